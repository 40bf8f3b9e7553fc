set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --config 4 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu > gpurun_out/r2d_bench_cfg4_n2.json 2> gpurun_out/r2d_bench_cfg4_n2.err; tail -c 300 gpurun_out/r2d_bench_cfg4_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d_bench_cfg4_n2.json"))
print("cfg4 n2 value %.3f G  ms %.1f  e2e %.3f G  ms %.1f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]), d.get("parity_check"), d.get("records_crc_check"), d["roofline"]["launches"])
PY
