set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 > gpurun_out/r2c_bench_n4.json 2> gpurun_out/r2c_bench_n4.err; tail -c 400 gpurun_out/r2c_bench_n4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_n4.json"))
print("n4 value %.3f G  ms %.1f  e2e %.3f G  ms %.1f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]), d.get("parity_check"), d.get("records_crc_check"), d["rank0_wall_ms_of_each_step"])
print(d["e2e"]["rank0_phases_ms_last_step"]); print(d["roofline"]["device_time_ms_per_step"])
PY
