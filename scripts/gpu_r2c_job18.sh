set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_reads or baseline_configs or dead_zone" 2>&1 | tail -4
timeout 300 python bench.py --config 2 --steps 3 --warmup 2 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/r2e_cfg2_split.json 2> gpurun_out/r2e_cfg2_split.err || tail -3 gpurun_out/r2e_cfg2_split.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_cfg2_split.json"))
print("split value %.1f M  ms %.1f" % (d["value"] / 1e6, d["ms_per_step"]), d.get("records_crc_check"), d["rank0_wall_ms_of_each_step"], d["roofline"]["device_time_ms_per_step"])
PY
