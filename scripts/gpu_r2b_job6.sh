set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_text.py tests/test_groups.py -m gpu -x -q -k "scattered or golden or text or groups or degenerate" 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 3 --e2e-steps 3 --no-cpu > gpurun_out/r2b_bench_n1_quick3.json 2> gpurun_out/r2b_bench_n1_quick3.err; tail -c 300 gpurun_out/r2b_bench_n1_quick3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2b_bench_n1_quick3.json"))
print("value %.2f G  ms %.1f  e2e %.2f G  ms %.1f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]))
print(d["e2e"]["rank0_phases_ms_last_step"])
print(d.get("parity_check"), d.get("records_crc_check"))
PY
