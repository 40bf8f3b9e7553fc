set -x
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_reads or dead_zone or large_alphabet or degenerate or thresholds or scattered" 2>&1 | tail -4
timeout 600 python bench.py --no-cpu > gpurun_out/r2e_bench_n1_final.json 2> gpurun_out/r2e_bench_n1_final.err; tail -c 300 gpurun_out/r2e_bench_n1_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench_n1_final.json"))
print("value %.2f G  ms %.1f  e2e %.2f G  ms %.1f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]), d.get("parity_check"), d.get("records_crc_check"), d["roofline"]["frac"], d["roofline"]["frac_useful"])
PY
