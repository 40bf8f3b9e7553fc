set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python tests/perf_probe_variants.py 5 6 4 2 > gpurun_out/r2b_variants4.jsonl 2> gpurun_out/r2b_variants4.err; tail -c 400 gpurun_out/r2b_variants4.err
python - <<'PY'
import json
for l in open("gpurun_out/r2b_variants4.jsonl"):
    d = json.loads(l)
    print(d["cfg"], "%-14s" % d["variant"], "ms %.1f" % d["ms_per_job"], "crc", d["crc_ok"], "dev %.1f lists %.1f first %.2f" % (d["device_ms"], d["lists_ms"], d["screen_ms"]), "live", d["live"], "pruned", d["pruned"])
PY
