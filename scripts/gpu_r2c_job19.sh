set -x
for t in 4 8; do
  ASB200_WRITER_THREADS=$t timeout 200 python bench.py --config 4 --steps 1 --warmup 1 --e2e-steps 2 --no-cpu --no-parity > gpurun_out/r2e_cfg4_w$t.json 2> gpurun_out/r2e_cfg4_w$t.err || tail -3 gpurun_out/r2e_cfg4_w$t.err
done
python - <<'PY'
import json
for t in (4, 8):
    try:
        d = json.load(open(f"gpurun_out/r2e_cfg4_w{t}.json"))
        print("writers", t, "e2e ms %.1f" % d["e2e"]["ms_per_step"], d.get("records_crc_check"), {k: v for k, v in d["e2e"]["rank0_phases_ms_last_step"].items() if "text" in k or "drain" in k or "slabs" in k})
    except Exception as e:
        print(t, "FAILED", e)
PY
