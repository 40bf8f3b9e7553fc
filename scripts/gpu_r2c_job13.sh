set -x
for c in 4 6; do timeout 600 python bench.py --config $c --steps 2 --warmup 3 --e2e-steps 3 --no-cpu > gpurun_out/r2d_bench_cfg${c}.json 2> gpurun_out/r2d_bench_cfg${c}.err; tail -c 200 gpurun_out/r2d_bench_cfg${c}.err; done
python - <<'PY'
import json
for f in ["cfg4", "cfg6"]:
    try:
        d = json.load(open(f"gpurun_out/r2d_bench_{f}.json"))
        print(f, "value %.3f G  ms %.1f  e2e %.3f G  ms %.1f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]), d.get("parity_check"), d.get("records_crc_check"), d["rank0_wall_ms_of_each_step"], d["roofline"]["launches"])
    except Exception as e:
        print(f, "FAILED", e)
PY
