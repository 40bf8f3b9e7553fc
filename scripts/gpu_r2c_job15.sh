set -x
for v in 0 3 4; do
  ASB200_LIB=$PWD/build_variants/libW$v.so timeout 300 python bench.py --config 2 --steps 2 --warmup 2 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/r2d_w$v.json 2> gpurun_out/r2d_w$v.err || tail -3 gpurun_out/r2d_w$v.err
done
python - <<'PY'
import json
for v in (0, 3, 4):
    try:
        d = json.load(open(f"gpurun_out/r2d_w{v}.json"))
        print("W32 =", v, "value %.1f M  ms %.1f" % (d["value"] / 1e6, d["ms_per_step"]), d.get("records_crc_check"), d["rank0_wall_ms_of_each_step"])
    except Exception as e:
        print(v, "FAILED", e)
PY
