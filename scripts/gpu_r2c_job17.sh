set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_reads or baseline_configs or dead_zone" 2>&1 | tail -4
timeout 300 python bench.py --config 2 --steps 3 --warmup 2 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/r2e_cfg2_split.json 2> gpurun_out/r2e_cfg2_split.err || tail -3 gpurun_out/r2e_cfg2_split.err
ASB200_LIB=$PWD/build_variants/libNOSPLIT.so timeout 300 python bench.py --config 2 --steps 3 --warmup 2 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2e_cfg2_nosplit.json 2> gpurun_out/r2e_cfg2_nosplit.err || tail -3 gpurun_out/r2e_cfg2_nosplit.err
python - <<'PY'
import json
for v in ("split", "nosplit"):
    try:
        d = json.load(open(f"gpurun_out/r2e_cfg2_{v}.json"))
        print(v, "value %.1f M  ms %.1f" % (d["value"] / 1e6, d["ms_per_step"]), d.get("records_crc_check"), d["rank0_wall_ms_of_each_step"], d["roofline"]["device_time_ms_per_step"])
    except Exception as e:
        print(v, "FAILED", e)
PY
