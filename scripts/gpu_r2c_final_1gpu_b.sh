set -x
timeout 900 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 300 gpurun_out/r2c_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_reference_arm.err
for c in 2 4 6; do timeout 600 python bench.py --config $c --steps 2 --warmup 3 --e2e-steps 3 --no-cpu > gpurun_out/r2c_bench_cfg${c}.json 2> gpurun_out/r2c_bench_cfg${c}.err; tail -c 200 gpurun_out/r2c_bench_cfg${c}.err; done
for c in 1 3; do timeout 600 python bench.py --config $c --steps 10 --warmup 8 --e2e-steps 5 --no-cpu > gpurun_out/r2c_bench_cfg${c}.json 2> gpurun_out/r2c_bench_cfg${c}.err; tail -c 200 gpurun_out/r2c_bench_cfg${c}.err; done
ncu --set full --clock-control none -k regex:asb_lists -s 14 -c 3 -o gpurun_out/r2c_lists_full -f python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2c_ncu_lists.log 2>&1
python profiles/summarize.py gpurun_out/r2c_lists_full.ncu-rep > gpurun_out/r2c_asb_lists_ncu_full.txt; rm -f gpurun_out/r2c_lists_full.ncu-rep
ncu --set full --clock-control none -k regex:asb_prune_rows -c 2 -o gpurun_out/r2c_prune_rows_full -f python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2c_ncu_prune.log 2>&1
python profiles/summarize.py gpurun_out/r2c_prune_rows_full.ncu-rep > gpurun_out/r2c_asb_prune_rows_ncu_full.txt; rm -f gpurun_out/r2c_prune_rows_full.ncu-rep
ncu --set full --clock-control none -k regex:asb_lists -s 8 -c 3 -o gpurun_out/r2c_lists_cfg2_full -f python bench.py --config 2 --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2c_ncu_lists_cfg2.log 2>&1
python profiles/summarize.py gpurun_out/r2c_lists_cfg2_full.ncu-rep > gpurun_out/r2c_asb_lists_cfg2_ncu_full.txt; rm -f gpurun_out/r2c_lists_cfg2_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches_cfg5.csv python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2c_ncu_launches.log 2>&1
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
python - <<'PY'
import json
for f in ["n1", "cfg1", "cfg2", "cfg3", "cfg4", "cfg6"]:
    try:
        d = json.load(open(f"gpurun_out/r2c_bench_{f}.json"))
        print(f, "value %.3f G  ms %.1f  e2e %.3f G  ms %.1f" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]), "frac %.3f useful %.3f" % (d["roofline"]["frac"], d["roofline"]["frac_useful"]), d.get("parity_check"), d.get("records_crc_check"), d["rank0_wall_ms_of_each_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
