set -x
ncu --set full --clock-control none --import-source on -k regex:asb_lists -s 14 -c 3 -o gpurun_out/r2_lists_full -f python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_ncu_lists.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_cfg5_final.csv python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_ncu_launches3.log 2>&1
ncu --set full --clock-control none -k regex:asb_prune -s 2 -c 1 -o gpurun_out/r2_prune_full -f python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_ncu_prune.log 2>&1
timeout 900 python tests/perf_reference_script.py --cfg2-scale 0.5 > gpurun_out/r2_reference_script.json 2> gpurun_out/r2_reference_script.err; tail -c 300 gpurun_out/r2_reference_script.err; cat gpurun_out/r2_reference_script.json
nproc; lscpu | head -20
