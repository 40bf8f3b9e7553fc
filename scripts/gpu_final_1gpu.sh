set -x
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; tail -c 300 gpurun_out/r2_bench_n1_final.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
for c in 1 2 3 4 6; do timeout 600 python bench.py --config $c --steps 2 --warmup 3 --e2e-steps 3 --no-cpu > gpurun_out/r2_bench_cfg${c}_final.json 2> gpurun_out/r2_bench_cfg${c}_final.err; tail -c 300 gpurun_out/r2_bench_cfg${c}_final.err; done
ncu --set full --clock-control none --import-source on -k regex:asb_lists -s 14 -c 3 -o gpurun_out/r2_lists_full_v2 -f python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_ncu_lists_v2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_cfg5_final_v2.csv python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_ncu_launches4.log 2>&1
