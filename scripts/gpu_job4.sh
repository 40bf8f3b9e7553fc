set -x
timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 3 --no-cpu > gpurun_out/r2_bench_d_n1.json 2> gpurun_out/r2_bench_d_n1.err; tail -c 400 gpurun_out/r2_bench_d_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_cfg5_prune.csv python bench.py --steps 1 --warmup 0 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/r2_ncu_launches2.log 2>&1
