timeout 600 python bench.py --config 4 --steps 2 --warmup 2 --e2e-steps 2 --no-cpu --no-parity > gpurun_out/r2_bench_g_cfg4.json 2> gpurun_out/r2_bench_g_cfg4.err; tail -c 400 gpurun_out/r2_bench_g_cfg4.err
timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 2 --no-cpu --no-parity > gpurun_out/r2_bench_g_n1.json 2> gpurun_out/r2_bench_g_n1.err; tail -c 400 gpurun_out/r2_bench_g_n1.err
