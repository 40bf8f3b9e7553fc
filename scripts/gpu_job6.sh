set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_text.py tests/test_gpu_fullsize.py -m gpu -x -q --durations=6 2>&1 | tail -20
timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 3 --no-cpu > gpurun_out/r2_bench_e_n1.json 2> gpurun_out/r2_bench_e_n1.err; tail -c 400 gpurun_out/r2_bench_e_n1.err
timeout 600 python bench.py --config 4 --steps 2 --warmup 2 --e2e-steps 2 --no-cpu --no-parity > gpurun_out/r2_bench_e_cfg4.json 2> gpurun_out/r2_bench_e_cfg4.err; tail -c 400 gpurun_out/r2_bench_e_cfg4.err
timeout 600 python bench.py --config 2 --steps 2 --warmup 2 --e2e-steps 2 --no-cpu --no-parity > gpurun_out/r2_bench_e_cfg2.json 2> gpurun_out/r2_bench_e_cfg2.err; tail -c 400 gpurun_out/r2_bench_e_cfg2.err
