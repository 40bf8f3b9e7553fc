set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_text.py tests/test_golden.py tests/test_groups.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 2 --warmup 2 --e2e-steps 3 --no-cpu > gpurun_out/r2_bench_f_n1.json 2> gpurun_out/r2_bench_f_n1.err; tail -c 400 gpurun_out/r2_bench_f_n1.err
for c in 4 2 3 1; do timeout 600 python bench.py --config $c --steps 2 --warmup 2 --e2e-steps 2 --no-cpu --no-parity > gpurun_out/r2_bench_f_cfg$c.json 2> gpurun_out/r2_bench_f_cfg$c.err; tail -c 400 gpurun_out/r2_bench_f_cfg$c.err; done
