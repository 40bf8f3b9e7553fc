set -x
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_pipeline_contract.py 2>&1 | tail -25
timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 2 --cpu-seconds 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 600 gpurun_out/r2_bench_b.err
timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu --no-parity --prune 0 > gpurun_out/r2_bench_b_noprune.json 2> gpurun_out/r2_bench_b_noprune.err; tail -c 300 gpurun_out/r2_bench_b_noprune.err
timeout 600 python bench.py --config 4 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu > gpurun_out/r2_bench_b_cfg4.json 2> gpurun_out/r2_bench_b_cfg4.err; tail -c 300 gpurun_out/r2_bench_b_cfg4.err
