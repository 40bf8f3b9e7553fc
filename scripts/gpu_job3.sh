set -x
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 3 --cpu-seconds 10 > gpurun_out/r2_bench_c_n1.json 2> gpurun_out/r2_bench_c_n1.err; tail -c 600 gpurun_out/r2_bench_c_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 3 > gpurun_out/r2_bench_c_n2.json 2> gpurun_out/r2_bench_c_n2.err; tail -c 1500 gpurun_out/r2_bench_c_n2.err
timeout 600 python bench.py --config 4 --steps 2 --warmup 2 --e2e-steps 2 --no-cpu > gpurun_out/r2_bench_c_cfg4.json 2> gpurun_out/r2_bench_c_cfg4.err; tail -c 300 gpurun_out/r2_bench_c_cfg4.err
