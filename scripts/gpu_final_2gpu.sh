set -x
timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; tail -c 300 gpurun_out/r2_bench_n2_final.err
