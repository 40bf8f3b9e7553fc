set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; tail -c 600 gpurun_out/r2_bench_n8_final.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err; tail -c 300 gpurun_out/r2_bench_n4_final.err
