set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_consensus_stage.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python tests/perf_probe_variants.py 5 2 3 6 4 1 > gpurun_out/r2b_variants.jsonl 2> gpurun_out/r2b_variants.err; tail -c 400 gpurun_out/r2b_variants.err; cat gpurun_out/r2b_variants.jsonl
timeout 600 python tests/perf_probe_shard.py 8 0 > gpurun_out/r2b_shard8.jsonl 2> gpurun_out/r2b_shard8.err; tail -c 300 gpurun_out/r2b_shard8.err; cat gpurun_out/r2b_shard8.jsonl
