set -x
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -16
timeout 900 python tests/perf_probe_variants.py 5 2 3 6 4 1 > gpurun_out/r2b_variants2.jsonl 2> gpurun_out/r2b_variants2.err; tail -c 400 gpurun_out/r2b_variants2.err; cat gpurun_out/r2b_variants2.jsonl
ASB200_TRACE=1 timeout 600 python tests/perf_probe_shard.py 8 0 > gpurun_out/r2b_shard8_2.jsonl 2> gpurun_out/r2b_shard8_2.err; tail -n 30 gpurun_out/r2b_shard8_2.err; cat gpurun_out/r2b_shard8_2.jsonl
