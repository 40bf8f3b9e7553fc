for v in A B C D E; do
  ASB200_LIB=$PWD/build_variants/lib$v.so timeout 300 python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_var_$v.json 2> gpurun_out/r2_var_$v.err || tail -3 gpurun_out/r2_var_$v.err
done
