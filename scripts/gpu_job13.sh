for v in H J K; do
  ASB200_LIB=$PWD/build_variants/lib$v.so timeout 300 python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_var_$v.json 2> gpurun_out/r2_var_$v.err || tail -3 gpurun_out/r2_var_$v.err
  ASB200_LIB=$PWD/build_variants/lib$v.so timeout 300 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu --no-parity --prune 0 --reads 30000 > gpurun_out/r2_var_${v}_screen.json 2> gpurun_out/r2_var_${v}_screen.err || tail -3 gpurun_out/r2_var_${v}_screen.err
  ASB200_LIB=$PWD/build_variants/lib$v.so timeout 300 python bench.py --config 2 --steps 2 --warmup 1 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_var_${v}_cfg2.json 2> gpurun_out/r2_var_${v}_cfg2.err || tail -3 gpurun_out/r2_var_${v}_cfg2.err
done
