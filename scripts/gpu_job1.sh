set -x
for pc in 27 28 29; do python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu --no-parity --pair-cap $((1<<pc)) > gpurun_out/r2_pc$pc.json 2> gpurun_out/r2_pc$pc.err; done
for c in 1 2 3 4; do python bench.py --config $c --steps 2 --warmup 3 --e2e-steps 3 --cpu-seconds 5 > gpurun_out/r2_cfg$c.json 2> gpurun_out/r2_cfg$c.err; done
ncu --set full --clock-control none --import-source on -k regex:asb_screen -s 20 -c 1 -o gpurun_out/r2_screen_full -f python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu --no-parity > gpurun_out/r2_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --steps 1 --warmup 0 --e2e-steps 1 --no-cpu --no-parity > gpurun_out/r2_ncu_launches.log 2>&1
ls -la gpurun_out | tail -20
