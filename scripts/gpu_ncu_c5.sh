mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pamld_whitelist -s 3 -c 1 -o gpurun_out/prof_c5_${1:-r01b} python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5.log 2>&1; tail -2 gpurun_out/ncu_full_c5.log | cut -c1-200
