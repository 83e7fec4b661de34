#!/bin/bash
# round 2, call C: per-kernel launch lists (ncu, durations only) of c1 / c3 / c4 with the prefilter scans, and full captures of the two prefilter kernels
mkdir -p gpurun_out
for w in c1 c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches_$w.csv \
      python bench.py --workload $w --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2c_ncu_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pamld_fast -c 2 -o gpurun_out/r2c_fast_c1 -f \
    python bench.py --workload c1 --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2c_full_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pamld_fast -c 3 -o gpurun_out/r2c_fast_c3 -f \
    python bench.py --workload c3 --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2c_full_c3.log 2>&1
ls -la gpurun_out | tail -12
