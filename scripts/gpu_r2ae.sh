#!/bin/bash
# round 2, call AE: c1's ncu evidence once more after the separable scan stopped pushing candidate lists (TIE_MASKS)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel' -c 64 --csv --log-file gpurun_out/r2u_launches_c1.csv \
    python bench.py --workload c1 --reads 16777216 --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2u_list_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pamld_fast_grid|pamld_grid_kernel|pamld_tie' -s 9 -c 3 -o /tmp/r2u_full_c1 -f \
    python bench.py --workload c1 --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2u_full_c1.log 2>&1
ncu -i /tmp/r2u_full_c1.ncu-rep --page raw --csv > gpurun_out/r2u_full_c1_raw.csv 2>/dev/null
for k in 1 2 3; do python scripts/ncu_lines.py /tmp/r2u_full_c1.ncu-rep $k 60 > gpurun_out/r2u_full_c1_lines_$k.txt 2>&1; done
ls -la gpurun_out | grep r2u_full_c1
