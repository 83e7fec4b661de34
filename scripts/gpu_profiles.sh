#!/bin/bash
# the ncu evidence kept under profiles/, re-taken on the final kernels of round 2 (thread-per-read tie pass, per-block index,
# more warps per SM in the short prefilter scans): launch lists of all five configs and full captures of their dominant kernels, reduced on the box
mkdir -p gpurun_out
for w in c1 c2 c3 c4 c5; do
  reads=16777216; [ $w = c5 ] && reads=4546560
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel' -c 64 --csv --log-file gpurun_out/profiles_launches_$w.csv \
      python bench.py --workload $w --reads $reads --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/profiles_list_$w.log 2>&1
done
cap() { # workload kernel-regex skip count reads
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/profiles_full_$1 -f \
      python bench.py --workload $1 --reads $5 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/profiles_full_$1.log 2>&1
  ncu -i /tmp/profiles_full_$1.ncu-rep --page raw --csv > gpurun_out/profiles_full_$1_raw.csv 2>/dev/null
  for k in $(seq 1 $4); do python scripts/ncu_lines.py /tmp/profiles_full_$1.ncu-rep $k 60 > gpurun_out/profiles_full_$1_lines_$k.txt 2>&1; done
}
cap c1 'pamld_fast_grid|pamld_grid_kernel|pamld_tie' 9 3 16777216
cap c2 'mdd_table' 3 1 16777216
cap c3 'pamld_fast_kernel|pamld_kernel|pamld_tie' 15 3 16777216
cap c4 'pamld_fast_kernel' 7 2 16777216
cap c5 'pamld_whitelist' 3 1 1136640
du -sh gpurun_out; ls gpurun_out | grep profiles | head -40
