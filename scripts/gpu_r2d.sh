#!/bin/bash
# round 2, call D: per-kernel launch lists (ncu, durations only, product kernels only)
mkdir -p gpurun_out
for w in c1 c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel|pack_kernel|tag_kernel' -c 40 --csv --log-file gpurun_out/r2d_launches_$w.csv \
      python bench.py --workload $w --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2d_ncu_$w.log 2>&1
  tail -2 gpurun_out/r2d_ncu_$w.log | cut -c1-300
done
python - <<'PY'
import csv, glob
for path in sorted(glob.glob("gpurun_out/r2d_launches_*.csv")):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = next(r for r in rows if "Kernel Name" in r)
    k, v = head.index("Kernel Name"), head.index("Metric Value")
    print(path)
    for r in rows[rows.index(head) + 1:][-14:]:
        print("   %-70s %s" % (r[k][:70], r[v]))
PY
