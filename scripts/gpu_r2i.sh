#!/bin/bash
# round 2, call I: whitelist scan (c5) launch list and full capture; tie pass full captures on c3
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|count_kernel' -c 24 --csv --log-file gpurun_out/r2i_launches_c5.csv \
    python bench.py --workload c5 --reads 4194304 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2i_ncu_c5.log 2>&1
tail -2 gpurun_out/r2i_ncu_c5.log | cut -c1-400
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'pamld_whitelist|pamld_tie' -s 6 -c 2 -o gpurun_out/r2i_whitelist_c5 -f \
    python bench.py --workload c5 --reads 1136640 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2i_full_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pamld_tie_kernel' -c 1 -o gpurun_out/r2i_tie_c3 -f \
    python bench.py --workload c3 --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2i_full_c3.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2i_launches_c5.csv")) if len(r) > 5]
head = next(r for r in rows if "Kernel Name" in r)
k, v = head.index("Kernel Name"), head.index("Metric Value")
for r in rows[rows.index(head) + 1:][-8:]:
    print("   %-70s %10.1f us" % (r[k][:70], float(r[v].replace(",", "")) / 1000))
PY
ls -la gpurun_out | grep r2i
