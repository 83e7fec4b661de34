#!/bin/bash
# round 2, call S: 32 warps per SM in the short uniform prefilter scan — parity tests, c3 / c4 bench lines and the c3 launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_whitelist_depth.py ) > gpurun_out/r2s_tests.log 2>&1
tail -3 gpurun_out/r2s_tests.log
for w in c3 c4; do
  timeout 600 python bench.py --workload $w --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2s_bench_$w.json 2> gpurun_out/r2s_bench_$w.err
  python scripts/bench_table.py gpurun_out/r2s_bench_$w.json | cut -c1-160 | tail -n +4 | head -2
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel' -c 40 --csv --log-file gpurun_out/r2s_launches_c3.csv \
    python bench.py --workload c3 --reads 16777216 --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2s_list_c3.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2s_launches_c3.csv")) if len(r) > 5]
head = next(r for r in rows if "Kernel Name" in r)
k, v = head.index("Kernel Name"), head.index("Metric Value")
seen = {}
for r in rows[rows.index(head) + 1:]:
    seen.setdefault(r[k][:70], []).append(float(r[v].replace(",", "")) / 1000)
for name, times in seen.items():
    print("   %-70s n=%2d median %9.1f us  all %s" % (name, len(times), sorted(times)[len(times) // 2], " ".join("%.0f" % t for t in times[:8])))
PY
