#!/bin/bash
# the loop of a kernel change: parity tests (short set), the quick bench line, ncu launch lists of c1 / c3 / c4
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_whitelist_depth.py ) > gpurun_out/iterate_tests.log 2>&1
tail -5 gpurun_out/iterate_tests.log
( time timeout 900 python bench.py --no-cpu-baseline --no-e2e ) > gpurun_out/iterate_bench.json 2> gpurun_out/iterate_bench.err
tail -c 600 gpurun_out/iterate_bench.err
python scripts/bench_table.py gpurun_out/iterate_bench.json | cut -c1-200
for w in c1 c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel' -c 40 --csv --log-file gpurun_out/iterate_launches_$w.csv \
      python bench.py --workload $w --reads 16777216 --steps 2 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/iterate_list_$w.log 2>&1
  python - $w <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/iterate_launches_%s.csv" % sys.argv[1])) if len(r) > 5]
head = next(r for r in rows if "Kernel Name" in r)
k, v = head.index("Kernel Name"), head.index("Metric Value")
seen = {}
for r in rows[rows.index(head) + 1:]:
    seen.setdefault(r[k][:70], []).append(float(r[v].replace(",", "")) / 1000)
for name, times in seen.items():
    print("   %-70s n=%2d median %9.1f us  all %s" % (name, len(times), sorted(times)[len(times) // 2], " ".join("%.0f" % t for t in times[:8])))
PY
done
