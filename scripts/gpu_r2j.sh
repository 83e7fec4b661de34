#!/bin/bash
# round 2, call E: all GPU tests (with durations), bench, per-kernel launch lists
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2j_tests.log 2>&1
tail -40 gpurun_out/r2j_tests.log
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 1500 gpurun_out/r2j_bench.err
python - <<'PY'
import json
try:
    line = json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[0])
    print("c1", line["value"], line["roofline"]["kernel_ms_per_launch_set"], "e2e", line["e2e"]["value"], line["e2e"]["frac_of_copy_ceiling"])
    for k, v in line.get("configs", {}).items():
        print(k, v.get("value"), v.get("error"), v.get("roofline", {}).get("kernel_ms_per_launch_set"))
        if "two_pass" in v: print("   pass2", v["two_pass"]["pass2"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
for w in c1 c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel|pack_kernel|tag_kernel' -c 40 --csv --log-file gpurun_out/r2j_launches_$w.csv \
      python bench.py --workload $w --reads 4194304 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2j_ncu_$w.log 2>&1
done
python - <<'PY'
import csv, glob
for path in sorted(glob.glob("gpurun_out/r2j_launches_*.csv")):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    head = next(r for r in rows if "Kernel Name" in r)
    k, v = head.index("Kernel Name"), head.index("Metric Value")
    print(path)
    for r in rows[rows.index(head) + 1:][-10:]:
        print("   %-70s %10.1f us" % (r[k][:70], float(r[v].replace(",", "")) / 1000))
PY
