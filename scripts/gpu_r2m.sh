#!/bin/bash
# round 2, call K: parity tests (short set), bench, launch lists of c1 pass 2 and c3
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_whitelist_depth.py ) > gpurun_out/r2m_tests.log 2>&1
tail -4 gpurun_out/r2m_tests.log
( time timeout 900 python bench.py --no-cpu-baseline --no-e2e ) > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -c 1000 gpurun_out/r2m_bench.err
python - <<'PY'
import json
try:
    line = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[0])
    print("c1", line["value"], line["roofline"]["kernel_ms_per_launch_set"], "pass2", line["two_pass"]["pass2"]["value"], line["two_pass"]["kernels_pass2"])
    for k, v in line.get("configs", {}).items():
        print(k, v.get("value"), v.get("error"), v.get("roofline", {}).get("kernel_ms_per_launch_set"))
        if "two_pass" in v: print("   pass2", v["two_pass"]["pass2"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pamld|mdd|count_kernel' -c 60 --csv --log-file gpurun_out/r2m_launches_c1.csv \
    python bench.py --workload c1 --reads 16777216 --steps 1 --warmup 3 --configs c3 --no-e2e --no-cpu-baseline > gpurun_out/r2m_ncu_c1.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2m_launches_c1.csv")) if len(r) > 5]
head = next(r for r in rows if "Kernel Name" in r)
k, v = head.index("Kernel Name"), head.index("Metric Value")
seen = {}
for r in rows[rows.index(head) + 1:]:
    seen.setdefault(r[k][:80], []).append(float(r[v].replace(",", "")) / 1000)
for name, times in seen.items():
    print("   %-80s n=%2d median %9.1f us" % (name, len(times), sorted(times)[len(times) // 2]))
PY
