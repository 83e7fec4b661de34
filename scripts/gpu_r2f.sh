#!/bin/bash
# round 2, call F: full ncu captures of the tie pass and the exact scan (index-list mode) on c1 and c3; quick bench of c1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pamld_tie_kernel|pamld_grid_kernel' -c 2 -o gpurun_out/r2f_tie_c1 -f \
    python bench.py --workload c1 --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2f_full_c1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pamld_tie_kernel|pamld_kernel' -c 2 -o gpurun_out/r2f_tie_c3 -f \
    python bench.py --workload c3 --reads 16777216 --steps 1 --warmup 3 --configs '' --no-e2e --no-cpu-baseline > gpurun_out/r2f_full_c3.log 2>&1
python bench.py --workload c1 --configs c3 --no-e2e --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
line = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[0])
print("c1", line["value"], line["roofline"]["kernel_ms_per_launch_set"])
for k, v in line.get("configs", {}).items():
    print(k, v.get("value"), v.get("error"), v.get("roofline", {}).get("kernel_ms_per_launch_set"))
PY
