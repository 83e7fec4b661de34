#!/bin/bash
# round 2, call L: experiment — the bit-sliced whitelist scan forced onto the small codecs of c3 / c4
mkdir -p gpurun_out
PHQ_WHITELIST_MINIMUM=1 timeout 600 python bench.py --workload c3 --configs c4,c1 --no-e2e --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -c 600 gpurun_out/r2l_bench.err
python - <<'PY'
import json
line = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[0])
print("c3", line["value"], line["roofline"]["kernel_ms_per_launch_set"], line["roofline"]["kernel"])
for k, v in line.get("configs", {}).items():
    print(k, v.get("value"), v.get("error"), v.get("roofline", {}).get("kernel_ms_per_launch_set"), v.get("roofline", {}).get("kernel"))
PY
