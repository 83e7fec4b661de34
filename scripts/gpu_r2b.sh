#!/bin/bash
# round 2, call B: parity with the prefilter scans, then the bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_tests.log 2>&1
tail -25 gpurun_out/r2b_tests.log
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 2000 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    line = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[0])
    print("c1", line["value"], line["roofline"]["kernel"], "e2e", line["e2e"]["value"], line["e2e"]["frac_of_copy_ceiling"])
    for k, v in line.get("configs", {}).items():
        print(k, v.get("value"), v.get("error"), v.get("roofline", {}).get("kernel_ms_per_launch_set"), v.get("roofline", {}).get("kernel"))
        if "two_pass" in v: print("   pass2", v["two_pass"]["pass2"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
