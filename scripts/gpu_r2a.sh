#!/bin/bash
# round 2, call A: GPU parity tests, then the default bench (all five configs in one line)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2a_topo.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_tests.log 2>&1
tail -5 gpurun_out/r2a_tests.log
( time timeout 900 python bench.py ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    line = json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[0])
    print("c1", line["value"], "e2e", line["e2e"]["value"], line["e2e"]["frac_of_copy_ceiling"])
    for k, v in line.get("configs", {}).items():
        print(k, v.get("value"), v.get("error"), v.get("roofline", {}).get("kernel_ms_per_launch_set"))
except Exception as e:
    print("bench parse failed", e)
PY
