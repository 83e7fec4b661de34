#!/bin/bash
# round 2, call H (2 GPUs): the collective on hardware — the 2-GPU collect test, then the full bench line at N = 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_collect.py -m gpu -q -x ) > gpurun_out/r2h_tests.log 2>&1
tail -8 gpurun_out/r2h_tests.log
( time timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -c 1500 gpurun_out/r2h_bench_n2.err
python - <<'PY'
import json
try:
    line = json.loads([l for l in open("gpurun_out/r2h_bench_n2.json").read().strip().splitlines() if l.startswith("{")][0])
    print("c1", line["value"], "collect", line.get("collect"), "verified", line.get("verified"), "e2e", line["e2e"]["value"], line["e2e"]["frac_of_copy_ceiling"])
    for k, v in line.get("configs", {}).items():
        print(k, v.get("value"), v.get("error"), v.get("collect"), v.get("verified"))
        if "two_pass" in v: print("   two_pass", {a: b for a, b in v["two_pass"].items() if a != "pass2"}, "pass2", v["two_pass"]["pass2"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
