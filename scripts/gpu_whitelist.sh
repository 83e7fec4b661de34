# whitelist kernel: parity tests, then c5 bench (no e2e) and its launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${1:-whitelist}" 2>&1 | tail -15
( time timeout 600 python bench.py --workload c5 --steps 3 --no-e2e --no-cpu-baseline ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-300 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_${2:-r01c}.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_c5.log 2>&1; tail -1 gpurun_out/ncu_launch_c5.log | cut -c1-200
grep -v "^==" gpurun_out/launches_c5_${2:-r01c}.csv | cut -d, -f5,12- | tail -3
