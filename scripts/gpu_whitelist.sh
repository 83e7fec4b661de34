# whitelist kernel: parity tests, then c5 bench (no e2e), then c3 with the whitelist kernel forced
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "whitelist" 2>&1 | tail -15
( time timeout 600 python bench.py --workload c5 --steps 3 --no-e2e --no-cpu-baseline ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-2500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
