# usage: bash scripts/gpu_tests_bench.sh [reads_log2]   -- GPU tests, then a short bench of c1
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --reads $((1<<${1:-26})) > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 2500 gpurun_out/bench_c1.json; tail -5 gpurun_out/bench_c1.err
