mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --workload c2 --steps 5 --reads $((1<<26)) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1800 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
