mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --reads $((1<<26)) > gpurun_out/bench_c1_first.json 2> gpurun_out/bench_c1_first.err; tail -c 3000 gpurun_out/bench_c1_first.json; tail -5 gpurun_out/bench_c1_first.err
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1.csv python bench.py --steps 2 --warmup 3 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:pamld_kernel -s 3 -c 1 -o gpurun_out/prof_pamld_c1 python bench.py --steps 1 --warmup 3 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
