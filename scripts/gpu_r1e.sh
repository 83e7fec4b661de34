mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_c1_quick.json 2> gpurun_out/bench_c1_quick.err; cut -c1-260 gpurun_out/bench_c1_quick.json; tail -2 gpurun_out/bench_c1_quick.err
for w in c3 c4; do python bench.py --workload $w --steps 5 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cut -c1-200 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1_r01b.csv python bench.py --steps 2 --warmup 3 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_c1.log 2>&1
grep -v "^==" gpurun_out/launches_c1_r01b.csv | cut -d, -f5,12- | tail -4
