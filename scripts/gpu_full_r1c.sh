# everything the driver runs at round end, plus the other workloads
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json; tail -4 gpurun_out/bench_reference.err
for w in c2 c3 c4; do python bench.py --workload $w --steps 5 --reads $((1<<24)) --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cut -c1-200 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
( time python bench.py --workload c5 --steps 3 --no-e2e ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-200 gpurun_out/bench_c5.json; tail -c 600 gpurun_out/bench_c5.json; tail -4 gpurun_out/bench_c5.err
