# default bench (as the driver runs it) + the reference arm + the other workloads briefly
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2600 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_default.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json; tail -4 gpurun_out/bench_reference.err
for w in c2 c3 c4; do python bench.py --workload $w --steps 5 --reads $((1<<24)) --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cut -c1-900 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
