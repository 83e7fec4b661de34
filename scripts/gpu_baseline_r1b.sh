# microbenchmark + current numbers of c3 c4 c5 (kernel descriptions included in the JSON)
mkdir -p gpurun_out
./scripts/microbench.bin > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt
for w in c3 c4; do python bench.py --workload $w --steps 5 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cut -c1-1800 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
( time python bench.py --workload c5 --steps 3 --reads 113664 --no-e2e ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-2500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
