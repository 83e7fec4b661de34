# what the driver runs at round end (tests, smoke, both bench arms) + the other workloads and the final whitelist profile
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-330 gpurun_out/bench_reference.json; tail -4 gpurun_out/bench_reference.err
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['roofline']['frac'], {k: d[k]['value'] for k in d if k.startswith('e2e')}, d['cpu_baseline']['value'], d['clocks'])"; tail -4 gpurun_out/bench_default.err
( time python bench.py --workload c5 --steps 3 --no-e2e ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print(d['value'], d['roofline']['pair_words_per_clk_per_sm'], d['cpu_baseline'])"; tail -3 gpurun_out/bench_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pamld_whitelist -s 3 -c 1 -o gpurun_out/prof_c5_final2 python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5.log 2>&1; tail -1 gpurun_out/ncu_full_c5.log | cut -c1-120
