mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -4
for w in c3 c4; do python bench.py --workload $w --steps 5 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cut -c1-200 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
python bench.py --workload c1 --steps 5 --reads $((1<<26)) --no-cpu-baseline > gpurun_out/bench_c1_tags.json 2> gpurun_out/bench_c1_tags.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c1_tags.json')); print({k: d[k]['value'] for k in d if k.startswith('e2e')}, d['value'])"; tail -3 gpurun_out/bench_c1_tags.err
timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_final.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_c5.log 2>&1
grep -v "^==" gpurun_out/launches_c5_final.csv | cut -d, -f5,12- | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pamld_whitelist -s 3 -c 1 -o gpurun_out/prof_c5_final python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5.log 2>&1; tail -1 gpurun_out/ncu_full_c5.log | cut -c1-200
