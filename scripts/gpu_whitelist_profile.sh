# c5: bench, then the ncu launch list (per kernel durations) and one full capture of the whitelist scan
mkdir -p gpurun_out
T=${1:-r01}
( time timeout 600 python bench.py --workload c5 --steps 3 --no-e2e --no-cpu-baseline ) > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-400 gpurun_out/bench_c5.json; tail -4 gpurun_out/bench_c5.err
timeout 900 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_${T}.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_c5.log 2>&1; tail -2 gpurun_out/ncu_launch_c5.log
grep -v "^==" gpurun_out/launches_c5_${T}.csv | cut -d, -f5,12- | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pamld_whitelist -s 3 -c 1 -o gpurun_out/prof_c5_${T} python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c5.log 2>&1; tail -2 gpurun_out/ncu_full_c5.log
