# usage: bash scripts/gpu_profile_tie.sh <tag>   -- one full ncu capture (with source counters) of the tie kernel on c1
T=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pamld_tie -s 3 -c 1 -o gpurun_out/prof_tie_${T} python bench.py --workload c1 --steps 1 --warmup 3 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_tie.log 2>&1; tail -2 gpurun_out/ncu_full_tie.log
