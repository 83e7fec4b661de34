# usage: bash scripts/gpu_profile.sh <workload> <kernel-regex> <tag>   -- launch list + one full ncu capture
W=${1:-c1}; K=${2:-pamld_kernel}; T=${3:-r01}
mkdir -p gpurun_out
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${W}_${T}.csv python bench.py --workload $W --steps 2 --warmup 3 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${W}.log 2>&1; tail -2 gpurun_out/ncu_launch_${W}.log
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${W}_${T} python bench.py --workload $W --steps 1 --warmup 3 --reads $((1<<24)) --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${W}.log 2>&1; tail -2 gpurun_out/ncu_full_${W}.log
