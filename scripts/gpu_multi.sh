#!/bin/bash
# N GPUs (N = $1; gpurun --gpus N): the full bench line under torchrun (NCCL) — every config, collect times, verified sums, two-pass workflows
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/multi_topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_collect.py -m gpu -q -x ) > gpurun_out/multi_tests_n2.log 2>&1
  tail -3 gpurun_out/multi_tests_n2.log
fi
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err
tail -c 800 gpurun_out/multi_bench_n$N.err
python scripts/bench_table.py gpurun_out/multi_bench_n$N.json | cut -c1-220
