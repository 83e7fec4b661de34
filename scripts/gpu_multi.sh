# usage: bash scripts/gpu_multi.sh N  -- the driver's launch line for N GPUs
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --reads $((1<<26)) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1800 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cut -c1-300 gpurun_out/bench_ref_n$N.json
