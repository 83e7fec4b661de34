"""The N > 1 path on CPU: world size 2 over gloo. Reads are sharded per rank (shard_range), each rank
accumulates its shard, the accumulator planes are all-reduced with the SAME function the NCCL path uses
(all_reduce_accumulators), and the result must equal a single-process run over all reads; prior estimation
from the reduced tables must match too. The per-shard tables come from the CPU oracle (no GPU here)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, n_reads, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from pheniqs_b200 import all_reduce_accumulators, shard_range, workload
    spec = workload.load("c4")
    compiled = O.compile_job(spec["job"])
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], n_reads, seed=31, sampling="zipf")
    begin, end = shard_range(n_reads, rank, world)
    keep = np.zeros(n_reads, dtype=bool)
    keep[begin:end] = True
    shard = O.ReadBatch(code, quality, offset).select(keep)
    oracle = O.PortOracle(compiled)
    oracle.decode(shard)
    planes_u, planes_f = [], []
    for k in range(oracle.n_decoders):
        u, f = oracle.accumulators(k)
        planes_u.append(u.reshape(-1))
        planes_f.append(f.reshape(-1))
    count, pf_count = oracle.totals()
    u64 = torch.from_numpy(np.concatenate(planes_u + [np.array([count, pf_count], dtype=np.uint64)]).view(np.int64).copy())
    f64 = torch.from_numpy(np.concatenate(planes_f).copy())
    all_reduce_accumulators(u64, f64)
    if rank == 0:
        np.save(os.path.join(out_dir, "u64.npy"), u64.numpy().view(np.uint64))
        np.save(os.path.join(out_dir, "f64.npy"), f64.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    n_reads, world, port = 6000, 2, 29731
    mp.spawn(_worker, args=(world, port, n_reads, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    from pheniqs_b200 import shard_range, workload
    spec = workload.load("c4")
    compiled = O.compile_job(spec["job"])
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], n_reads, seed=31, sampling="zipf")
    whole = O.PortOracle(compiled)
    whole.decode(O.ReadBatch(code, quality, offset))
    u64 = np.load(tmp_path / "u64.npy")
    f64 = np.load(tmp_path / "f64.npy")
    at_u = at_f = 0
    for k in range(whole.n_decoders):
        u, f = whole.accumulators(k)
        got_u = u64[at_u:at_u + u.size].reshape(u.shape)
        got_f = f64[at_f:at_f + f.size].reshape(f.shape)
        assert np.array_equal(got_u, u)
        assert np.allclose(got_f, f, rtol=1e-12, atol=0)
        if whole.chain[k][1]["algorithm"] == "pamld":
            noise, concentration = whole.estimate_priors(k, (got_u, got_f))
            want_noise, want_concentration = whole.estimate_priors(k)
            assert noise == want_noise and np.array_equal(concentration, want_concentration)
        at_u += u.size
        at_f += f.size
    assert tuple(u64[at_u:at_u + 2]) == whole.totals()
    assert [shard_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
