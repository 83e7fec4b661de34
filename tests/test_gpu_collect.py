"""The path's one collective on hardware: phq_collect (one grouped in-place ncclAllReduce over the two accumulator
planes, classifier.h:87-93 / transcode.cpp:162-179 across GPUs) on two B200s over NCCL.

Two ranks decode the two halves of one synthetic batch (config 4: two PAMLD decoders, Zipf sampling, so the priors are
worth estimating), collect, and every rank's DEVICE tables must then equal (a) a one-GPU run over the whole batch —
integers bit for bit, f64 sums to 1e-12 — and (b) the CPU oracle; the priors estimated from the collected tables must
match on both ranks, and the second pass (priors installed, tables reset) must again match the oracle run with the
oracle's own estimates. Skipped with fewer than two devices."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

N_READS = 30000


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pheniqs_b200 import DecoderChain, compile_job, shard_range, workload
    from pheniqs_b200.binding import PheniqsError
    spec = workload.load("c4")
    compiled = compile_job(spec["job"])
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], N_READS, seed=41, sampling="zipf")
    begin, end = shard_range(N_READS, rank, world)
    part_offset = [o[begin:end + 1] - o[begin] for o in offset]
    part_code = [c[o[begin]:o[end]] for c, o in zip(code, offset)]
    part_quality = [q[o[begin]:o[end]] for q, o in zip(quality, offset)]
    chain = DecoderChain(compiled, device=rank)
    out = {}
    for pass_index in (1, 2):
        tiles = chain.pack(part_code, part_quality, part_offset)
        chain.decode(tiles, end - begin)
        chain.collect()
        torch.cuda.synchronize()
        refused = False
        try:
            chain.decode(tiles, end - begin)
        except PheniqsError:
            refused = True          # collected tables must not accumulate again before a reset
        assert refused
        tables = [chain.accumulators(k) for k in range(chain.n_decoders)]
        out["u%d" % pass_index] = np.concatenate([u.reshape(-1) for u, _ in tables])
        out["f%d" % pass_index] = np.concatenate([f.reshape(-1) for _, f in tables])
        out["totals%d" % pass_index] = np.array(chain.totals(), dtype=np.uint64)
        if pass_index == 1:
            priors = [chain.estimate_priors(k) for k, info in enumerate(chain.info) if info.algorithm == 0]
            out["noise"] = np.array([p[0] for p in priors])
            out["concentration"] = np.concatenate([p[1] for p in priors])
            chain.adjust_priors()
            chain.reset()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_collect_equals_one_gpu_and_the_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29741, str(tmp_path)), nprocs=2, join=True)
    from oracle import oracle as O
    from pheniqs_b200 import DecoderChain, compile_job, workload
    spec = workload.load("c4")
    compiled = compile_job(spec["job"])
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], N_READS, seed=41, sampling="zipf")
    ranks = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(2)]
    for key in ranks[0].files:
        if key.startswith("f"):
            assert np.allclose(ranks[0][key], ranks[1][key], rtol=0, atol=0), key       # an all-reduce leaves every rank with the same bits
        else:
            assert np.array_equal(ranks[0][key], ranks[1][key]), key

    # one GPU over the whole batch, and the CPU oracle, pass 1
    single = DecoderChain(compiled, device=0)
    single.decode(single.pack(code, quality, offset), N_READS)
    checker = O.best_oracle(compiled, len(code))
    checker.decode(O.ReadBatch(code, quality, offset))

    def compare(chain, oracle, got_u, got_f, got_totals):
        at_u = at_f = 0
        for k in range(chain.n_decoders):
            u, f = chain.accumulators(k)
            eu, ef = oracle.accumulators(k)
            assert np.array_equal(got_u[at_u:at_u + u.size].reshape(u.shape), u), k
            assert np.array_equal(u, eu), k
            assert np.allclose(got_f[at_f:at_f + f.size].reshape(f.shape), f, rtol=1e-12, atol=0), k
            assert np.allclose(f, ef, rtol=1e-9, atol=0), k
            at_u += u.size
            at_f += f.size
        assert tuple(int(v) for v in got_totals) == chain.totals() == oracle.totals()

    compare(single, checker, ranks[0]["u1"], ranks[0]["f1"], ranks[0]["totals1"])
    estimates = [checker.estimate_priors(k) for k, info in enumerate(single.info) if info.algorithm == 0]
    assert np.allclose(ranks[0]["noise"], [e[0] for e in estimates], rtol=1e-9, atol=0)
    assert np.allclose(ranks[0]["concentration"], np.concatenate([e[1] for e in estimates]), rtol=1e-9, atol=0)

    # pass 2 with the estimated priors (docs/pamld.md:38-44, classifier.h:125-160)
    adjusted = O.compile_job(spec["job"])
    for k, (topic, decoder) in enumerate(O.decoder_chain(adjusted)):
        if decoder["algorithm"] != "pamld":
            continue
        noise, concentration = checker.estimate_priors(k)
        decoder["noise"] = noise
        decoder["undetermined"]["concentration"] = noise
        for record in decoder["codec"].values():
            record["concentration"] = float(concentration[record["index"] - 1])
    single.adjust_priors()
    single.reset()
    single.decode(single.pack(code, quality, offset), N_READS)
    second = O.best_oracle(adjusted, len(code))
    second.decode(O.ReadBatch(code, quality, offset))
    compare(single, second, ranks[0]["u2"], ranks[0]["f2"], ranks[0]["totals2"])
