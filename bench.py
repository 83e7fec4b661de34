#!/usr/bin/env python3
"""bench.py — reads/s decoded on the barcode classification path, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1] [--reads R] [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic reads resident in HBM:
reset accumulators -> classify every read of the batch against every barcode of every decoder of
the workload (one kernel per decoder) -> all-reduce the accumulators across ranks (the path's only
collective). Reads are sharded per rank (weak scaling: --reads is per GPU). The default workload
is BASELINE.json's PAMLD headline configuration, c1 (96 x [8,8] dual index).

Prints ONE JSON line (rank 0). `value` is device-timed (CUDA events on the launching stream, max
over ranks); `e2e` is the same metric through the host-buffer C-ABI call phq_decode_batch (pinned
host tiles in, results + qcfail out, copies inside the timed region); `roofline` relates the
dominant kernel to the measured HBM peak; `cpu_baseline` is the reference's own decoder classes
(oracle/_ref, else the C port) timed on this host's cores on a bounded sample of the same reads.

`--impl reference` times only that CPU implementation (all host threads), same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reads/sec decoded (PAMLD)"
UNIT = "reads/s"
DEFAULT_READS = {"c1": 1 << 28, "c2": 1 << 28, "c3": 1 << 26, "c4": 1 << 26, "c5": 148 * 15 * 32 * 8}
WORKLOAD_LABEL = {
    "c1": "C1: Illumina dual-index (i7+i5, 8 bp each) 96-sample PAMLD, noise 0.05, confidence threshold 0.95",
    "c2": "C2: same 96-sample dual-index set, MDD, distance tolerance [1,1]",
    "c3": "C3: SPLiT-seq 3 x 96 x [8] + 4 x [6] PAMLD cellular + naive 10 bp UMI",
    "c4": "C4: sci-RNA-seq 96 x [10] + 196 x [10,10] PAMLD cellular + naive 8 bp UMI",
    "c5": "C5: 16 bp cellular PAMLD against a 737,280 barcode whitelist + naive 12 bp UMI",
}


def algorithmic_bytes_per_read(chain, compiled) -> int:
    """per tiled decoder ceil(L/4) + ceil(L/8) bytes of bases and no-call mask, L quality bytes and
    16 bytes out; + 1 byte qcfail in and out per read. An MDD decoder without quality masking never needs the quality
    plane (absent positions travel in the base words), so its L quality bytes are not counted."""
    total = 2
    for info in chain.info:
        if info.has_tile:
            topic = compiled[{0: "sample", 1: "molecular", 2: "cellular"}[info.topic]]
            spec = topic[info.index] if isinstance(topic, list) else topic
            L = info.nucleotide_cardinality
            total += (L + 3) // 4 + (L + 7) // 8 + 16
            if not (info.algorithm == 1 and int(spec.get("quality masking threshold", 0)) <= 0):
                total += L
    return total


def pair_words_per_read(chain) -> int:
    return sum(info.barcode_cardinality for info in chain.info if info.has_tile)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one
    `nvidia-smi -lms 100` process streaming CSV lines while the timed steps run."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.process = None
        self.thread = None
        self.active = threading.Event()

    def _run(self):
        for line in self.process.stdout:
            if self.active.is_set():
                self.samples.append([x.strip() for x in line.strip().split(",")])

    def __enter__(self):
        try:
            self.process = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
            time.sleep(0.35)            # let the first samples arrive before the timed region starts
        except Exception:
            self.process = None
        return self

    def begin(self):
        self.active.set()

    def end(self):
        self.active.clear()

    def __exit__(self, *a):
        if self.process is not None:
            self.process.terminate()
            try:
                self.process.wait(timeout=5)
            except Exception:
                self.process.kill()

    def summary(self):
        rows = [s for s in self.samples if len(s) >= 7 and s[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in rows)]
        power = [float(s[2]) for s in rows if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(power) if power else None}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload_name, n_reads):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full`
    capture (profiles/traffic.json: bytes per read), scaled to this launch; None when no capture exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        per_read = json.load(open(path))[workload_name]["dram_bytes_per_read"]
        return per_read * n_reads
    except Exception:
        return None


def host_sample(compiled, spec, n_reads, seed):
    from pheniqs_b200 import workload
    return workload.synthesize(compiled, spec["input segment length"], n_reads, seed=seed, sampling="zipf" if spec["name"] == "c4" else "prior")


def time_cpu(compiled, spec, seconds_target=12.0, threads=None, seed=99):
    """The reference's CPU implementation of the path on this host's cores, on a bounded sample."""
    from oracle import oracle as O
    threads = threads or (os.cpu_count() or 1)
    probe_n = 2000 if spec["name"] != "c5" else 16
    code, quality, offset, _ = host_sample(compiled, spec, probe_n, seed)
    checker = O.best_oracle(compiled, len(code))
    probe = checker.decode(O.ReadBatch(code, quality, offset), threads=1, want_outputs=False)
    rate = probe_n / max(probe.seconds, 1e-9)
    n = int(min(max(rate * threads * seconds_target, threads * 4), 4e7))
    code, quality, offset, _ = host_sample(compiled, spec, n, seed + 1)
    checker = O.best_oracle(compiled, len(code))
    out = checker.decode(O.ReadBatch(code, quality, offset), threads=threads, want_outputs=False)
    return {"value": n / out.seconds, "unit": UNIT, "cores": threads, "kind": checker.kind,
            "sample": "%d synthetic reads of the same workload, %d threads each with private decoders (transcode.cpp:2296), %.1f s" % (n, threads, out.seconds)}, n, out.seconds


def run_reference(args, rank, world):
    """The reference's own decoder classes on all host threads: one bounded sample of the workload (synthesized once),
    decoded args.warmup + args.steps times with fresh decoder sets; sized so the whole run stays within a few minutes."""
    from oracle import oracle as O
    from pheniqs_b200 import compile_job, workload
    if rank != 0:
        return
    spec = workload.load(args.workload)
    compiled = compile_job(spec["job"])
    threads = os.cpu_count() or 1
    passes = max(args.steps + args.warmup, 1)
    per_step = max(1.5, min(15.0, 100.0 / passes))
    probe_n = 2000 if spec["name"] != "c5" else 16
    code, quality, offset, _ = host_sample(compiled, spec, probe_n, 999)
    probe = O.best_oracle(compiled, len(code)).decode(O.ReadBatch(code, quality, offset), threads=1, want_outputs=False)
    rate = probe_n / max(probe.seconds, 1e-9)
    n = int(min(max(rate * threads * per_step, threads * 4), 4e7))
    code, quality, offset, _ = host_sample(compiled, spec, n, 1000)
    batch = O.ReadBatch(code, quality, offset)
    total_seconds, kind = 0.0, None
    for step in range(passes):
        checker = O.best_oracle(compiled, len(code))
        kind = checker.kind
        out = checker.decode(batch, threads=threads, want_outputs=False)
        if step >= args.warmup:
            total_seconds += out.seconds
    steps = max(passes - args.warmup, 1)
    value = n * steps / max(total_seconds, 1e-9)
    baseline = {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                "sample": "%d synthetic reads of the same workload per step, %d threads each with private decoders (transcode.cpp:2296), %.1f s per step" % (n, threads, total_seconds / steps)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_seconds / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD_LABEL[args.workload], "reads_per_step": n},
            "cpu_baseline": baseline, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--workload", default="c1", choices=sorted(DEFAULT_READS))
    parser.add_argument("--reads", type=int, default=0, help="reads per GPU per step (default: per workload)")
    parser.add_argument("--e2e-reads", type=int, default=0, help="reads per GPU per end-to-end step (default min(reads, 2^26))")
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-e2e", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pheniqs_b200 import DecoderChain, compile_job, workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the classification path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    spec = workload.load(args.workload)
    compiled = compile_job(spec["job"])
    chain = DecoderChain(compiled, device=local_rank)
    n = args.reads or DEFAULT_READS[args.workload]
    sampling = "zipf" if args.workload == "c4" else "prior"
    tiles = workload.synthesize_device_tiles(chain, compiled, n, device, seed=workload.SEED + 17 * rank, sampling=sampling)
    flags = torch.zeros(n, dtype=torch.uint8, device=device)
    results = [torch.empty((n, 2), dtype=torch.float64, device=device) if info.has_tile else None for info in chain.info]
    stream = torch.cuda.current_stream(device)
    u64_plane, f64_plane = chain.accumulator_tensors()

    # timing rule: inputs larger than L2, or L2 flushed between steps. The tile planes of a step exceed the 126 MB L2 for
    # c1-c4 at their default sizes; where they do not (c5: a few MB of reads against a table that is MEANT to live in L2),
    # a 256 MB write between steps evicts them (about 40 us inside a step of tens of ms).
    input_bytes = algorithmic_bytes_per_read(chain, compiled) * n
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if input_bytes < (192 << 20) else None

    def step():
        if flush is not None:
            flush.zero_()
        flags.zero_()
        u64_plane.zero_()
        f64_plane.zero_()
        chain.decode_device(tiles, n, flags, results, stream)
        if world > 1:
            dist.all_reduce(u64_plane)
            dist.all_reduce(f64_plane)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        step()
    barrier()
    launches_before = chain.statistics()["kernel_launches"]
    kernel_ms = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.begin()
        torch.cuda.nvtx.range_push("timed")
        start.record(stream)
        for _ in range(args.steps):
            step()
        stop.record(stream)
        barrier()
        torch.cuda.nvtx.range_pop()
        launches = chain.statistics()["kernel_launches"] - launches_before
        # keep the GPU under the same load a little longer when the timed region was too short to sample
        extra = 0
        while len(clocks.samples) < 3 and extra < 200:
            step()
            extra += 1
            if extra % 4 == 0:
                torch.cuda.synchronize(device)
        torch.cuda.synchronize(device)
        clocks.end()
    elapsed_ms = start.elapsed_time(stop)
    # the dominant kernel alone, timed live with CUDA events on the launching stream (phq_last_kernel_milliseconds)
    for _ in range(min(args.steps, 5)):
        flags.zero_()
        u64_plane.zero_()
        f64_plane.zero_()
        chain.decode_device(tiles, n, flags, results, stream)
        kernel_ms.append(chain.last_kernel_milliseconds())
    kernel_ms_mean = float(np.mean(kernel_ms))
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = n * world * args.steps / (elapsed_ms * 1e-3)

    # sanity on the timed work: every read was classified
    if world == 1:
        u, _ = chain.accumulators(next(k for k, info in enumerate(chain.info) if info.has_tile))
        assert int(u[:, 0].sum()) == n, "accumulators do not cover the batch"

    bytes_per_read = algorithmic_bytes_per_read(chain, compiled)
    peak, peak_source = measured_peaks()
    achieved = bytes_per_read * n / (kernel_ms_mean * 1e-3) / 1e9
    pairs = pair_words_per_read(chain)
    clock_summary = clocks.summary()
    sm_mhz = clock_summary.get("sm_mhz") or 0
    sm_count = torch.cuda.get_device_properties(device).multi_processor_count
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(args.workload, n),
                "peak_source": peak_source, "kernel": "; ".join(chain.kernel_description(k) for k in range(chain.n_decoders)),
                "kernel_ms_per_launch_set": kernel_ms_mean, "algorithmic_bytes_per_read": bytes_per_read,
                "pair_words_per_read": pairs, "pair_words_per_s": pairs * n / (kernel_ms_mean * 1e-3),
                "pair_words_per_clk_per_sm": (pairs * n / (kernel_ms_mean * 1e-3)) / (sm_mhz * 1e6 * sm_count) if sm_mhz else None,
                "note": "the path is issue/shared-memory bound, not HBM bound (SURVEY.md §8d); the HBM fraction is reported as the contract asks"}

    # SURVEY.md §8d also asks for the integer / issue roofline: the exhaustive formulation costs 19 integer-pipe operations
    # per (read, barcode) pair-word for PAMLD and 7 for MDD; the chip issues 4 warp instructions per clock and SM (2 on the
    # ALU pipe + 2 on the FMA pipe, measured: profiles/r01_microbench.txt). The kernels restructure the arithmetic (separable
    # grids, lookups, pruned bit-sliced scans), so the EQUIVALENT rate can exceed the peak: that ratio is the algorithmic gain.
    equivalent_ops = sum((19 if info.algorithm == 0 else 7) * info.barcode_cardinality for info in chain.info if info.has_tile)
    if sm_mhz:
        issue_peak = sm_count * 4 * 32 * sm_mhz * 1e6
        roofline["int"] = {"equivalent_ops_per_read": equivalent_ops, "equivalent_ops_per_s": equivalent_ops * n / (kernel_ms_mean * 1e-3),
                           "issue_peak_lane_ops_per_s": issue_peak, "frac": equivalent_ops * n / (kernel_ms_mean * 1e-3) / issue_peak,
                           "note": "operations of the exhaustive formulation per second over the measured issue peak; above 1 = work the kernels avoid"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_LABEL[args.workload], "reads_per_gpu": n, "decoders": chain.n_decoders, "barcodes": [info.barcode_cardinality for info in chain.info],
                       "l2": ("inputs (%d MB per GPU) exceed L2; no flush needed" % (bytes_per_read * n >> 20)) if flush is None else
                             ("inputs are %d MB per GPU: L2 flushed with a 256 MB write between steps (inside the timed region)" % (bytes_per_read * n >> 20)), "parallelism": "reads sharded x%d, accumulators all-reduced" % world},
            "roofline": roofline, "gpu_launches": int(launches), "clocks": clock_summary}

    # ---------------------------------------------------------------- end to end through the host-buffer C-ABI calls
    # Pinned host tiles in, per-read results out, every copy inside the timed region. Two forms of the same call:
    #   e2e       phq_decode_batch_compact with the smallest quality form that fits the batch (here 2-bit codebook
    #             indices: the synthetic reads, like current Illumina output, have 4 distinct qualities) and the 8-byte
    #             records that carry what the reference's output carries (index, distance, qcfail, float(1 - confidence));
    #             used when every topic has one decoder, where those records are lossless w.r.t. the reference's output
    #   e2e_full  phq_decode_batch with Phred bytes in and 16-byte {index, distance, f64 confidence} + qcfail byte out
    if not args.no_e2e:
        from pheniqs_b200 import COMPACT_DTYPE, RESULT_DTYPE
        # per rank; smaller with many ranks on one host (pinned staging is ~170 B per read and rank, and the ranks share the host's memory system)
        m = args.e2e_reads or min(n, (1 << 26) // max(1, world // 2))
        host_tiles = chain.allocate_tiles(m, pinned=True)
        for k, t in enumerate(tiles):
            if t is None:
                continue
            host_tiles[k].bases[:] = t[0][:, :m].cpu().numpy().view(np.uint32)
            host_tiles[k].nmask[:] = t[1][:, :m].cpu().numpy().view(np.uint16)
            host_tiles[k].quality[:] = t[2][:, :m].cpu().numpy().view(np.uint32)
        e2e_steps = max(3, min(args.steps, 5))

        def time_host(call):
            for _ in range(2):
                call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                call()
            torch.cuda.synchronize(device)
            seconds = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([seconds], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                seconds = float(t.item())
            return m * world * e2e_steps / seconds

        keep = []
        full_results = []
        for info in chain.info:
            if info.has_tile:
                buffer = torch.zeros((m, 2), dtype=torch.float64).pin_memory()
                keep.append(buffer)
                full_results.append(buffer.numpy().view(RESULT_DTYPE).reshape(-1))
            else:
                full_results.append(None)
        qc_buffer = torch.zeros(m, dtype=torch.uint8).pin_memory()
        qc_out = qc_buffer.numpy()
        h2d_full = sum(t.bytes_per_read() * m for t in host_tiles if t is not None)
        d2h_full = sum(16 * m for info in chain.info if info.has_tile) + m
        full_value = time_host(lambda: chain.decode(host_tiles, m, None, results=full_results, qcfail_out=qc_out))
        e2e_full = {"value": full_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_full), "d2h_bytes_per_step": int(d2h_full),
                    "reads_per_gpu_per_step": m, "steps": e2e_steps, "call": "phq_decode_batch (Phred byte tiles in; 16-byte results + qcfail byte out)"}

        topics = [info.topic for info in chain.info if info.has_tile]
        lossless = len(topics) == len(set(topics))
        if lossless:
            forms = [t.compress_quality() for t in host_tiles if t is not None]
            compact_results = []
            for info in chain.info:
                if info.has_tile:
                    buffer = torch.zeros(m, dtype=torch.float64).pin_memory()
                    keep.append(buffer)
                    compact_results.append(buffer.numpy().view(COMPACT_DTYPE).reshape(-1))
                else:
                    compact_results.append(None)
            h2d = sum(t.bytes_per_read() * m for t in host_tiles if t is not None)
            d2h = sum(8 * m for info in chain.info if info.has_tile)
            value = time_host(lambda: chain.decode_compact(host_tiles, m, None, results=compact_results))
            line["e2e"] = {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                           "reads_per_gpu_per_step": m, "steps": e2e_steps, "quality_bits": forms,
                           "call": "phq_decode_batch_compact (2-bit + mask + codebook-index tiles in; 8-byte records out: index, distance, qcfail, float(1 - confidence))"}
            line["e2e_full"] = e2e_full
        else:
            line["e2e"] = e2e_full
        # e2e_raw: the bytes of the FASTQ records in (ASCII nucleotides and qualities of the barcode-bearing segments,
        # pinned), packing on the device (phq_decode_batch_raw_compact): the host does no per-read work at all
        raw = workload.raw_segments_from_device_tiles(chain, compiled, tiles, m) if lossless else None
        if raw is not None:
            segments = [None if g is None else g[:4] for g in raw]
            raw_results = []
            for info in chain.info:
                if info.has_tile:
                    buffer = torch.zeros(m, dtype=torch.float64).pin_memory()
                    keep.append(buffer)
                    raw_results.append(buffer.numpy().view(COMPACT_DTYPE).reshape(-1))
                else:
                    raw_results.append(None)
            raw_value = time_host(lambda: chain.decode_raw(segments, m, 33, None, compact=True, results=raw_results))
            assert all(a is None or np.array_equal(a["packed"], b["packed"]) for a, b in zip(raw_results, compact_results))
            line["e2e_raw"] = {"value": raw_value, "unit": UNIT, "h2d_bytes_per_step": int(sum(2 * g[3] * m for g in raw if g is not None)), "d2h_bytes_per_step": int(d2h),
                               "reads_per_gpu_per_step": m, "steps": e2e_steps,
                               "call": "phq_decode_batch_raw_compact (FASTQ bytes of the barcode segments in, packed on the device; 8-byte records out)"}
            # e2e_tags: the same bytes in, the BAM auxiliary block of every read out (RG BC QT XB ... as Read::flush and
            # Auxiliary::encode write them) + the qcfail byte: phq_decode_batch_raw_tags
            stride = chain.tag_record_bytes()
            aux_buffer = torch.zeros((m, stride), dtype=torch.uint8).pin_memory()
            length_buffer = torch.zeros(m, dtype=torch.int32).pin_memory()
            flag_buffer = torch.zeros(m, dtype=torch.uint8).pin_memory()
            keep += [aux_buffer, length_buffer, flag_buffer]
            tags_value = time_host(lambda: chain.decode_raw_tags(segments, m, 33, None, stride=stride, aux=aux_buffer.numpy(), aux_length=length_buffer.numpy(), qcfail_out=flag_buffer.numpy()))
            line["e2e_tags"] = {"value": tags_value, "unit": UNIT, "h2d_bytes_per_step": int(sum(2 * g[3] * m for g in raw if g is not None)), "d2h_bytes_per_step": int(m * (stride + 5)),
                                "reads_per_gpu_per_step": m, "steps": e2e_steps, "record_bytes": stride,
                                "call": "phq_decode_batch_raw_tags (FASTQ bytes of the barcode segments in; the auxiliary block RG BC QT XB ... of every read + its length + qcfail out)"}
        del host_tiles, full_results, keep, raw

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        baseline, _, _ = time_cpu(compiled, spec)
        line["cpu_baseline"] = baseline
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
