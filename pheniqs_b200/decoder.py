"""decoder.py — host-side mirror of the reference's decoder interface over the C ABI.

`DecoderChain` stands where one thread's `TranscodingDecoder` stands in the reference
(transcode.h:40-65): it is built from the compiled decoder ontology, classifies reads
(a batch per call instead of one read per call), owns per-barcode accumulators, is merged
with its peers by `collect` (here: one all-reduce over NVLink instead of a serial loop over
threads, transcode.cpp:162-179) and estimates priors in `finalize` (classifier.h:94-124).

All arithmetic happens inside libpheniqs_b200.so; this file only marshals buffers.
PyTorch is used for device memory, streams and torch.distributed, nothing else.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import binding
from .binding import DecoderInfo, RawSegment, Tile, check, library

RESULT_DTYPE = np.dtype([("index", np.int32), ("distance", np.int32), ("confidence", np.float64)])
COMPACT_DTYPE = np.dtype([("packed", np.uint32), ("error_probability", np.float32)])
ACC_U64_COLUMNS = ("count", "pf count", "accumulated distance", "low conditional confidence count", "low confidence count", "accumulated pf distance")
ACC_F64_COLUMNS = ("accumulated confidence", "accumulated pf confidence")


def compile_job(job) -> dict:
    """Transcode::compile for the decoder sections (phq_compile_job). Host only."""
    lib = library()
    text = job if isinstance(job, (str, bytes)) else json.dumps(job)
    if isinstance(text, str):
        text = text.encode()
    out = C.c_void_p()
    check(lib.phq_compile_job(text, C.byref(out)))
    try:
        return json.loads(C.string_at(out).decode())
    finally:
        lib.phq_free(out)


def load_job(path: str) -> dict:
    """The job file at `path` with its imports merged in (phq_load_job; Job::load_instruction_with_import). Host only."""
    lib = library()
    out = C.c_void_p()
    check(lib.phq_load_job(str(path).encode(), C.byref(out)))
    try:
        return json.loads(C.string_at(out).decode())
    finally:
        lib.phq_free(out)


def adjust_job(job, report, precision: int = 15, text: bool = False):
    """The prior adjusted job (phq_adjust_job): tool/pheniqs-prior-api.py:39-56 / classifier.h:125-160. Host only."""
    lib = library()
    as_bytes = lambda v: (v if isinstance(v, (str, bytes)) else json.dumps(v))
    a, b = as_bytes(job), as_bytes(report)
    out = C.c_void_p()
    check(lib.phq_adjust_job(a.encode() if isinstance(a, str) else a, b.encode() if isinstance(b, str) else b, precision, C.byref(out)))
    try:
        raw = C.string_at(out).decode()
        return raw if text else json.loads(raw)
    finally:
        lib.phq_free(out)


def parse_auxiliary(record) -> dict:
    """The tags of one BAM auxiliary block (bytes): Z tags as str, f tags as numpy float32; order kept."""
    data = bytes(record)
    out, at = {}, 0
    while at + 3 <= len(data):
        tag, kind = data[at:at + 2].decode("latin-1"), chr(data[at + 2])
        at += 3
        if kind == "Z":
            end = data.index(b"\0", at)
            out[tag] = data[at:end].decode("latin-1")
            at = end + 1
        elif kind == "f":
            out[tag] = np.frombuffer(data[at:at + 4], dtype="<f4")[0]
            at += 4
        else:
            raise ValueError("unexpected auxiliary type " + kind)
    return out


def shard_range(n_reads: int, rank: int, world_size: int):
    """Contiguous read range of one rank, as the reference slices nothing but threads pull in turn
    (transcode.cpp:287-316); reads are independent, any partition is valid."""
    return n_reads * rank // world_size, n_reads * (rank + 1) // world_size


class HostTile:
    """Host planes of one decoder for n_reads reads (numpy arrays, optionally pinned)."""

    def __init__(self, info: DecoderInfo, n_reads: int, pinned: bool = False):
        self.n_reads = n_reads
        self.pitch = max(n_reads, 1)
        self.words = info.word_cardinality
        self.quality_words = info.quality_word_cardinality
        self.nucleotides = info.nucleotide_cardinality
        self.quality_bits = 8
        self.quality_codebook = bytes(16)
        self._pinned = []
        self.bases = self._allocate((self.words, self.pitch), np.uint32, pinned)
        self.nmask = self._allocate((self.words, self.pitch), np.uint16, pinned)
        self.quality = self._allocate((self.quality_words, self.pitch), np.uint32, pinned)

    def _allocate(self, shape, dtype, pinned):
        if not pinned:
            return np.zeros(shape, dtype=dtype)
        import torch
        t = torch.zeros(shape, dtype={np.uint32: torch.int32, np.uint16: torch.int16}[dtype]).pin_memory()
        self._pinned.append(t)
        return t.numpy().view(dtype)

    def as_struct(self) -> Tile:
        return Tile(self.bases.ctypes.data, self.nmask.ctypes.data, self.quality.ctypes.data, self.pitch,
                    self.quality_bits, (C.c_uint8 * 16)(*self.quality_codebook))

    @property
    def packed_quality_words(self) -> int:
        """rows of the quality plane the chosen form occupies"""
        return (self.nucleotides * self.quality_bits + 31) // 32

    def bytes_per_read(self) -> int:
        return self.words * 6 + self.packed_quality_words * 4

    def compress_quality(self) -> int:
        """Re-encode a byte-form quality plane in place as indices into a codebook of its distinct values
        (what phq_pack does with quality_bits = -1), when at most 16 distinct values occur. Returns the form."""
        if self.quality_bits != 8:
            return self.quality_bits
        L = self.nucleotides
        present = np.zeros(256, dtype=bool)
        for w in range(self.quality_words):
            for k in range(4):
                if 4 * w + k < L:
                    present[np.unique((self.quality[w] >> (8 * k)) & 0xff)] = True
        values = np.nonzero(present)[0]
        bits = 2 if values.size <= 4 else (4 if values.size <= 16 else 8)
        if bits == 8:
            return 8
        index_of = np.zeros(256, dtype=np.uint32)
        index_of[values] = np.arange(values.size, dtype=np.uint32)
        per_word = 32 // bits
        packed = np.zeros(((L * bits + 31) // 32, self.pitch), dtype=np.uint32)
        for j in range(L):
            q = (self.quality[j // 4] >> (8 * (j % 4))) & 0xff
            packed[j // per_word] |= index_of[q] << np.uint32(bits * (j % per_word))
        self.quality[:packed.shape[0]] = packed
        self.quality_bits = bits
        self.quality_codebook = bytes(values.astype(np.uint8).tolist() + [0] * (16 - values.size))
        return bits


class DecoderChain:
    def __init__(self, compiled_job, device: int = 0):
        self.lib = library()
        text = compiled_job if isinstance(compiled_job, (str, bytes)) else json.dumps(compiled_job)
        if isinstance(text, str):
            text = text.encode()
        self.handle = C.c_void_p()
        self.device = device
        check(self.lib.phq_create(text, device, C.byref(self.handle)))
        self.n_decoders = self.lib.phq_decoder_count(self.handle)
        self.info = []
        for k in range(self.n_decoders):
            info = DecoderInfo()
            check(self.lib.phq_decoder_describe(self.handle, k, C.byref(info)), self.handle)
            self.info.append(info)

    @classmethod
    def from_job(cls, job, device: int = 0) -> "DecoderChain":
        return cls(compile_job(job), device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.phq_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ feed seam
    def allocate_tiles(self, n_reads: int, pinned: bool = False):
        return [HostTile(info, n_reads, pinned) if info.has_tile else None for info in self.info]

    def _tile_array(self, tiles):
        array = (Tile * self.n_decoders)()
        for k, t in enumerate(tiles):
            if t is not None:
                array[k] = t if isinstance(t, Tile) else t.as_struct()
        return array

    def pack(self, code, quality, offset, tiles=None, pinned: bool = False, quality_bits: int = 8):
        """Rule::apply + packing for a batch held as per-segment flat arrays (phq_pack).
        quality_bits: 8 = Phred bytes, 4 / 2 = codebook indices, -1 = the smallest form that fits."""
        n_segments = len(code)
        n_reads = int(offset[0].shape[0] - 1)
        if tiles is None:
            tiles = self.allocate_tiles(n_reads, pinned)
        code = [np.ascontiguousarray(c, dtype=np.uint8) for c in code]
        quality = [np.ascontiguousarray(q, dtype=np.uint8) for q in quality]
        offset = [np.ascontiguousarray(o, dtype=np.int64) for o in offset]
        pc = (C.c_void_p * n_segments)(*[c.ctypes.data for c in code])
        pq = (C.c_void_p * n_segments)(*[q.ctypes.data for q in quality])
        po = (C.c_void_p * n_segments)(*[o.ctypes.data for o in offset])
        for t in tiles:
            if t is not None:
                t.quality_bits = quality_bits
        array = self._tile_array(tiles)
        check(self.lib.phq_pack(self.handle, n_reads, n_segments, pc, pq, po, array), self.handle)
        for k, t in enumerate(tiles):
            if t is not None:
                t.quality_bits = int(array[k].quality_bits)
                t.quality_codebook = bytes(array[k].quality_codebook)
        return tiles

    # ------------------------------------------------------------------ classification
    def decode(self, tiles, n_reads: int, qcfail_in=None, want_results: bool = True, results=None, qcfail_out=None):
        """TranscodingDecoder::classify over host buffers (phq_decode_batch)."""
        if results is None:
            results = [np.zeros(n_reads, dtype=RESULT_DTYPE) if want_results else None for _ in range(self.n_decoders)]
        if qcfail_out is None:
            qcfail_out = np.zeros(n_reads, dtype=np.uint8)
        pointers = (C.c_void_p * self.n_decoders)(*[None if r is None else r.ctypes.data for r in results])
        qin = None if qcfail_in is None else np.ascontiguousarray(qcfail_in, dtype=np.uint8)
        check(self.lib.phq_decode_batch(self.handle, n_reads, self._tile_array(tiles), None if qin is None else qin.ctypes.data,
                                        pointers, qcfail_out.ctypes.data), self.handle)
        return results, qcfail_out

    def decode_compact(self, tiles, n_reads: int, qcfail_in=None, results=None):
        """phq_decode_batch_compact: 8-byte records (index | distance | qcfail, float error probability)."""
        if results is None:
            results = [np.zeros(n_reads, dtype=COMPACT_DTYPE) if info.has_tile else None for info in self.info]
        pointers = (C.c_void_p * self.n_decoders)(*[None if r is None else r.ctypes.data for r in results])
        qin = None if qcfail_in is None else np.ascontiguousarray(qcfail_in, dtype=np.uint8)
        check(self.lib.phq_decode_batch_compact(self.handle, n_reads, self._tile_array(tiles), None if qin is None else qin.ctypes.data, pointers), self.handle)
        return results

    def _raw_segment_array(self, segments):
        array = (RawSegment * max(len(segments), 1))()
        keep = []
        for i, g in enumerate(segments):
            if g is None:
                continue
            sequence, quality, offset, length = g[:4]
            sequence = np.ascontiguousarray(sequence, dtype=np.uint8)
            quality = np.ascontiguousarray(quality, dtype=np.uint8)
            offset = None if offset is None else np.ascontiguousarray(offset, dtype=np.int64)
            keep.append((sequence, quality, offset))
            array[i] = RawSegment(sequence.ctypes.data, quality.ctypes.data, None if offset is None else offset.ctypes.data, int(length))
        return array, keep

    def decode_raw(self, segments, n_reads: int, phred_offset: int = 33, qcfail_in=None, compact: bool = False, results=None, qcfail_out=None, bam: bool = False):
        """phq_decode_batch_raw[_compact]: the bytes of the FASTQ records in, packing on the device; with bam=True
        phq_decode_batch_bam[_compact]: the reference's own decoded form in (one BAM code and one Phred byte per base).

        segments[i] = (sequence uint8 [bytes], quality uint8 [bytes], offset int64 [n_reads + 1] or None, length) or None
        for an input segment no token refers to."""
        array, keep = self._raw_segment_array(segments)
        if results is None:
            dtype = COMPACT_DTYPE if compact else RESULT_DTYPE
            results = [np.zeros(n_reads, dtype=dtype) if (info.has_tile or not compact) else None for info in self.info]
        pointers = (C.c_void_p * self.n_decoders)(*[None if r is None else r.ctypes.data for r in results])
        qin = None if qcfail_in is None else np.ascontiguousarray(qcfail_in, dtype=np.uint8)
        qin_pointer = None if qin is None else qin.ctypes.data
        if compact:
            if bam:
                check(self.lib.phq_decode_batch_bam_compact(self.handle, n_reads, len(segments), array, qin_pointer, pointers), self.handle)
            else:
                check(self.lib.phq_decode_batch_raw_compact(self.handle, n_reads, len(segments), array, phred_offset, qin_pointer, pointers), self.handle)
            return results
        if qcfail_out is None:
            qcfail_out = np.zeros(n_reads, dtype=np.uint8)
        if bam:
            check(self.lib.phq_decode_batch_bam(self.handle, n_reads, len(segments), array, qin_pointer, pointers, qcfail_out.ctypes.data), self.handle)
        else:
            check(self.lib.phq_decode_batch_raw(self.handle, n_reads, len(segments), array, phred_offset, qin_pointer, pointers, qcfail_out.ctypes.data), self.handle)
        return results, qcfail_out

    def tag_record_bytes(self) -> int:
        """The smallest auxiliary record stride the job needs (phq_tag_record_bytes)."""
        value = C.c_int32()
        check(self.lib.phq_tag_record_bytes(self.handle, C.byref(value)), self.handle)
        return value.value

    def decode_raw_tags(self, segments, n_reads: int, phred_offset: int = 33, qcfail_in=None, stride: int = 0, want_results: bool = False,
                        aux=None, aux_length=None, qcfail_out=None, bam: bool = False):
        """phq_decode_batch_raw_tags (bam=True: phq_decode_batch_bam_tags): feed bytes in, the BAM auxiliary block of every
        read out.
        Returns (aux uint8 [n_reads, stride], aux_length int32 [n_reads], qcfail uint8 [n_reads][, results])."""
        array, keep = self._raw_segment_array(segments)
        stride = stride or self.tag_record_bytes()
        aux = np.zeros((max(n_reads, 1), stride), dtype=np.uint8) if aux is None else aux
        aux_length = np.zeros(max(n_reads, 1), dtype=np.int32) if aux_length is None else aux_length
        qcfail_out = np.zeros(max(n_reads, 1), dtype=np.uint8) if qcfail_out is None else qcfail_out
        results = [np.zeros(n_reads, dtype=RESULT_DTYPE) if (want_results and info.has_tile) else None for info in self.info]
        pointers = (C.c_void_p * self.n_decoders)(*[None if r is None else r.ctypes.data for r in results])
        qin = None if qcfail_in is None else np.ascontiguousarray(qcfail_in, dtype=np.uint8)
        qin_pointer = None if qin is None else qin.ctypes.data
        if bam:
            check(self.lib.phq_decode_batch_bam_tags(self.handle, n_reads, len(segments), array, qin_pointer,
                                                     aux.ctypes.data, stride, aux_length.ctypes.data, qcfail_out.ctypes.data, pointers), self.handle)
        else:
            check(self.lib.phq_decode_batch_raw_tags(self.handle, n_reads, len(segments), array, phred_offset, qin_pointer,
                                                     aux.ctypes.data, stride, aux_length.ctypes.data, qcfail_out.ctypes.data, pointers), self.handle)
        out = (aux[:n_reads], aux_length[:n_reads], qcfail_out[:n_reads])
        return out + (results,) if want_results else out

    def decode_device(self, device_tiles, n_reads: int, qcfail, results=None, stream=None):
        """Same over device-resident torch tensors; asynchronous on `stream` (phq_decode_batch_device).

        device_tiles[k] = (bases int32 [words, pitch], nmask int16 [words, pitch], quality int32 [qwords, pitch]) or None."""
        array = (Tile * self.n_decoders)()
        for k, t in enumerate(device_tiles):
            if t is not None:
                bases, nmask, quality = t[:3]
                bits, codebook = (t[3], t[4]) if len(t) > 3 else (8, bytes(16))
                array[k] = Tile(bases.data_ptr(), nmask.data_ptr(), quality.data_ptr(), bases.shape[-1], bits, (C.c_uint8 * 16)(*codebook))
        pointers = (C.c_void_p * self.n_decoders)(*[None if (results is None or r is None) else r.data_ptr() for r in (results or [None] * self.n_decoders)])
        handle = 0 if stream is None else stream.cuda_stream
        check(self.lib.phq_decode_batch_device(self.handle, n_reads, array, qcfail.data_ptr(), pointers, C.c_void_p(handle)), self.handle)

    def upload(self, tiles, n_reads: int):
        """Host tiles -> device tensors (torch), for keeping a batch resident in HBM."""
        import torch
        device = torch.device("cuda", self.device)
        out = []
        for t in tiles:
            if t is None:
                out.append(None)
                continue
            out.append((torch.from_numpy(t.bases.view(np.int32)).to(device), torch.from_numpy(t.nmask.view(np.int16)).to(device),
                        torch.from_numpy(t.quality.view(np.int32)).to(device), t.quality_bits, t.quality_codebook))
        return out

    def kernel_description(self, k: int) -> str:
        """Names of the kernels decoder k launches with its current tables."""
        buffer = C.create_string_buffer(256)
        check(self.lib.phq_kernel_description(self.handle, k, buffer, len(buffer)), self.handle)
        return buffer.value.decode()

    def last_kernel_milliseconds(self) -> float:
        ms = C.c_float()
        check(self.lib.phq_last_kernel_milliseconds(self.handle, C.byref(ms)), self.handle)
        return ms.value

    # ------------------------------------------------------------------ accumulators / priors
    def accumulators(self, k: int):
        rows = self.info[k].barcode_cardinality + 1
        u = np.zeros((rows, 6), dtype=np.uint64)
        f = np.zeros((rows, 2), dtype=np.float64)
        check(self.lib.phq_accumulators(self.handle, k, u.ctypes.data, f.ctypes.data), self.handle)
        return u, f

    def totals(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self.lib.phq_totals(self.handle, C.byref(a), C.byref(b)), self.handle)
        return a.value, b.value

    def reset(self, stream=None):
        """Clear the accumulators (and the collected state): synchronously, or in `stream` order when one is given."""
        if stream is None:
            check(self.lib.phq_reset_accumulators(self.handle), self.handle)
        else:
            check(self.lib.phq_reset_accumulators_async(self.handle, C.c_void_p(stream.cuda_stream)), self.handle)

    def accumulator_tensors(self):
        """The handle's accumulator buffer as two torch views (int64 bit pattern of the u64 plane, float64 plane)."""
        import torch
        pointer, n_u64, n_f64 = C.c_void_p(), C.c_int64(), C.c_int64()
        check(self.lib.phq_accumulator_buffer(self.handle, C.byref(pointer), C.byref(n_u64), C.byref(n_f64)), self.handle)

        class _Raw:
            def __init__(self, address, count, typestr):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (address, False), "version": 2}
        device = torch.device("cuda", self.device)
        u = torch.as_tensor(_Raw(pointer.value, n_u64.value, "<i8"), device=device)
        f = torch.as_tensor(_Raw(pointer.value + 8 * n_u64.value, n_f64.value, "<f8"), device=device)
        return u, f

    def communicator(self, group=None):
        """The ncclComm_t phq_collect reduces over: one rank per process of `group`, created once per (group, device)
        from an id that rank 0 draws (phq_comm_unique_id) and torch.distributed carries to the other ranks."""
        import torch.distributed as dist
        key = (id(group), self.device)
        if key not in _COMMUNICATORS:
            ident = (C.c_uint8 * 128)()
            if dist.get_rank(group) == 0:
                check(self.lib.phq_comm_unique_id(ident))
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            ident = (C.c_uint8 * 128)(*box[0])
            comm = C.c_void_p()
            check(self.lib.phq_comm_create(ident, dist.get_rank(group), dist.get_world_size(group), self.device, C.byref(comm)))
            _COMMUNICATORS[key] = comm
        return _COMMUNICATORS[key]

    def collect(self, group=None, stream=None):
        """Classifier::collect across ranks (classifier.h:87-93): phq_collect, one grouped in-place ncclAllReduce(sum)
        over the two accumulator planes, asynchronous on `stream` (default: torch's current stream). Afterwards the
        handle answers for the whole job and refuses to accumulate again until reset()."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        stream = stream if stream is not None else torch.cuda.current_stream(torch.device("cuda", self.device))
        check(self.lib.phq_collect(self.handle, self.communicator(group), C.c_void_p(stream.cuda_stream)), self.handle)

    def report(self, incoming=(0, 0), precision: int = 15, text: bool = False):
        """The decoder sections of the job report (phq_report) from this handle's device accumulators."""
        out = C.c_void_p()
        check(self.lib.phq_report(self.handle, int(incoming[0]), int(incoming[1]), precision, C.byref(out)), self.handle)
        try:
            raw = C.string_at(out).decode()
            return raw if text else json.loads(raw)
        finally:
            self.lib.phq_free(out)

    def encode_report(self, tables, totals, incoming=(0, 0), precision: int = 15, text: bool = False):
        """phq_encode_report: the same report from caller-held tables [(u64 [(N+1), 6], f64 [(N+1), 2]) per decoder]
        and chain totals (count, pf_count). Pure host work: also valid on a host-only handle."""
        u = [np.ascontiguousarray(t[0], dtype=np.uint64) for t in tables]
        f = [np.ascontiguousarray(t[1], dtype=np.float64) for t in tables]
        pu = (C.c_void_p * len(u))(*[a.ctypes.data for a in u])
        pf = (C.c_void_p * len(f))(*[a.ctypes.data for a in f])
        out = C.c_void_p()
        check(self.lib.phq_encode_report(self.handle, pu, pf, int(totals[0]), int(totals[1]), int(incoming[0]), int(incoming[1]),
                                         precision, C.byref(out)), self.handle)
        try:
            raw = C.string_at(out).decode()
            return raw if text else json.loads(raw)
        finally:
            self.lib.phq_free(out)

    def estimate_priors(self, k: int):
        noise = C.c_double()
        concentration = np.zeros(self.info[k].barcode_cardinality, dtype=np.float64)
        check(self.lib.phq_estimate_priors(self.handle, k, C.byref(noise), concentration.ctypes.data), self.handle)
        return noise.value, concentration

    def set_priors(self, k: int, noise: float, concentration):
        c = np.ascontiguousarray(concentration, dtype=np.float64)
        check(self.lib.phq_set_priors(self.handle, k, float(noise), c.ctypes.data), self.handle)

    def adjust_priors(self):
        """The two-pass workflow of docs/pamld.md:38-44 / Classifier::adjust_prior (classifier.h:125-160):
        replace every PAMLD decoder's priors by the estimates from the accumulated (collected) counts."""
        for k, info in enumerate(self.info):
            if info.algorithm == 0:
                noise, concentration = self.estimate_priors(k)
                self.set_priors(k, noise, concentration)

    def reference_power(self, sigma):
        """pow(B, sigma) as the tie pass forms it (phq_reference_power)."""
        sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        out = np.zeros_like(sigma)
        check(self.lib.phq_reference_power(self.handle, sigma.size, sigma.ctypes.data, out.ctypes.data), self.handle)
        return out

    def statistics(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self.lib.phq_statistics(self.handle, C.byref(a), C.byref(b), C.byref(c)), self.handle)
        return {"kernel_launches": a.value, "exact_path_reads": b.value, "threshold_band_reads": c.value}


_COMMUNICATORS = {}


def all_reduce_accumulators(u64_plane, f64_plane, group=None):
    """Sum accumulator planes held as torch tensors in place over torch.distributed: what phq_collect does inside the
    library, for hosts that keep their tables elsewhere (the gloo test reduces CPU tables with it). u64 counters
    travel as int64 bit patterns (two's complement addition is the same operation)."""
    import torch.distributed as dist
    dist.all_reduce(u64_plane, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(f64_plane, op=dist.ReduceOp.SUM, group=group)
