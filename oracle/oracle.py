"""oracle.py — TEST INFRASTRUCTURE: Python front-end of the CPU checkers.

Two checkers live behind the same `decode()` signature:

* ``PortOracle``  — oracle/libpheniqs_oracle.so, the plain-C restatement (pheniqs_oracle.c).
* ``RefOracle``   — oracle/_ref/libpheniqs_ref.so, the reference's OWN decoder classes compiled
  from /root/reference by oracle/Makefile (present in the build container; the prebuilt
  .so travels to the GPU box).

Also here, because the checkers need them and the product must not be used to check itself:
a Python restatement of the reference's decoder *compile* step (transcode.cpp:735-768,
824-1039; metric.h:87-111,216-242; defaults configuration.json:423-501), token / knit parsing
(transform.cpp:100-331) and the FASTQ byte conventions (fastq.h:55-78, iupac.h:153-171).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. Nothing under pheniqs_b200/ does.
"""
from __future__ import annotations

import copy
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIBRARY = os.path.join(HERE, "libpheniqs_oracle.so")
REF_LIBRARY = os.path.join(HERE, "_ref", "libpheniqs_ref.so")
BINDING_LIBRARY = os.path.join(HERE, "_ref", "libpheniqs_binding.so")

ALGORITHM = {"pamld": 0, "mdd": 1, "naive": 2, "passthrough": 3}
TOPIC = {"sample": 0, "molecular": 1, "cellular": 2}
TOPIC_ORDER = ("sample", "molecular", "cellular")        # transcode.h:51-60

# iupac.h:153-171 AsciiToAmbiguousBam
_IUPAC = {"=": 0, "A": 1, "C": 2, "M": 3, "G": 4, "R": 5, "S": 6, "V": 7, "T": 8, "W": 9, "Y": 10, "H": 11, "K": 12, "D": 13, "B": 14, "N": 15}
ASCII_TO_BAM = np.full(256, 15, dtype=np.uint8)
for _c, _v in _IUPAC.items():
    ASCII_TO_BAM[ord(_c)] = _v
    ASCII_TO_BAM[ord(_c.lower())] = _v
for _i, _v in enumerate((1, 2, 4, 8)):
    ASCII_TO_BAM[ord("0") + _i] = _v
BAM_TO_ASCII = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)


def build(target: str = "all") -> None:
    """Compile the checkers (make -C oracle). Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


# ------------------------------------------------------------------ decoder compile (restated)
DECODER_DEFAULT = {          # configuration.json:423-501 (projection *:decoder), :368-376 (default)
    "sample": {"algorithm": "pamld"},
    "cellular": {"algorithm": "pamld"},
    "molecular": {"algorithm": "naive"},
}
COMMON_DEFAULT = {
    "confidence threshold": 0.95,
    "high quality distance threshold": 0,
    "high quality threshold": 30,
    "noise": 0.01,
    "quality masking threshold": 0,
    "corrected quality": 30,
}
_TOKEN = re.compile(r"^([0-9]+):(-?[0-9]+)?:(-?[0-9]+)?$")       # configuration.json:1427


def parse_token(pattern: str):
    """transform.cpp:100-127: 'segment:start:end' -> (segment, start, end, end_terminated)."""
    m = _TOKEN.match(pattern)
    if not m:
        raise ValueError("illegal token syntax " + pattern)
    segment = int(m.group(1))
    start = int(m.group(2)) if m.group(2) else 0
    end_terminated = bool(m.group(3))
    end = int(m.group(3)) if m.group(3) else 0
    return segment, start, end, end_terminated


def token_length(start, end, end_terminated):
    """Token::length / constant / empty (transform.h:47-64); None when not fixed width."""
    if end_terminated:
        constant = (start >= 0 and end >= 0) or (start < 0 and end < 0)
        if not constant:
            return None
        return 0 if start >= end else end - start
    return -start if start < 0 else None


def parse_rule(transform: dict):
    """Rule decode (transform.cpp:252-331): knit elements 'i:~j:k' -> ordered transforms."""
    tokens = [parse_token(t) for t in transform["token"]]
    knit = transform.get("knit") or [str(i) for i in range(len(tokens))]     # transcode.cpp:735-768
    transforms = []
    for output_segment, element in enumerate(knit):
        for part in element.split(":"):
            reverse = part.startswith("~")
            index = int(part[1:] if reverse else part)
            segment, start, end, end_terminated = tokens[index]
            transforms.append((segment, start, end, int(end_terminated), output_segment, int(reverse)))
    return tokens, knit, transforms, len(knit)


def shannon_bound(words):
    """WordMetric::find_shannon_bound over the DISTINCT words of one segment (metric.h:87-111)."""
    words = sorted(set(words))
    if not words:
        return 0
    minimum = len(words[0])
    a = np.array([list(w.encode()) for w in words], dtype=np.uint8)
    for i in range(len(words) - 1):
        d = (a[i + 1:] != a[i]).sum(axis=1).min()
        minimum = min(minimum, int(d))
    return int((minimum - 1) / 2)           # C++ integer division truncates toward zero


def compile_decoder(decoder: dict, topic: str, index: int = 0, compute_tolerance: bool = True) -> dict:
    """Python restatement of Transcode::compile_decoder (+_transformation) for ONE decoder."""
    value = dict(COMMON_DEFAULT)
    value.update(DECODER_DEFAULT[topic])
    value.update(copy.deepcopy(decoder))
    value["index"] = index
    value.setdefault("multiplexing classifier", False)

    tokens, knit, transforms, segment_cardinality = parse_rule(value["transform"])
    value["transform"] = {"token": list(value["transform"]["token"]), "knit": knit}
    barcode_length = [0] * segment_cardinality
    nucleotide_cardinality = 0
    for (segment, start, end, end_terminated, output_segment, reverse) in transforms:
        length = token_length(start, end, bool(end_terminated))
        if length is None:
            raise ValueError("token is not fixed width")
        if length == 0:
            raise ValueError("token is empty")
        barcode_length[output_segment] += length
        nucleotide_cardinality += length
    value["segment cardinality"] = segment_cardinality
    value["nucleotide cardinality"] = nucleotide_cardinality
    value["barcode length"] = barcode_length
    lower_bound = 1.0 / float(pow(4, nucleotide_cardinality))
    if "random barcode probability" in value:
        if value["random barcode probability"] < lower_bound:
            raise ValueError("random barcode probability is smaller than lower bound")
    else:
        value["random barcode probability"] = lower_bound

    noise = float(value["noise"])
    undetermined = dict(value.get("undetermined") or {})
    undetermined.update({"index": 0, "concentration": noise, "segment cardinality": segment_cardinality,
                         "barcode": ["=" * n for n in barcode_length]})
    undetermined.setdefault("ID", "undetermined")
    value["undetermined"] = undetermined

    codec = value.get("codec") or {}
    if codec:
        compiled = {}
        total = 0.0
        seen = set()
        for position, key in enumerate(sorted(codec.keys())):         # json.cpp:875-893 key sort -> index order
            record = dict(codec[key])
            barcode = list(record["barcode"]) if "barcode" in record else None
            if barcode is None:
                raise ValueError("barcode missing in " + key)
            if len(barcode) != segment_cardinality:
                raise ValueError("expected %d segments in barcode %s" % (segment_cardinality, key))
            for i, segment in enumerate(barcode):
                if len(segment) != barcode_length[i]:
                    raise ValueError("expected %d nucleotides in segment %d of barcode %s" % (barcode_length[i], i, key))
            flat = "".join(barcode)
            if flat in seen:
                raise ValueError("duplicate barcode sequence " + flat)
            seen.add(flat)
            record["index"] = position + 1
            record["segment cardinality"] = segment_cardinality
            record.setdefault("concentration", 1)
            if record["concentration"] < 0:
                raise ValueError("barcode concentration must be a positive number")
            total += float(record["concentration"])
            record.setdefault("ID", "-".join(barcode))
            compiled[key] = record
        if not total > 0:
            raise ValueError("total pool concentration is not a positive number")
        factor = (1.0 - noise) / total
        for record in compiled.values():
            record["concentration"] = float(record["concentration"]) * factor
        value["codec"] = compiled
        value["barcode cardinality"] = len(compiled) + 1
        if compute_tolerance:
            bound = [shannon_bound([r["barcode"][i] for r in compiled.values()]) for i in range(segment_cardinality)]
            value["shannon bound"] = bound
            if "distance tolerance" in value and value["distance tolerance"] is not None:
                tolerance = list(value["distance tolerance"])
                if len(tolerance) != segment_cardinality:
                    raise ValueError("distance tolerance cardinality inconsistant with barcode segment cardinality")
                for i, t in enumerate(tolerance):
                    if t > bound[i]:
                        raise ValueError("barcode tolerance for segment %d is higher than shannon bound %d" % (i, bound[i]))
            else:
                value["distance tolerance"] = bound
        else:
            value.setdefault("distance tolerance", [0] * segment_cardinality)
    else:
        value.setdefault("distance tolerance", [0] * segment_cardinality)
    for key in ("confidence threshold", "noise"):               # transcode.cpp:1540-1565
        if not 0 <= value[key] <= 1:
            raise ValueError(key + " out of range")
    return value


def compile_job(job: dict, compute_tolerance: bool = True) -> dict:
    """Compile {'sample': {...}, 'molecular': [...], 'cellular': [...]} the way compile_topic does."""
    out = {}
    for topic in TOPIC_ORDER:
        if topic not in job or job[topic] is None:
            continue
        element = job[topic]
        if isinstance(element, dict):
            out[topic] = compile_decoder(element, topic, 0, compute_tolerance)
        else:
            out[topic] = [compile_decoder(e, topic, i, compute_tolerance) for i, e in enumerate(element)]
    return out


def decoder_chain(compiled_job: dict):
    """[(topic, decoder)] in classification order (transcode.h:51-60)."""
    chain = []
    for topic in TOPIC_ORDER:
        element = compiled_job.get(topic)
        if element is None:
            continue
        for decoder in ([element] if isinstance(element, dict) else element):
            chain.append((topic, decoder))
    return chain


# ------------------------------------------------------------------ read batches
class ReadBatch:
    """Input reads in the reference's in-memory convention: per input segment a flat array of
    BAM 4-bit codes (one byte per base), a flat array of Phred values (offset removed) and
    n_reads + 1 offsets; plus the incoming qcfail flags."""

    def __init__(self, code, quality, offset, qcfail=None):
        self.code = [np.ascontiguousarray(c, dtype=np.uint8) for c in code]
        self.quality = [np.ascontiguousarray(q, dtype=np.uint8) for q in quality]
        self.offset = [np.ascontiguousarray(o, dtype=np.int64) for o in offset]
        self.n_segments = len(self.code)
        self.n_reads = int(self.offset[0].shape[0] - 1) if self.offset else 0
        self.qcfail = None if qcfail is None else np.ascontiguousarray(qcfail, dtype=np.uint8)

    @classmethod
    def from_fixed(cls, code_matrices, quality_matrices, qcfail=None):
        """Segments of fixed width: one [n_reads, width] matrix per segment."""
        code, quality, offset = [], [], []
        for c, q in zip(code_matrices, quality_matrices):
            n, w = c.shape
            code.append(c.reshape(-1))
            quality.append(q.reshape(-1))
            offset.append(np.arange(n + 1, dtype=np.int64) * w)
        return cls(code, quality, offset, qcfail)

    @classmethod
    def from_fastq(cls, paths, phred_offset=33):
        """fastq.h:55-78: ASCII -> BAM code, quality byte - phred offset; comment 'x:Y:...' -> qcfail (fastq.h:253-285)."""
        code, quality, offset = [], [], []
        qcfail = None
        for path in paths:
            with open(path, "rb") as f:
                lines = f.read().split(b"\n")
            if lines and lines[-1] == b"":
                lines.pop()
            names, sequences, qualities = lines[0::4], lines[1::4], lines[3::4]
            lengths = np.array([len(s) for s in sequences], dtype=np.int64)
            o = np.zeros(len(sequences) + 1, dtype=np.int64)
            np.cumsum(lengths, out=o[1:])
            code.append(ASCII_TO_BAM[np.frombuffer(b"".join(sequences), dtype=np.uint8)])
            quality.append((np.frombuffer(b"".join(qualities), dtype=np.uint8) - phred_offset).astype(np.uint8))
            offset.append(o)
            if qcfail is None:
                qcfail = np.zeros(len(names), dtype=np.uint8)
                for i, name in enumerate(names):
                    parts = name.split(b" ", 1)
                    if len(parts) == 2:
                        fields = parts[1].split(b":")
                        if len(fields) > 1 and fields[1] == b"Y":
                            qcfail[i] = 1
        return cls(code, quality, offset, qcfail)

    def select(self, keep):
        """Sub-batch of the reads where keep[r] is true (order preserved)."""
        keep = np.asarray(keep, dtype=bool)
        code, quality, offset = [], [], []
        for c, q, o in zip(self.code, self.quality, self.offset):
            lengths = (o[1:] - o[:-1])[keep]
            index = np.concatenate([np.arange(o[r], o[r + 1]) for r in np.nonzero(keep)[0]]) if keep.any() else np.zeros(0, dtype=np.int64)
            no = np.zeros(lengths.shape[0] + 1, dtype=np.int64)
            np.cumsum(lengths, out=no[1:])
            code.append(c[index])
            quality.append(q[index])
            offset.append(no)
        return ReadBatch(code, quality, offset, None if self.qcfail is None else self.qcfail[keep])

    def _pointers(self):
        P8 = C.POINTER(C.c_uint8)
        P64 = C.POINTER(C.c_int64)
        code = (P8 * self.n_segments)(*[c.ctypes.data_as(P8) for c in self.code])
        quality = (P8 * self.n_segments)(*[q.ctypes.data_as(P8) for q in self.quality])
        offset = (P64 * self.n_segments)(*[o.ctypes.data_as(P64) for o in self.offset])
        qcfail = None if self.qcfail is None else self.qcfail.ctypes.data_as(P8)
        return code, quality, offset, qcfail


class DecodeResult:
    def __init__(self, n_reads, n_decoders):
        self.index = np.zeros((n_reads, n_decoders), dtype=np.int32)
        self.distance = np.zeros((n_reads, n_decoders), dtype=np.int32)
        self.confidence = np.zeros((n_reads, n_decoders), dtype=np.float64)
        self.qcfail = np.zeros(n_reads, dtype=np.uint8)
        self.read_distance = np.zeros((n_reads, 3), dtype=np.uint32)
        self.read_confidence = np.zeros((n_reads, 3), dtype=np.float64)
        self.channel = np.zeros(n_reads, dtype=np.int32)
        self.seconds = 0.0


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ------------------------------------------------------------------ the C restatement
class _Transform(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("input_segment_index", "start", "end", "end_terminated", "output_segment_index", "reverse_complement")]


class _Decoder(C.Structure):
    _fields_ = [
        ("algorithm", C.c_int32), ("topic", C.c_int32), ("n_barcodes", C.c_int32), ("n_segments", C.c_int32),
        ("nucleotide_cardinality", C.c_int32), ("n_transforms", C.c_int32),
        ("transform", C.POINTER(_Transform)), ("segment_length", C.POINTER(C.c_int32)),
        ("barcode", C.POINTER(C.c_uint8)), ("concentration", C.POINTER(C.c_double)),
        ("noise", C.c_double), ("confidence_threshold", C.c_double), ("random_barcode_probability", C.c_double),
        ("high_quality_threshold", C.c_int32), ("high_quality_distance_threshold", C.c_int32),
        ("quality_masking_threshold", C.c_int32), ("distance_tolerance", C.POINTER(C.c_int32)),
        ("multiplexing_classifier", C.c_int32),
    ]


def flat_spec(topic: str, decoder: dict):
    """Compiled decoder JSON -> (ctypes struct, keep-alive list, barcode matrix, priors)."""
    _, _, transforms, n_segments = parse_rule(decoder["transform"])
    keep = []
    tarray = (_Transform * len(transforms))(*[_Transform(*t) for t in transforms])
    lengths = np.array(decoder.get("barcode length") or [0] * n_segments, dtype=np.int32)
    codec = decoder.get("codec") or {}
    records = sorted(codec.values(), key=lambda r: r["index"])
    L = int(decoder.get("nucleotide cardinality", int(lengths.sum())))
    barcode = np.zeros((len(records), L), dtype=np.uint8)
    prior = np.zeros(len(records), dtype=np.float64)
    for i, record in enumerate(records):
        barcode[i] = ASCII_TO_BAM[np.frombuffer("".join(record["barcode"]).encode(), dtype=np.uint8)]
        prior[i] = record["concentration"]
    tolerance = np.array(decoder.get("distance tolerance") or [0] * n_segments, dtype=np.int32)
    keep += [tarray, lengths, barcode, prior, tolerance]
    spec = _Decoder(
        ALGORITHM[decoder["algorithm"]], TOPIC[topic], len(records), n_segments, L, len(transforms),
        tarray, _p(lengths, C.c_int32), _p(barcode, C.c_uint8), _p(prior, C.c_double),
        float(decoder.get("noise", 0)), float(decoder.get("confidence threshold", 0)), float(decoder.get("random barcode probability", 0)),
        int(decoder.get("high quality threshold", 0)), int(decoder.get("high quality distance threshold", 0)),
        int(decoder.get("quality masking threshold", 0)), _p(tolerance, C.c_int32),
        int(bool(decoder.get("multiplexing classifier", False))))
    return spec, keep, barcode, prior


class PortOracle:
    """The plain-C restatement, driven through ctypes."""
    kind = "port"

    def __init__(self, compiled_job: dict):
        if not os.path.exists(PORT_LIBRARY):
            build("port")
        self.lib = C.CDLL(PORT_LIBRARY)
        self.lib.phqo_create.restype = C.c_void_p
        self.lib.phqo_decode_threaded.restype = C.c_double
        self.chain = decoder_chain(compiled_job)
        self.n_decoders = len(self.chain)
        self._keep = []
        specs = (_Decoder * self.n_decoders)()
        self.n_barcodes = []
        for k, (topic, decoder) in enumerate(self.chain):
            spec, keep, barcode, _ = flat_spec(topic, decoder)
            specs[k] = spec
            self._keep.append(keep)
            self.n_barcodes.append(barcode.shape[0])
        self.handle = C.c_void_p(self.lib.phqo_create(self.n_decoders, specs))

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.phqo_destroy(self.handle)
            self.handle = None

    def decode(self, batch: ReadBatch, threads: int = 1, want_outputs: bool = True) -> DecodeResult:
        out = DecodeResult(batch.n_reads if want_outputs else 0, self.n_decoders)
        code, quality, offset, qcfail = batch._pointers()
        if threads <= 1:
            import time
            t0 = time.perf_counter()
            if want_outputs:
                self.lib.phqo_decode(self.handle, C.c_int64(batch.n_reads), batch.n_segments, code, quality, offset, qcfail,
                                     _p(out.index, C.c_int32), _p(out.distance, C.c_int32), _p(out.confidence, C.c_double),
                                     _p(out.qcfail, C.c_uint8), _p(out.read_distance, C.c_uint32), _p(out.read_confidence, C.c_double), _p(out.channel, C.c_int32))
            else:
                self.lib.phqo_decode(self.handle, C.c_int64(batch.n_reads), batch.n_segments, code, quality, offset, qcfail,
                                     None, None, None, None, None, None, None)
            out.seconds = time.perf_counter() - t0
        else:
            args = (None, None, None, None) if not want_outputs else (_p(out.index, C.c_int32), _p(out.distance, C.c_int32), _p(out.confidence, C.c_double), _p(out.qcfail, C.c_uint8))
            out.seconds = self.lib.phqo_decode_threaded(self.handle, threads, C.c_int64(batch.n_reads), batch.n_segments, code, quality, offset, qcfail, *args)
        return out

    def extract(self, k: int, batch: ReadBatch):
        L = int(self.chain[k][1]["nucleotide cardinality"])
        ns = int(self.chain[k][1]["segment cardinality"])
        code = np.zeros((batch.n_reads, L), dtype=np.uint8)
        quality = np.zeros((batch.n_reads, L), dtype=np.uint8)
        length = np.zeros((batch.n_reads, ns), dtype=np.int32)
        c, q, o, _ = batch._pointers()
        self.lib.phqo_extract(self.handle, k, C.c_int64(batch.n_reads), batch.n_segments, c, q, o,
                              _p(code, C.c_uint8), _p(quality, C.c_uint8), _p(length, C.c_int32))
        return code, quality, length

    def accumulators(self, k: int):
        nb = self.n_barcodes[k]
        u = np.zeros((nb + 1, 6), dtype=np.uint64)
        f = np.zeros((nb + 1, 2), dtype=np.float64)
        self.lib.phqo_accumulators(self.handle, k, _p(u, C.c_uint64), _p(f, C.c_double))
        return u, f

    def totals(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.phqo_totals(self.handle, C.byref(a), C.byref(b))
        return a.value, b.value

    def reset(self):
        self.lib.phqo_reset(self.handle)

    def estimate_priors(self, k: int, tables=None):
        u, f = tables if tables is not None else self.accumulators(k)
        u = np.ascontiguousarray(u, dtype=np.uint64)
        f = np.ascontiguousarray(f, dtype=np.float64)
        nb = u.shape[0] - 1
        noise = C.c_double()
        concentration = np.zeros(nb, dtype=np.float64)
        self.lib.phqo_estimate_priors(nb, _p(u, C.c_uint64), _p(f, C.c_double), C.byref(noise), _p(concentration, C.c_double))
        return noise.value, concentration


def phred_tables():
    if not os.path.exists(PORT_LIBRARY):
        build("port")
    lib = C.CDLL(PORT_LIBRARY)
    tq = np.zeros(128, dtype=np.float64)
    u, b = C.c_double(), C.c_double()
    lib.phqo_phred_tables(_p(tq, C.c_double), C.byref(u), C.byref(b))
    return tq, u.value, b.value


# ------------------------------------------------------------------ the reference's own classes
def ref_available() -> bool:
    return os.path.exists(REF_LIBRARY)


class RefOracle:
    """oracle/_ref: the reference's own PamlDecoder / MdDecoder / NaiveMolecularDecoder classes."""
    kind = "reference"

    def __init__(self, compiled_job: dict, input_segment_cardinality: int = 0):
        if not ref_available():
            raise RuntimeError("oracle/_ref/libpheniqs_ref.so is not built (needs /root/reference; run make -C oracle ref)")
        self.lib = C.CDLL(REF_LIBRARY)
        self.lib.phq_ref_create.restype = C.c_void_p
        self.lib.phq_ref_decode.restype = C.c_double
        self.lib.phq_ref_report.restype = C.c_char_p
        self.lib.phq_ref_last_error.restype = C.c_char_p
        self.chain = decoder_chain(compiled_job)
        self.n_decoders = len(self.chain)
        error = C.create_string_buffer(4096)
        text = json.dumps(compiled_job).encode()
        self.handle = C.c_void_p(self.lib.phq_ref_create(text, input_segment_cardinality, error, 4096))
        if not self.handle:
            raise RuntimeError("reference decoder construction failed: " + error.value.decode())
        self.n_barcodes = [self.lib.phq_ref_barcode_count(self.handle, k) for k in range(self.n_decoders)]

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.phq_ref_destroy(self.handle)
            self.handle = None

    def decode(self, batch: ReadBatch, threads: int = 1, want_outputs: bool = True) -> DecodeResult:
        out = DecodeResult(batch.n_reads if want_outputs else 0, self.n_decoders)
        code, quality, offset, qcfail = batch._pointers()
        if want_outputs:
            args = (_p(out.index, C.c_int32), _p(out.distance, C.c_int32), _p(out.confidence, C.c_double),
                    _p(out.qcfail, C.c_uint8), _p(out.read_distance, C.c_uint32), _p(out.read_confidence, C.c_double), _p(out.channel, C.c_int32))
        else:
            args = (None,) * 7
        seconds = self.lib.phq_ref_decode(self.handle, C.c_int64(batch.n_reads), batch.n_segments, code, quality, offset, qcfail, threads, *args)
        if seconds < 0:
            raise RuntimeError(self.lib.phq_ref_last_error(self.handle).decode())
        out.seconds = seconds
        return out

    TAG_NAMES = ("RG", "BC", "QT", "RX", "QX", "OX", "BZ", "CB", "CR", "CY")

    def tags(self, batch: ReadBatch, stride: int = 256):
        """Per read, the tags Read::flush / Auxiliary::encode produce (read.h:187-237, auxiliary.cpp:320-361):
        [{"RG": ..., "BC": ..., ..., "XB": float32, ...}] with absent tags left out, and the final qcfail flags."""
        code, quality, offset, qcfail = batch._pointers()
        text = np.zeros((batch.n_reads, 10, stride), dtype=np.uint8)
        probability = np.zeros((batch.n_reads, 3), dtype=np.float32)
        flags = np.zeros(batch.n_reads, dtype=np.uint8)
        status = self.lib.phq_ref_tags(self.handle, C.c_int64(batch.n_reads), batch.n_segments, code, quality, offset, qcfail, stride,
                                       _p(text, C.c_char), _p(probability, C.c_float), _p(flags, C.c_uint8))
        if status != 0:
            raise RuntimeError(self.lib.phq_ref_last_error(self.handle).decode())
        out = []
        for r in range(batch.n_reads):
            record = {}
            for t, name in enumerate(self.TAG_NAMES):
                value = text[r, t].tobytes().split(b"\0", 1)[0]
                if value:
                    record[name] = value.decode("latin-1")
            for t, name in enumerate(("XB", "XM", "XC")):
                if probability[r, t] > 0:
                    record[name] = np.float32(probability[r, t])
            out.append(record)
        return out, flags

    def accumulators(self, k: int):
        nb = self.n_barcodes[k]
        u = np.zeros((nb + 1, 6), dtype=np.uint64)
        f = np.zeros((nb + 1, 2), dtype=np.float64)
        self.lib.phq_ref_accumulators(self.handle, k, _p(u, C.c_uint64), _p(f, C.c_double))
        return u, f

    def totals(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.phq_ref_totals(self.handle, C.byref(a), C.byref(b))
        return a.value, b.value

    def estimate_priors(self, k: int):
        nb = self.n_barcodes[k]
        noise = C.c_double()
        concentration = np.zeros(nb, dtype=np.float64)
        self.lib.phq_ref_estimated_priors(self.handle, k, C.byref(noise), _p(concentration, C.c_double))
        return noise.value, concentration

    def report(self, k: int, precision: int = 15) -> dict:
        return json.loads(self.lib.phq_ref_report(self.handle, k, precision).decode())


def binding_available() -> bool:
    return os.path.exists(BINDING_LIBRARY)


def batched_binding(compiled_job: dict, batch: ReadBatch, device: int = 0, batch_reads: int = 2048, stride: int = 256):
    """oracle/_ref/libpheniqs_binding.so (batched_binding.cpp): the reference-side binding — the reference's own Read /
    Segment / decoder classes and Read::flush around the product's phq_decode_batch_bam on `device`. Returns what
    RefOracle.tags returns (per read tag dict, final qcfail flags) plus the job report of the device accumulators.
    Needs a GPU: the binding has no scoring code of its own."""
    if not binding_available():
        raise RuntimeError("oracle/_ref/libpheniqs_binding.so is not built (needs /root/reference; run make -C oracle ref)")
    lib = C.CDLL(BINDING_LIBRARY)
    lib.phq_binding_last_error.restype = C.c_char_p
    code, quality, offset, qcfail = batch._pointers()
    text = np.zeros((batch.n_reads, 10, stride), dtype=np.uint8)
    probability = np.zeros((batch.n_reads, 3), dtype=np.float32)
    flags = np.zeros(batch.n_reads, dtype=np.uint8)
    report = C.c_void_p()
    status = lib.phq_binding_run(json.dumps(compiled_job).encode(), device, C.c_int64(batch.n_reads), batch.n_segments, C.c_int64(batch_reads),
                                 code, quality, offset, qcfail, stride, _p(text, C.c_char), _p(probability, C.c_float), _p(flags, C.c_uint8), C.byref(report))
    if status != 0:
        raise RuntimeError(lib.phq_binding_last_error().decode())
    report_json = json.loads(C.string_at(report).decode())
    C.CDLL(None).free(report)
    out = []
    for r in range(batch.n_reads):
        record = {}
        for t, name in enumerate(RefOracle.TAG_NAMES):
            value = text[r, t].tobytes().split(b"\0", 1)[0]
            if value:
                record[name] = value.decode("latin-1")
        for t, name in enumerate(("XB", "XM", "XC")):
            if probability[r, t] > 0:
                record[name] = np.float32(probability[r, t])
        out.append(record)
    return out, flags, report_json


def best_oracle(compiled_job: dict, input_segment_cardinality: int = 0):
    """The reference's own classes when their build is present, else the C restatement."""
    if ref_available():
        return RefOracle(compiled_job, input_segment_cardinality)
    return PortOracle(compiled_job)
