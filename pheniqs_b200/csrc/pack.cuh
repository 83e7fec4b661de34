/*  pack.cuh — feed bytes -> tiles on the device.

    The host path (phq_pack, api.cu) needs the feed's decoded form (one BAM code and one Phred byte per base). This is
    the same step from the bytes of the FASTQ record itself: AsciiToAmbiguousBam and `quality - phred offset`
    (fastq.h:55-78), token slicing and reverse complement (transform.h:65-80, 142-169), and the tile packing, in one
    kernel, so the host only copies the barcode-bearing segments as they sit in its feed buffers. */
#ifndef PHQ_PACK_CUH
#define PHQ_PACK_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pheniqs_b200.h"

namespace phq {

constexpr int PACK_MAX_TOKENS = 16;
constexpr int PACK_MAX_INPUT_SEGMENTS = 8;

/* one input segment of the batch on the device; `sequence` / `quality` are biased so that offset[r] (or r * length)
   of the FIRST read of the launch indexes its first byte */
struct RawSegmentView {
    const uint8_t* sequence;
    const uint8_t* quality;
    const long long* offset;        /* [reads + 1] absolute offsets, or NULL */
    long long length;               /* bytes per read when offset == NULL */
    long long first;                /* index of the launch's first read in `offset` / in units of `length` */
};

struct PackToken {                  /* TransformSpec (spec.hpp) */
    int32_t input_segment;
    int32_t start;
    int32_t end;
    int32_t end_terminated;
    int32_t output_segment;
    int32_t reverse_complement;
};

struct PackPlan {
    int32_t token_cardinality;
    int32_t segment_cardinality;
    int32_t nucleotide_cardinality;
    int32_t stale_semantics;        /* PAMLD: expected length is read, terminator then bytes of earlier reads (barcode.h:150) */
    int32_t phred_offset;
    int32_t bam_input;              /* the segments hold the reference's in-memory form (one BAM 4-bit code and one Phred byte per base, sequence.h:264-300) instead of FASTQ text */
    int32_t segment_offset[PHQ_MAX_SEGMENTS + 1];
    PackToken token[PACK_MAX_TOKENS];
    /* the Observation as the reads before this launch left it, by concatenated position: BAM code and Phred byte */
    uint8_t carry_code[PHQ_MAX_NUCLEOTIDES];
    uint8_t carry_quality[PHQ_MAX_NUCLEOTIDES];
    RawSegmentView input[PACK_MAX_INPUT_SEGMENTS];
};

/* ------------------------------------------------------------------ tag synthesis
   What Read::flush and Auxiliary::encode (read.h:187-237, auxiliary.cpp:320-361) append to every output record
   for the decoders of this path, as the BAM auxiliary bytes themselves, in the reference's order:
       RG:Z  BC:Z QT:Z XB:f  RX:Z QX:Z OX:Z BZ:Z XM:f  CB:Z CR:Z CY:Z XC:f
   from the raw segments (raw barcodes and their qualities), the per-decoder results (read group, corrected
   barcodes, error probabilities) and the barcode tables. */
constexpr int TAG_MAX_DECODERS = 6;
constexpr int TAG_MAX_TOKENS = 8;

struct TagDecoder {
    int32_t topic;                  /* phq_topic: 0 sample, 1 molecular, 2 cellular */
    int32_t algorithm;              /* phq_algorithm */
    int32_t corrected_quality;
    int32_t token_cardinality;
    int32_t segment_cardinality;
    int32_t nucleotide_cardinality;
    int32_t segment_offset[PHQ_MAX_SEGMENTS + 1];
    PackToken token[TAG_MAX_TOKENS];
    const phq_result* results;      /* device, [reads of the launch]; NULL for a naive decoder */
    const uint8_t* barcode_code;    /* device, [(N + 1)][nucleotide_cardinality] BAM codes; row 0 = undetermined ('=') */
};

struct TagPlan {
    int32_t decoder_cardinality;
    int32_t phred_offset;
    int32_t bam_input;              /* as PackPlan */
    int32_t stride;                 /* bytes per record in `aux` (multiple of 4) */
    const uint8_t* read_group_text; /* device: read group IDs of the sample decoder, row i at [offset[i], offset[i + 1]) */
    const int32_t* read_group_offset;
    RawSegmentView input[PACK_MAX_INPUT_SEGMENTS];
    TagDecoder decoder[TAG_MAX_DECODERS];
};

/* aux[r * stride ..] receives the auxiliary bytes of read r (zero padded), aux_length[r] their count */
cudaError_t launch_tags(const TagPlan& plan, long long n_reads, uint8_t* aux, int32_t* aux_length, int multiprocessor_count, cudaStream_t stream);

/* tile planes of one decoder for `n_reads` reads from the raw segments; asynchronous on `stream` */
cudaError_t launch_pack(const PackPlan& plan, long long n_reads, uint32_t* bases, uint16_t* nmask, uint32_t* quality, long long pitch,
                        int multiprocessor_count, cudaStream_t stream);

}   /* namespace phq */
#endif
