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
    int32_t segment_offset[PHQ_MAX_SEGMENTS + 1];
    PackToken token[PACK_MAX_TOKENS];
    /* the Observation as the reads before this launch left it, by concatenated position: BAM code and Phred byte */
    uint8_t carry_code[PHQ_MAX_NUCLEOTIDES];
    uint8_t carry_quality[PHQ_MAX_NUCLEOTIDES];
    RawSegmentView input[PACK_MAX_INPUT_SEGMENTS];
};

/* tile planes of one decoder for `n_reads` reads from the raw segments; asynchronous on `stream` */
cudaError_t launch_pack(const PackPlan& plan, long long n_reads, uint32_t* bases, uint16_t* nmask, uint32_t* quality, long long pitch,
                        int multiprocessor_count, cudaStream_t stream);

}   /* namespace phq */
#endif
