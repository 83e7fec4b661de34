/*  pheniqs_oracle.h — TEST INFRASTRUCTURE. CPU restatement of the reference's
    barcode classification path (PAMLD / MDD / naive / passthrough decoders).

    This is the checker the CUDA path is compared against. It is NOT part of the
    product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference legs may load it. The product (pheniqs_b200/) never links,
    imports or calls anything declared here.

    Parity status: PINNED. The restatement is checked (tests/test_oracle_*.py)
    against the reference's own golden vectors (test/BDGGG/valid/annotated.out,
    annotated.err; test/api/prior/valid/BDGGG_annotated_estimated.json, stored
    under tests/golden/) and, in the build container, against the reference's
    own classes compiled into oracle/_ref/libpheniqs_ref.so.

    All file:line citations are relative to the reference tree (/root/reference).
*/
#ifndef PHENIQS_ORACLE_H
#define PHENIQS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PHQO_PAMLD = 0, PHQO_MDD = 1, PHQO_NAIVE = 2, PHQO_PASSTHROUGH = 3 };      /* atom.h:348-355 */
enum { PHQO_SAMPLE = 0, PHQO_MOLECULAR = 1, PHQO_CELLULAR = 2 };                  /* classifier.h:28-33 (chain order differs: transcode.h:51-60) */

/* one Transform of a Rule: a token bound to an output segment (transform.h:36-121) */
typedef struct {
    int32_t input_segment_index;
    int32_t start;
    int32_t end;
    int32_t end_terminated;
    int32_t output_segment_index;
    int32_t reverse_complement;
} phqo_transform;

/* the fields a compiled decoder ontology carries (classifier.h:54-60, decoder.h:44-52, pamld.cpp:24-31, mdd.cpp:24-27) */
typedef struct {
    int32_t algorithm;
    int32_t topic;
    int32_t n_barcodes;                     /* codec cardinality, excluding undetermined */
    int32_t n_segments;                     /* "segment cardinality" */
    int32_t nucleotide_cardinality;         /* sum of segment lengths */
    int32_t n_transforms;
    const phqo_transform* transform;        /* [n_transforms], knit order */
    const int32_t* segment_length;          /* [n_segments] "barcode length" */
    const uint8_t* barcode;                 /* [n_barcodes][nucleotide_cardinality] BAM 4-bit codes, segments concatenated */
    const double* concentration;            /* [n_barcodes] compiled priors (sum to 1 - noise) */
    double noise;
    double confidence_threshold;
    double random_barcode_probability;
    int32_t high_quality_threshold;
    int32_t high_quality_distance_threshold;
    int32_t quality_masking_threshold;
    const int32_t* distance_tolerance;      /* [n_segments] (MDD) */
    int32_t multiplexing_classifier;
} phqo_decoder;

typedef struct phqo_job phqo_job;

/* decoders are given in chain order: sample, molecular[], cellular[] (transcode.h:51-60) */
phqo_job* phqo_create(int32_t n_decoders, const phqo_decoder* decoders);
void phqo_destroy(phqo_job* job);

/*  Decode reads [0, n_reads) sequentially on one thread, in order, carrying the
    Observation scratch from read to read exactly as one reference thread does.
    Segment s of read r is code[s][offset[s][r] .. offset[s][r+1]).
    Outputs (any may be NULL): per read per decoder index / distance / confidence
    ([n_reads][n_decoders]), final qcfail [n_reads], Read-level distance and
    confidence per topic ([n_reads][3], order sample, molecular, cellular) and
    channel index. Accumulates into the job's tables. */
void phqo_decode(phqo_job* job, int64_t n_reads, int32_t n_input_segments,
                 const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                 const uint8_t* qcfail_in,
                 int32_t* out_index, int32_t* out_distance, double* out_confidence,
                 uint8_t* out_qcfail, uint32_t* out_read_distance, double* out_read_confidence, int32_t* out_channel);

/*  Same over n_threads private copies of the job (contiguous read slices), merged
    with the reference's collect() afterwards (transcode.cpp:317-320). Returns the
    wall-clock seconds spent decoding. Per-read outputs as above (may be NULL). */
double phqo_decode_threaded(phqo_job* job, int32_t n_threads, int64_t n_reads, int32_t n_input_segments,
                 const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                 const uint8_t* qcfail_in,
                 int32_t* out_index, int32_t* out_distance, double* out_confidence, uint8_t* out_qcfail);

/*  Extracted observations (Rule::apply output) of decoder k for the reads decoded
    by the LAST phqo_decode call are not retained; this helper re-applies the rule
    of decoder k to a batch and returns the fixed-width concatenated observation
    [n_reads][nucleotide_cardinality] (code and quality), with the same short-token
    semantics (terminator + stale bytes) a sequential reference thread shows. */
void phqo_extract(const phqo_job* job, int32_t decoder, int64_t n_reads, int32_t n_input_segments,
                  const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                  uint8_t* out_code, uint8_t* out_quality, int32_t* out_length /* [n_reads][n_segments] observed lengths */);

/* raw accumulator tables of decoder k; row 0 undetermined, rows 1..NB codec order.
   u64 columns: count, pf_count, accumulated_distance, low_conditional_confidence_count,
   low_confidence_count, accumulated_pf_distance; f64 columns: accumulated_confidence,
   accumulated_pf_confidence (selector.h:34-41) */
void phqo_accumulators(const phqo_job* job, int32_t decoder, uint64_t* u64_table, double* f64_table);
void phqo_totals(const phqo_job* job, uint64_t* count, uint64_t* pf_count);
void phqo_reset(phqo_job* job);

/*  Prior estimation from accumulator tables (classifier.h:94-124 with pamld.h:40-48,
    decoder.h:77-83, selector.cpp:78-101). Pure function of the tables. */
void phqo_estimate_priors(int32_t n_barcodes, const uint64_t* u64_table, const double* f64_table,
                          double* estimated_noise, double* estimated_concentration /* [n_barcodes] */);

/* the Phred tables of phred.cpp:24-72: true_positive_quality[128] and the scalar constants */
void phqo_phred_tables(double* true_positive_quality /* [128] */, double* uniform_base_quality, double* phred_probability_base);

#ifdef __cplusplus
}
#endif
#endif
