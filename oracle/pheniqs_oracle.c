/*  pheniqs_oracle.c — TEST INFRASTRUCTURE (see pheniqs_oracle.h).

    Plain C restatement of the reference's barcode classification path, written
    from the reference's observable behaviour, one function per reference unit.
    Compiled with -ffp-contract=off so every double operation rounds exactly as
    the reference's x86-64 -O3 build does (no FMA contraction in the Kahan sums).
    All file:line citations are relative to /root/reference.
*/
#define _POSIX_C_SOURCE 200809L
#include "pheniqs_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SCRATCH_CAPACITY 512

/* ---------------------------------------------------------------- Phred scale
   phred.h:33-34, phred.cpp:24-72. The lookup is keyed q<<8 | expected<<4 | observed;
   rows for q = 0 are never written and stay 0.0 (static storage). */
typedef struct {
    double uniform_base_quality;            /* 10 * log10(4) */
    double phred_probability_base;          /* 10 ^ -0.1 */
    double false_positive_probability[128];
    double true_positive_probability[128];
    double true_positive_quality[128];
    double substitution_lookup[0x8000];
} phred_scale;

static phred_scale SCALE;
static pthread_once_t SCALE_ONCE = PTHREAD_ONCE_INIT;

static void assemble_scale(void) {
    memset(&SCALE, 0, sizeof(SCALE));
    SCALE.uniform_base_quality = 10.0 * log10(4);
    SCALE.phred_probability_base = pow(10.0, -0.1);
    for(int q = 1; q < 0x80; ++q) {
        SCALE.false_positive_probability[q] = pow(SCALE.phred_probability_base, q);
    }
    for(int q = 1; q < 0x80; ++q) {
        SCALE.true_positive_probability[q] = 1.0 - SCALE.false_positive_probability[q];
    }
    for(int q = 1; q < 0x80; ++q) {
        SCALE.true_positive_quality[q] = -10.0 * log10(SCALE.true_positive_probability[q]);
    }
    for(int q = 1; q < 0x80; ++q) {
        for(int e = 0; e < 0x10; ++e) {
            for(int o = 0; o < 0x10; ++o) {
                const int unambiguous_e = (e == 1 || e == 2 || e == 4 || e == 8);
                const int unambiguous_o = (o == 1 || o == 2 || o == 4 || o == 8);
                double value;
                if(unambiguous_e && unambiguous_o) {
                    value = (e == o) ? SCALE.true_positive_quality[q] : (double)q;
                } else {
                    value = SCALE.uniform_base_quality;
                }
                SCALE.substitution_lookup[q << 8 | e << 4 | o] = value;
            }
        }
    }
}
static inline double substitution_quality(uint8_t expected, uint8_t observed, uint8_t quality) {
    /* phred.h:58-60. quality >= 128 indexes past the reference's table (undefined there); 0.0 here. */
    if(quality >= 0x80) { return 0.0; }
    return SCALE.substitution_lookup[quality << 8 | expected << 4 | observed];
}
void phqo_phred_tables(double* true_positive_quality, double* uniform_base_quality, double* phred_probability_base) {
    pthread_once(&SCALE_ONCE, assemble_scale);
    if(true_positive_quality != NULL) { memcpy(true_positive_quality, SCALE.true_positive_quality, sizeof(double) * 128); }
    if(uniform_base_quality != NULL) { *uniform_base_quality = SCALE.uniform_base_quality; }
    if(phred_probability_base != NULL) { *phred_probability_base = SCALE.phred_probability_base; }
}

/* iupac.h:107-124 */
static const uint8_t BAM_REVERSE_COMPLEMENT[16] = { 0x0, 0x8, 0x4, 0xc, 0x2, 0xa, 0x6, 0xe, 0x1, 0x9, 0x5, 0xd, 0x3, 0xb, 0x7, 0xf };

/* ---------------------------------------------------------------- accumulators
   selector.h:32-60 (AccumulatingOption), fields in declaration order */
typedef struct {
    uint64_t count;
    uint64_t pf_count;
    uint64_t accumulated_distance;
    double accumulated_confidence;
    uint64_t low_conditional_confidence_count;
    uint64_t low_confidence_count;
    uint64_t accumulated_pf_distance;
    double accumulated_pf_confidence;
} option;

static void option_collect(option* to, const option* from) {     /* selector.cpp:68-77 */
    to->count += from->count;
    to->pf_count += from->pf_count;
    to->accumulated_distance += from->accumulated_distance;
    to->accumulated_confidence += from->accumulated_confidence;
    to->low_conditional_confidence_count += from->low_conditional_confidence_count;
    to->low_confidence_count += from->low_confidence_count;
    to->accumulated_pf_distance += from->accumulated_pf_distance;
    to->accumulated_pf_confidence += from->accumulated_pf_confidence;
}

/* ---------------------------------------------------------------- decoder state */
typedef struct {
    uint8_t code[SCRATCH_CAPACITY];
    uint8_t quality[SCRATCH_CAPACITY];
    int32_t length;
} observed_sequence;                        /* sequence.h:264-300; never re-zeroed between reads */

typedef struct {
    phqo_decoder spec;                      /* deep copy */
    phqo_transform* transform;
    int32_t* segment_length;
    int32_t* segment_offset;
    uint8_t* barcode;
    double* concentration;
    int32_t* distance_tolerance;
    double adjusted_noise_probability;      /* pamld.cpp:29 */

    observed_sequence* observation;         /* [n_segments] */
    option* tag;                            /* [n_barcodes + 1]; row 0 = unclassified */

    /* per-decoder state that the reference keeps across reads */
    int32_t decoded;                        /* 0 = unclassified, else 1-based barcode row (classifier.h:47,62) */
    int32_t edit_distance;                  /* decoder.h:34 */
    int32_t high_quality_edit_distance;     /* decoder.h:35 */
    double conditional_decoding_probability;/* pamld.h:34, not reset per read */
    double decoding_confidence;             /* pamld.h:35 */
} decoder;

struct phqo_job {
    int32_t n_decoders;
    decoder* chain;
    uint64_t count;                         /* transcode.h:44-45 */
    uint64_t pf_count;
};

/* the part of Read the decoders write (read.h:152-158, 187-199) */
typedef struct {
    int qcfail;
    int32_t channel_index;
    uint32_t distance[3];
    double confidence[3];
} read_state;

static void* xcalloc(size_t n, size_t size) {
    void* p = calloc(n ? n : 1, size);
    if(p == NULL) { abort(); }
    return p;
}

static void decoder_init(decoder* d, const phqo_decoder* spec) {
    memset(d, 0, sizeof(*d));
    d->spec = *spec;
    const int32_t ns = spec->n_segments;
    const int32_t nb = spec->n_barcodes;
    const int32_t nc = spec->nucleotide_cardinality;
    d->transform = xcalloc(spec->n_transforms, sizeof(phqo_transform));
    memcpy(d->transform, spec->transform, sizeof(phqo_transform) * spec->n_transforms);
    d->segment_length = xcalloc(ns, sizeof(int32_t));
    d->segment_offset = xcalloc(ns + 1, sizeof(int32_t));
    for(int32_t i = 0; i < ns; ++i) {
        d->segment_length[i] = spec->segment_length ? spec->segment_length[i] : 0;
        d->segment_offset[i + 1] = d->segment_offset[i] + d->segment_length[i];
    }
    d->barcode = xcalloc((size_t)nb * nc, 1);
    if(nb > 0 && nc > 0) { memcpy(d->barcode, spec->barcode, (size_t)nb * nc); }
    d->concentration = xcalloc(nb, sizeof(double));
    if(nb > 0 && spec->concentration) { memcpy(d->concentration, spec->concentration, sizeof(double) * nb); }
    d->distance_tolerance = xcalloc(ns, sizeof(int32_t));
    if(spec->distance_tolerance) { memcpy(d->distance_tolerance, spec->distance_tolerance, sizeof(int32_t) * ns); }
    d->adjusted_noise_probability = spec->noise * spec->random_barcode_probability;
    d->observation = xcalloc(ns, sizeof(observed_sequence));
    d->tag = xcalloc((size_t)nb + 1, sizeof(option));
    d->spec.transform = d->transform;
    d->spec.segment_length = d->segment_length;
    d->spec.barcode = d->barcode;
    d->spec.concentration = d->concentration;
    d->spec.distance_tolerance = d->distance_tolerance;
}
static void decoder_free(decoder* d) {
    free(d->transform); free(d->segment_length); free(d->segment_offset); free(d->barcode);
    free(d->concentration); free(d->distance_tolerance); free(d->observation); free(d->tag);
}

phqo_job* phqo_create(int32_t n_decoders, const phqo_decoder* decoders) {
    pthread_once(&SCALE_ONCE, assemble_scale);
    phqo_job* job = xcalloc(1, sizeof(phqo_job));
    job->n_decoders = n_decoders;
    job->chain = xcalloc(n_decoders, sizeof(decoder));
    for(int32_t k = 0; k < n_decoders; ++k) { decoder_init(&job->chain[k], &decoders[k]); }
    return job;
}
void phqo_destroy(phqo_job* job) {
    if(job == NULL) { return; }
    for(int32_t k = 0; k < job->n_decoders; ++k) { decoder_free(&job->chain[k]); }
    free(job->chain);
    free(job);
}
static phqo_job* job_clone(const phqo_job* job) {
    phqo_decoder* specs = xcalloc(job->n_decoders, sizeof(phqo_decoder));
    for(int32_t k = 0; k < job->n_decoders; ++k) { specs[k] = job->chain[k].spec; }
    phqo_job* copy = phqo_create(job->n_decoders, specs);
    free(specs);
    return copy;
}
void phqo_reset(phqo_job* job) {
    job->count = 0;
    job->pf_count = 0;
    for(int32_t k = 0; k < job->n_decoders; ++k) {
        memset(job->chain[k].tag, 0, sizeof(option) * ((size_t)job->chain[k].spec.n_barcodes + 1));
    }
}

/* ---------------------------------------------------------------- token slicing
   transform.h:65-80 (python-slice like, with the reference's clamping) */
static int32_t absolute_end(const phqo_transform* t, int32_t length) {
    if(t->end_terminated) {
        if(t->end < 0) {
            int32_t value = length + t->end;
            return value < 0 ? 0 : value;
        } else { return t->end > length ? length : t->end; }
    } else { return length; }
}
static int32_t absolute_start(const phqo_transform* t, int32_t length) {
    if(t->start < 0) {
        int32_t value = length + t->start;
        return value < 0 ? 0 : value;
    } else { return t->start > length ? 0 : t->start; }
}

typedef struct {
    int32_t n_segments;
    const uint8_t* const* code;
    const uint8_t* const* quality;
    const int64_t* const* offset;
} batch;

/* Observation::clear (sequence.h:296-300) followed by Rule::apply (transform.h:142-169) */
static void rule_apply(decoder* d, const batch* b, int64_t r) {
    for(int32_t i = 0; i < d->spec.n_segments; ++i) {
        d->observation[i].length = 0;
        d->observation[i].code[0] = 0;
        d->observation[i].quality[0] = 0;
    }
    for(int32_t k = 0; k < d->spec.n_transforms; ++k) {
        const phqo_transform* t = &d->transform[k];
        const int64_t from = b->offset[t->input_segment_index][r];
        const int32_t from_length = (int32_t)(b->offset[t->input_segment_index][r + 1] - from);
        const uint8_t* from_code = b->code[t->input_segment_index] + from;
        const uint8_t* from_quality = b->quality[t->input_segment_index] + from;
        observed_sequence* to = &d->observation[t->output_segment_index];
        const int32_t start = absolute_start(t, from_length);
        const int32_t end = absolute_end(t, from_length);
        const int32_t size = end - start;
        if(size > 0 && to->length + size + 1 < SCRATCH_CAPACITY) {
            if(!t->reverse_complement) {
                memcpy(to->code + to->length, from_code + start, size);
                memcpy(to->quality + to->length, from_quality + start, size);
            } else {
                for(int32_t i = 0; i < size; ++i) {
                    to->code[to->length + i] = BAM_REVERSE_COMPLEMENT[from_code[end - i - 1] & 0xf];
                    to->quality[to->length + i] = from_quality[end - i - 1];
                }
            }
            to->length += size;
            to->code[to->length] = 0;
            to->quality[to->length] = 0;
        }
    }
}

/* ---------------------------------------------------------------- P(r|b)
   Barcode::compensated_decoding_probability (barcode.h:131-164). Iterates the EXPECTED
   length of every segment, so a short observation contributes its terminator byte and
   then whatever the scratch still holds from earlier reads. */
static void compensated_decoding_probability(const decoder* d, int32_t b, double* probability, int32_t* distance, int32_t* high_quality_distance) {
    double y = 0, t = 0, sigma_q = 0, compensation = 0;
    *distance = 0;
    *high_quality_distance = 0;
    const uint8_t* expected = d->barcode + (size_t)b * d->spec.nucleotide_cardinality;
    for(int32_t i = 0; i < d->spec.n_segments; ++i) {
        const observed_sequence* observed = &d->observation[i];
        const uint8_t* e = expected + d->segment_offset[i];
        for(int32_t j = 0; j < d->segment_length[i]; ++j) {
            y = substitution_quality(e[j], observed->code[j], observed->quality[j]) - compensation;
            t = sigma_q + y;
            compensation = (t - sigma_q) - y;
            sigma_q = t;
            if(observed->code[j] != e[j]) {
                ++(*distance);
                if(observed->quality[j] >= d->spec.high_quality_threshold) {
                    ++(*high_quality_distance);
                }
            }
        }
    }
    *probability = pow(SCALE.phred_probability_base, sigma_q);
}

/* Decoder::classify (decoder.h:68-76) then Classifier::classify (classifier.h:78-86) */
static void base_classify(decoder* d, read_state* output) {
    option* decoded = &d->tag[d->decoded];
    if(d->decoded > 0 && d->edit_distance) {
        decoded->accumulated_distance += (uint64_t)d->edit_distance;
        if(!output->qcfail) {
            decoded->accumulated_pf_distance += (uint64_t)d->edit_distance;
        }
    }
    ++decoded->count;
    if(!output->qcfail) {
        ++decoded->pf_count;
    }
    if(d->spec.multiplexing_classifier) {
        output->channel_index = d->decoded;
    }
}

/* read.h:279-285 and siblings */
static void update_confidence(double* field, double confidence) {
    if(*field == 1) { *field = confidence; } else { *field *= confidence; }
}

/* PamlDecoder::classify (pamld.cpp:37-123) + the three routing subclasses (pamld.cpp:133-180) */
static void pamld_classify(decoder* d, const batch* b, int64_t r, read_state* output) {
    rule_apply(d, b, r);
    double p = 0, y = 0, t = 0;
    int32_t distance = 0, hqd = 0;
    double sigma_p = 0, compensation = 0, conditional_probability = 0;
    double adjusted_conditional_decoding_probability = 0;

    for(int32_t i = 0; i < d->spec.n_barcodes; ++i) {
        compensated_decoding_probability(d, i, &conditional_probability, &distance, &hqd);
        p = conditional_probability * d->concentration[i];
        y = p - compensation;
        t = sigma_p + y;
        compensation = (t - sigma_p) - y;
        sigma_p = t;
        if(p > adjusted_conditional_decoding_probability) {
            d->decoded = i + 1;
            d->edit_distance = distance;
            d->high_quality_edit_distance = hqd;
            adjusted_conditional_decoding_probability = p;
            d->conditional_decoding_probability = conditional_probability;
        }
    }
    y = d->adjusted_noise_probability - compensation;
    t = sigma_p + y;
    compensation = (t - sigma_p) - y;
    sigma_p = t;

    d->decoding_confidence = adjusted_conditional_decoding_probability / sigma_p;

    if(d->conditional_decoding_probability > d->spec.random_barcode_probability) {
        if(d->decoding_confidence > d->spec.confidence_threshold) {
            d->tag[d->decoded].accumulated_confidence += d->decoding_confidence;
            if(d->spec.high_quality_distance_threshold > 0 && d->high_quality_edit_distance >= d->spec.high_quality_distance_threshold) {
                output->qcfail = 1;
            }
            if(!output->qcfail) {
                d->tag[d->decoded].accumulated_pf_confidence += d->decoding_confidence;
            }
        } else {
            ++d->tag[d->decoded].low_confidence_count;
            output->qcfail = 1;
        }
    } else {
        ++d->tag[d->decoded].low_conditional_confidence_count;
        output->qcfail = 1;
        d->decoded = 0;
        d->edit_distance = 0;
        d->high_quality_edit_distance = 0;
        d->decoding_confidence = 0;
    }
    base_classify(d, output);

    switch(d->spec.topic) {
        case PHQO_SAMPLE:
            output->distance[0] += (uint32_t)d->edit_distance;
            update_confidence(&output->confidence[0], d->decoding_confidence);
            break;
        case PHQO_CELLULAR:
            if(d->decoded > 0) {
                update_confidence(&output->confidence[2], d->decoding_confidence);
                output->distance[2] += (uint32_t)d->edit_distance;
            } else {
                output->confidence[2] = 0;
                output->distance[2] = 0;
            }
            break;
        case PHQO_MOLECULAR:
            if(d->decoded > 0) {
                update_confidence(&output->confidence[1], d->decoding_confidence);
                output->distance[1] += (uint32_t)d->edit_distance;
            } else {
                output->confidence[1] = 0;
                output->distance[1] += 0;       /* Read::set_molecular_distance adds (read.h:319-321) */
            }
            break;
    }
}

/* MdDecoder::classify (mdd.cpp:37-86) + routing subclasses (mdd.cpp:96-138) */
static void mdd_classify(decoder* d, const batch* b, int64_t r, read_state* output) {
    rule_apply(d, b, r);
    d->decoded = 0;
    d->edit_distance = 0;

    /* exact match on the concatenated BAM code string (mdd.cpp:44-46, sequence.h:483-493, barcode.h:46-56) */
    int32_t observed_total = 0;
    for(int32_t i = 0; i < d->spec.n_segments; ++i) { observed_total += d->observation[i].length; }
    int32_t exact = -1;
    if(observed_total == d->spec.nucleotide_cardinality) {
        uint8_t key[SCRATCH_CAPACITY];
        int32_t at = 0;
        for(int32_t i = 0; i < d->spec.n_segments; ++i) {
            memcpy(key + at, d->observation[i].code, d->observation[i].length);
            at += d->observation[i].length;
        }
        for(int32_t i = 0; i < d->spec.n_barcodes; ++i) {
            if(memcmp(key, d->barcode + (size_t)i * d->spec.nucleotide_cardinality, observed_total) == 0) {
                exact = i;
                break;
            }
        }
    }
    if(exact >= 0) {
        d->decoded = exact + 1;
    } else {
        for(int32_t k = 0; k < d->spec.n_barcodes; ++k) {
            const uint8_t* expected = d->barcode + (size_t)k * d->spec.nucleotide_cardinality;
            int32_t distance = 0;
            int successful = 1;
            for(int32_t i = 0; i < d->spec.n_segments; ++i) {
                const observed_sequence* observed = &d->observation[i];
                const uint8_t* e = expected + d->segment_offset[i];
                int32_t error = 0;
                if(d->spec.quality_masking_threshold > 0) {
                    /* ObservedSequence::masked_distance_from (sequence.h:321-332): over the OBSERVED length */
                    for(int32_t j = 0; j < observed->length; ++j) {
                        if(observed->quality[j] < d->spec.quality_masking_threshold) { ++error; }
                        else if(observed->code[j] != e[j]) { ++error; }
                    }
                } else {
                    /* Sequence::distance_from (sequence.h:90-98) */
                    for(int32_t j = 0; j < observed->length; ++j) {
                        if(observed->code[j] != e[j]) { ++error; }
                    }
                }
                if(error > d->distance_tolerance[i]) { successful = 0; break; }
                else { distance += error; }
            }
            if(successful) {
                d->edit_distance = distance;
                d->decoded = k + 1;
                break;
            }
        }
    }
    if(d->decoded == 0) { output->qcfail = 1; }
    base_classify(d, output);

    switch(d->spec.topic) {
        case PHQO_SAMPLE:
            output->distance[0] += (uint32_t)d->edit_distance;
            break;
        case PHQO_CELLULAR:
            if(d->decoded > 0) { output->distance[2] += (uint32_t)d->edit_distance; }
            else { output->distance[2] = 0; }
            break;
        case PHQO_MOLECULAR:
            if(d->decoded > 0) { output->distance[1] += (uint32_t)d->edit_distance; }
            else { output->distance[2] = 0; }   /* MdMolecularDecoder clears the CELLULAR distance (mdd.cpp:136) */
            break;
    }
}

/* NaiveMolecularDecoder::classify (naive.h:40-45): rule + base bookkeeping on the unclassified row */
static void naive_classify(decoder* d, const batch* b, int64_t r, read_state* output) {
    rule_apply(d, b, r);
    base_classify(d, output);
}
/* Classifier< Barcode >::classify used directly for "passthrough" (transcode.cpp:78-80, classifier.h:78-86) */
static void passthrough_classify(decoder* d, read_state* output) {
    option* decoded = &d->tag[d->decoded];
    ++decoded->count;
    if(!output->qcfail) { ++decoded->pf_count; }
    if(d->spec.multiplexing_classifier) { output->channel_index = d->decoded; }
}

typedef struct {
    int32_t* index; int32_t* distance; double* confidence;
    uint8_t* qcfail; uint32_t* read_distance; double* read_confidence; int32_t* channel;
} outputs;

/* TranscodingThread::run body (transcode.h:202-225) with TranscodingDecoder::classify (transcode.h:51-65) */
static void run_slice(phqo_job* job, const batch* b, const uint8_t* qcfail_in, int64_t begin, int64_t end, const outputs* out) {
    const int32_t nd = job->n_decoders;
    for(int64_t r = begin; r < end; ++r) {
        read_state output;
        output.qcfail = (qcfail_in != NULL && qcfail_in[r]) ? 1 : 0;
        output.channel_index = 0;
        for(int i = 0; i < 3; ++i) { output.distance[i] = 0; output.confidence[i] = 1; }   /* Read::clear, read.h:166-186 */

        for(int32_t k = 0; k < nd; ++k) {
            decoder* d = &job->chain[k];
            switch(d->spec.algorithm) {
                case PHQO_PAMLD:        pamld_classify(d, b, r, &output); break;
                case PHQO_MDD:          mdd_classify(d, b, r, &output); break;
                case PHQO_NAIVE:        naive_classify(d, b, r, &output); break;
                default:                passthrough_classify(d, &output); break;
            }
            if(out->index != NULL) {
                out->index[r * nd + k] = d->decoded;
                out->distance[r * nd + k] = (d->spec.algorithm == PHQO_PAMLD || d->spec.algorithm == PHQO_MDD) ? d->edit_distance : 0;
                out->confidence[r * nd + k] = (d->spec.algorithm == PHQO_PAMLD) ? d->decoding_confidence : 0;
            }
        }
        ++job->count;
        if(!output.qcfail) { ++job->pf_count; }
        if(out->qcfail != NULL) { out->qcfail[r] = (uint8_t)output.qcfail; }
        if(out->read_distance != NULL) { for(int i = 0; i < 3; ++i) { out->read_distance[r * 3 + i] = output.distance[i]; } }
        if(out->read_confidence != NULL) { for(int i = 0; i < 3; ++i) { out->read_confidence[r * 3 + i] = output.confidence[i]; } }
        if(out->channel != NULL) { out->channel[r] = output.channel_index; }
    }
}

void phqo_decode(phqo_job* job, int64_t n_reads, int32_t n_input_segments,
                 const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                 const uint8_t* qcfail_in,
                 int32_t* out_index, int32_t* out_distance, double* out_confidence,
                 uint8_t* out_qcfail, uint32_t* out_read_distance, double* out_read_confidence, int32_t* out_channel) {
    batch b = { n_input_segments, code, quality, offset };
    outputs out = { out_index, out_distance, out_confidence, out_qcfail, out_read_distance, out_read_confidence, out_channel };
    run_slice(job, &b, qcfail_in, 0, n_reads, &out);
}

typedef struct {
    phqo_job* job; const batch* b; const uint8_t* qcfail_in; int64_t begin, end; const outputs* out;
} slice_argument;
static void* slice_main(void* p) {
    slice_argument* a = p;
    run_slice(a->job, a->b, a->qcfail_in, a->begin, a->end, a->out);
    return NULL;
}
static void job_collect(phqo_job* to, const phqo_job* from) {      /* transcode.cpp:162-179, classifier.h:87-93 */
    to->count += from->count;
    to->pf_count += from->pf_count;
    for(int32_t k = 0; k < to->n_decoders; ++k) {
        for(int32_t i = 0; i <= to->chain[k].spec.n_barcodes; ++i) {
            option_collect(&to->chain[k].tag[i], &from->chain[k].tag[i]);
        }
    }
}
double phqo_decode_threaded(phqo_job* job, int32_t n_threads, int64_t n_reads, int32_t n_input_segments,
                 const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                 const uint8_t* qcfail_in,
                 int32_t* out_index, int32_t* out_distance, double* out_confidence, uint8_t* out_qcfail) {
    if(n_threads < 1) { n_threads = 1; }
    batch b = { n_input_segments, code, quality, offset };
    outputs out = { out_index, out_distance, out_confidence, out_qcfail, NULL, NULL, NULL };
    phqo_job** part = xcalloc(n_threads, sizeof(phqo_job*));
    slice_argument* argument = xcalloc(n_threads, sizeof(slice_argument));
    pthread_t* thread = xcalloc(n_threads, sizeof(pthread_t));
    for(int32_t t = 0; t < n_threads; ++t) { part[t] = job_clone(job); }

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for(int32_t t = 0; t < n_threads; ++t) {
        slice_argument a = { part[t], &b, qcfail_in, n_reads * t / n_threads, n_reads * (t + 1) / n_threads, &out };
        argument[t] = a;
        pthread_create(&thread[t], NULL, slice_main, &argument[t]);
    }
    for(int32_t t = 0; t < n_threads; ++t) { pthread_join(thread[t], NULL); }
    clock_gettime(CLOCK_MONOTONIC, &t1);

    for(int32_t t = 0; t < n_threads; ++t) { job_collect(job, part[t]); phqo_destroy(part[t]); }
    free(part); free(argument); free(thread);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void phqo_extract(const phqo_job* job, int32_t k, int64_t n_reads, int32_t n_input_segments,
                  const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                  uint8_t* out_code, uint8_t* out_quality, int32_t* out_length) {
    phqo_decoder spec = job->chain[k].spec;
    decoder d;
    decoder_init(&d, &spec);
    batch b = { n_input_segments, code, quality, offset };
    const int32_t nc = spec.nucleotide_cardinality;
    for(int64_t r = 0; r < n_reads; ++r) {
        rule_apply(&d, &b, r);
        for(int32_t i = 0; i < spec.n_segments; ++i) {
            /* what a decoder iterating the EXPECTED length sees, stale tail included */
            memcpy(out_code + r * nc + d.segment_offset[i], d.observation[i].code, d.segment_length[i]);
            memcpy(out_quality + r * nc + d.segment_offset[i], d.observation[i].quality, d.segment_length[i]);
            if(out_length != NULL) { out_length[r * spec.n_segments + i] = d.observation[i].length; }
        }
    }
    decoder_free(&d);
}

void phqo_accumulators(const phqo_job* job, int32_t k, uint64_t* u, double* f) {
    const decoder* d = &job->chain[k];
    for(int32_t i = 0; i <= d->spec.n_barcodes; ++i) {
        const option* o = &d->tag[i];
        u[i * 6 + 0] = o->count;
        u[i * 6 + 1] = o->pf_count;
        u[i * 6 + 2] = o->accumulated_distance;
        u[i * 6 + 3] = o->low_conditional_confidence_count;
        u[i * 6 + 4] = o->low_confidence_count;
        u[i * 6 + 5] = o->accumulated_pf_distance;
        f[i * 2 + 0] = o->accumulated_confidence;
        f[i * 2 + 1] = o->accumulated_pf_confidence;
    }
}
void phqo_totals(const phqo_job* job, uint64_t* count, uint64_t* pf_count) {
    *count = job->count;
    *pf_count = job->pf_count;
}

/*  PamlDecoder::finalize (pamld.h:40-48) -> Decoder::finalize (decoder.h:77-83) ->
    Classifier::finalize (classifier.h:94-124) -> AccumulatingOption::finalize (selector.cpp:78-101).
    Only the quantities the prior estimate depends on are produced. */
void phqo_estimate_priors(int32_t n_barcodes, const uint64_t* u, const double* f, double* estimated_noise, double* estimated_concentration) {
    (void)f;
    uint64_t classified_count = 0, pf_classified_count = 0;
    uint64_t low_conditional_confidence_count = 0, low_confidence_count = 0;
    for(int32_t i = 1; i <= n_barcodes; ++i) {
        low_conditional_confidence_count += u[i * 6 + 3];
        low_confidence_count += u[i * 6 + 4];
    }
    for(int32_t i = 1; i <= n_barcodes; ++i) {
        classified_count += u[i * 6 + 0];
        pf_classified_count += u[i * 6 + 1];
    }
    const uint64_t count = classified_count + u[0];

    double estimated_noise_count = (double)low_conditional_confidence_count;
    double confident_noise_ratio = estimated_noise_count / (estimated_noise_count + pf_classified_count);
    if(low_confidence_count > 0) {
        estimated_noise_count += (double)low_confidence_count * confident_noise_ratio;
    }
    const double estimated_noise_prior = estimated_noise_count / (double)count;
    const double estimated_not_noise_prior = 1.0 - estimated_noise_prior;
    for(int32_t i = 1; i <= n_barcodes; ++i) {
        double pf_pooled_classified_fraction = 0;
        const uint64_t pf_count = u[i * 6 + 1];
        if(pf_count > 0 && pf_classified_count > 0) {
            pf_pooled_classified_fraction = (double)pf_count / (double)pf_classified_count;
        }
        estimated_concentration[i - 1] = estimated_not_noise_prior * pf_pooled_classified_fraction;
    }
    *estimated_noise = estimated_noise_prior;
}
