/*  pheniqs_b200.hpp — header-only C++ host wrapper over the C ABI (pheniqs_b200.h).

    Mirrors the shape of the reference's per-thread decoder object so the call site in
    TranscodingThread::run (transcode.h:202-225) changes from "classify one Read" to
    "classify one batch of Reads":

        reference                                   here
        ---------                                   ----
        TranscodingDecoder(const Value& ontology)   phq::BatchDecoder(compiled_json, device)
        classify(const Read&, Read&)                classify(batch)            (transcode.h:51-65)
        collect(const TranscodingDecoder&)          collect(nccl_comm, stream): one all-reduce (transcode.cpp:162-179)
        finalize()                                  estimate_priors(k)         (classifier.h:94-124)
        Transcode::finalize report                  report()                   (transcode.cpp:1811-1863)
        Error subclasses with ErrorCode             phq::Error subclasses with the same codes (error.h:32-136)

    No CUDA or torch types appear; link with -lpheniqs_b200.
*/
#ifndef PHENIQS_B200_HPP
#define PHENIQS_B200_HPP

#include "pheniqs_b200.h"

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace phq {

/* error.h:46-136: one exception class per ErrorCode the path can raise */
class Error : public std::runtime_error {
    public:
        const int code;
        Error(int code, const std::string& message) : std::runtime_error(message), code(code) {}
};
class InternalError : public Error { public: explicit InternalError(const std::string& m) : Error(PHQ_INTERNAL_ERROR, m) {} };
class ConfigurationError : public Error { public: explicit ConfigurationError(const std::string& m) : Error(PHQ_CONFIGURATION_ERROR, m) {} };
class OutOfMemoryError : public Error { public: explicit OutOfMemoryError(const std::string& m) : Error(PHQ_OUT_OF_MEMORY_ERROR, m) {} };
class SequenceError : public Error { public: explicit SequenceError(const std::string& m) : Error(PHQ_SEQUENCE_ERROR, m) {} };
class OverflowError : public Error { public: explicit OverflowError(const std::string& m) : Error(PHQ_OVERFLOW_ERROR, m) {} };

/* status first, then the message it left behind: never both as arguments of one call (their evaluation order is unspecified) */
inline void raise(int status, const char* message) {
    const std::string text(message != NULL ? message : "");
    switch(status) {
        case PHQ_OK: return;
        case PHQ_CONFIGURATION_ERROR: throw ConfigurationError(text);
        case PHQ_OUT_OF_MEMORY_ERROR: throw OutOfMemoryError(text);
        case PHQ_SEQUENCE_ERROR: throw SequenceError(text);
        case PHQ_OVERFLOW_ERROR: throw OverflowError(text);
        case PHQ_INTERNAL_ERROR: throw InternalError(text);
        default: throw Error(status, text);
    }
}

/* Transcode::compile for the decoder sections of a job */
inline std::string compile_job(const std::string& job_json) {
    char* out(NULL);
    const int status(phq_compile_job(job_json.c_str(), &out));
    raise(status, phq_last_global_error());
    std::string compiled(out);
    phq_free(out);
    return compiled;
}

/* the job file with its imports merged underneath (Job::load_instruction_with_import, job.cpp:160-224) */
inline std::string load_job(const std::string& path) {
    char* out(NULL);
    const int status(phq_load_job(path.c_str(), &out));
    raise(status, phq_last_global_error());
    std::string job(out);
    phq_free(out);
    return job;
}

/* the prior adjusted job (tool/pheniqs-prior-api.py:39-56, classifier.h:125-160) */
inline std::string adjust_job(const std::string& job_json, const std::string& report_json, int precision = 15) {
    char* out(NULL);
    const int status(phq_adjust_job(job_json.c_str(), report_json.c_str(), precision, &out));
    raise(status, phq_last_global_error());
    std::string adjusted(out);
    phq_free(out);
    return adjusted;
}

/* host planes of one decoder for a batch, owned by the caller (pinned when `pinned`) */
class TileBuffer {
    public:
        TileBuffer() : n_reads_(0), pinned_(false) { memset(&tile_, 0, sizeof(tile_)); }
        TileBuffer(const TileBuffer&) = delete;
        void operator=(const TileBuffer&) = delete;
        ~TileBuffer() { release(); }
        void allocate(const phq_decoder_info& info, int64_t n_reads, bool pinned) {
            release();
            n_reads_ = n_reads;
            pinned_ = pinned;
            const size_t pitch(static_cast< size_t >(n_reads > 0 ? n_reads : 1));
            tile_.pitch = static_cast< int64_t >(pitch);
            tile_.bases = static_cast< const uint32_t* >(get(pitch * info.word_cardinality * sizeof(uint32_t)));
            tile_.nmask = static_cast< const uint16_t* >(get(pitch * info.word_cardinality * sizeof(uint16_t)));
            tile_.quality = static_cast< const uint32_t* >(get(pitch * info.quality_word_cardinality * sizeof(uint32_t)));
        }
        const phq_tile& tile() const { return tile_; }
    private:
        phq_tile tile_;
        int64_t n_reads_;
        bool pinned_;
        void* get(size_t bytes) {
            void* p(NULL);
            if(pinned_) { const int status(phq_host_alloc(&p, bytes)); raise(status, phq_last_global_error()); }
            else { p = ::operator new(bytes ? bytes : 1); }
            return p;
        }
        void drop(const void* p) {
            if(p == NULL) { return; }
            if(pinned_) { phq_host_free(const_cast< void* >(p)); } else { ::operator delete(const_cast< void* >(p)); }
        }
        void release() {
            drop(tile_.bases); drop(tile_.nmask); drop(tile_.quality);
            tile_.bases = NULL; tile_.nmask = NULL; tile_.quality = NULL;
        }
};

/* one GPU's decoder chain: what a TranscodingThread's TranscodingDecoder is in the reference */
class BatchDecoder {
    public:
        BatchDecoder(const std::string& compiled_job_json, int device) : handle_(NULL) {
            const int status(phq_create(compiled_job_json.c_str(), device, &handle_));
            raise(status, phq_last_global_error());
            const int n(phq_decoder_count(handle_));
            info_.resize(static_cast< size_t >(n));
            for(int k(0); k < n; ++k) { check(phq_decoder_describe(handle_, k, &info_[k])); }
        }
        BatchDecoder(const BatchDecoder&) = delete;
        void operator=(const BatchDecoder&) = delete;
        ~BatchDecoder() { phq_destroy(handle_); }

        size_t decoder_cardinality() const { return info_.size(); }
        const phq_decoder_info& info(size_t k) const { return info_[k]; }

        /* Rule::apply + packing (transform.h:142-169) for reads held one code byte and one Phred byte per base;
           tiles[k].quality_bits chooses the quality form on the way in and reports it on the way out */
        void pack(int64_t n_reads, int32_t n_input_segments, const uint8_t* const* code, const uint8_t* const* quality,
                  const int64_t* const* offset, std::vector< phq_tile >& tiles) {
            check(phq_pack(handle_, n_reads, n_input_segments, code, quality, offset, tiles.data()));
        }
        /* the 8-byte records of the reference's output (index, distance, qcfail, float error probability) */
        void classify_compact(int64_t n_reads, const std::vector< phq_tile >& tiles, const uint8_t* qcfail_in,
                              const std::vector< phq_compact_result* >& results) {
            check(phq_decode_batch_compact(handle_, n_reads, tiles.data(), qcfail_in, results.data()));
        }
        /* TranscodingDecoder::classify (transcode.h:51-65) for a batch; host buffers */
        void classify(int64_t n_reads, const std::vector< phq_tile >& tiles, const uint8_t* qcfail_in,
                      const std::vector< phq_result* >& results, uint8_t* qcfail_out) {
            check(phq_decode_batch(handle_, n_reads, tiles.data(), qcfail_in, results.data(), qcfail_out));
        }
        /* FASTQ bytes of the barcode bearing segments in; the BAM auxiliary block of every read out (Read::flush +
           Auxiliary::encode, read.h:187-237, auxiliary.cpp:320-361): aux[r * stride .. + aux_length[r]) */
        int32_t tag_record_bytes() {
            int32_t bytes(0);
            check(phq_tag_record_bytes(handle_, &bytes));
            return bytes;
        }
        void classify_raw_tags(int64_t n_reads, const std::vector< phq_raw_segment >& segments, int32_t phred_offset, const uint8_t* qcfail_in,
                               uint8_t* aux, int32_t aux_stride, int32_t* aux_length, uint8_t* qcfail_out) {
            check(phq_decode_batch_raw_tags(handle_, n_reads, static_cast< int32_t >(segments.size()), segments.data(), phred_offset, qcfail_in,
                                            aux, aux_stride, aux_length, qcfail_out, NULL));
        }
        /* the reference's own decoded form of the barcode bearing segments in (Segment::code BAM bytes and Phred bytes,
           sequence.h:264-300): slicing, reverse complement and packing happen on the device */
        void classify_bam(int64_t n_reads, const std::vector< phq_raw_segment >& segments, const uint8_t* qcfail_in,
                          const std::vector< phq_result* >& results, uint8_t* qcfail_out) {
            check(phq_decode_batch_bam(handle_, n_reads, static_cast< int32_t >(segments.size()), segments.data(), qcfail_in, results.data(), qcfail_out));
        }
        /* Classifier::collect across GPUs: in-place all-reduce of the accumulator planes over the host's ncclComm_t */
        void collect(void* nccl_comm, void* stream) { check(phq_collect(handle_, nccl_comm, stream)); }
        /* device pointers, asynchronous on `stream` */
        void classify_device(int64_t n_reads, const std::vector< phq_tile >& tiles, uint8_t* qcfail, const std::vector< phq_result* >& results, void* stream) {
            check(phq_decode_batch_device(handle_, n_reads, tiles.data(), qcfail, results.data(), stream));
        }
        /* AccumulatingOption tables of decoder k (selector.h:32-60) */
        void accumulators(size_t k, std::vector< uint64_t >& u64_table, std::vector< double >& f64_table) {
            const size_t rows(static_cast< size_t >(info_[k].barcode_cardinality) + 1);
            u64_table.assign(rows * 6, 0);
            f64_table.assign(rows * 2, 0);
            check(phq_accumulators(handle_, static_cast< int >(k), u64_table.data(), f64_table.data()));
        }
        void totals(uint64_t& count, uint64_t& pf_count) { check(phq_totals(handle_, &count, &pf_count)); }
        /* the buffer to all-reduce(sum) across GPUs in place of collect() (transcode.cpp:162-179) */
        void accumulator_buffer(void*& device_pointer, int64_t& n_u64, int64_t& n_f64) { check(phq_accumulator_buffer(handle_, &device_pointer, &n_u64, &n_f64)); }
        void reset() { check(phq_reset_accumulators(handle_)); }
        /* Classifier::finalize (classifier.h:94-124) */
        double estimate_priors(size_t k, std::vector< double >& concentration) {
            double noise(0);
            concentration.assign(static_cast< size_t >(info_[k].barcode_cardinality), 0);
            check(phq_estimate_priors(handle_, static_cast< int >(k), &noise, concentration.data()));
            return noise;
        }
        /* Classifier::adjust_prior (classifier.h:125-160) applied to the live tables */
        void set_priors(size_t k, double noise, const std::vector< double >& concentration) {
            check(phq_set_priors(handle_, static_cast< int >(k), noise, concentration.data()));
        }
        /* the decoder sections of Transcode::finalize's report (transcode.cpp:1811-1863) from the device accumulators */
        std::string report(uint64_t incoming_count = 0, uint64_t incoming_pf_count = 0, int precision = 15) {
            char* out(NULL);
            check(phq_report(handle_, incoming_count, incoming_pf_count, precision, &out));
            std::string text(out);
            phq_free(out);
            return text;
        }
        phq_handle* handle() { return handle_; }

    private:
        phq_handle* handle_;
        std::vector< phq_decoder_info > info_;
        void check(int status) { raise(status, phq_last_error(handle_)); }
};

}   /* namespace phq */
#endif
