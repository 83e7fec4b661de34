/*  oracle/_ref batched binding — TEST INFRASTRUCTURE, not product code.

    The reference-side binding a Pheniqs maintainer would add to use the B200 path, compiled
    against the reference's OWN headers and linked with its own unmodified objects (oracle/Makefile) plus the product
    library through its public C++ wrapper (include/pheniqs_b200.hpp). It stands where the per-read loop of

        TranscodingThread::run        /root/reference/transcode.h:202-225
        TranscodingDecoder::classify  /root/reference/transcode.h:51-65

    stands, with the loop turned inside out: a batch of the reference's `Read` objects is filled as a feed fills them,
    the `Segment` buffers of the barcode bearing input segments (BAM code bytes and Phred bytes, sequence.h:264-300)
    are handed to phq_decode_batch_bam as they are, and the per-read verdicts are scattered back through the
    reference's own `Read::append_to_* / update_* / set_RG` calls and `Read::flush` (read.h:187-285) — exactly the
    calls the Paml- / Md- / NaiveMolecularDecoder::classify members make after their scoring loop (pamld.cpp:133-180, mdd.cpp:96-138,
    naive.h:40-45). What comes out is what the reference's writer would emit: the auxiliary tags of every read and the
    QC fail flag; tests/test_gpu_binding.py diffs them with test/BDGGG/valid/annotated.out.

    Nothing of the reference is copied: the decoder subclasses below only reach protected members of the reference
    classes they derive from. The scoring itself (PamlDecoder / MdDecoder::classify) is never called here.
*/

#include "include.h"
#include "pamld.h"
#include "mdd.h"
#include "naive.h"

#include <pheniqs_b200.hpp>

#include <memory>

namespace {

enum Topic { T_SAMPLE = 0, T_MOLECULAR = 1, T_CELLULAR = 2 };

/* one decoder of the chain as the scatter step sees it */
class Scatter {
    public:
        virtual ~Scatter() {}
        virtual void apply(const Read& input, Read& output, const phq_result* verdict) = 0;
};

/*  A reference decoder D whose classify() is replaced by "take the verdict the device returned". The Observation is
    still extracted on the host by the reference's own Rule::apply (transform.h:142-169): the raw barcode tags are
    built from it, as in the reference. PROBABILISTIC = D carries decoding_confidence (PAMLD). */
template < class D, Topic TOPIC, bool PROBABILISTIC > class Batched : public D, public Scatter {
    public:
        Batched(const Value& ontology) : D(ontology) {}
        void apply(const Read& input, Read& output, const phq_result* verdict) override {
            this->observation.clear();
            this->rule.apply(input, this->observation);
            this->decoded = verdict->index > 0 ? &this->tag_array[static_cast< size_t >(verdict->index) - 1] : &this->unclassified;
            this->edit_distance = verdict->distance;
            route(output, verdict->confidence);
        }
    private:
        /* the statements that follow the base classify() call in pamld.cpp:133-180 / mdd.cpp:96-138, on the same members */
        template < Topic T = TOPIC > typename std::enable_if< T == T_SAMPLE >::type route(Read& output, double confidence) {
            output.append_to_raw_sample_barcode(this->observation);
            output.append_to_corrected_sample_barcode_sequence(*this->decoded, this->observation, this->corrected_quality);
            output.update_sample_distance(this->edit_distance);
            if(PROBABILISTIC) { output.update_sample_decoding_confidence(confidence); }
            output.set_RG(this->rg_by_barcode_index[this->decoded->index]);
        }
        template < Topic T = TOPIC > typename std::enable_if< T == T_CELLULAR >::type route(Read& output, double confidence) {
            output.append_to_raw_cellular_barcode(this->observation);
            output.append_to_corrected_cellular_barcode_sequence(*this->decoded, this->observation, this->corrected_quality);
            if(this->decoded->is_classified()) {
                if(PROBABILISTIC) { output.update_cellular_decoding_confidence(confidence); }
                output.update_cellular_distance(this->edit_distance);
            } else {
                if(PROBABILISTIC) { output.set_cellular_decoding_confidence(0); }
                output.set_cellular_distance(0);
            }
        }
        template < Topic T = TOPIC > typename std::enable_if< T == T_MOLECULAR >::type route(Read& output, double confidence) {
            output.append_to_raw_molecular_barcode(this->observation);
            output.append_to_corrected_molecular_barcode_sequence(*this->decoded, this->observation, this->corrected_quality);
            if(this->decoded->is_classified()) {
                if(PROBABILISTIC) { output.update_molecular_decoding_confidence(confidence); }
                output.update_molecular_distance(this->edit_distance);
            } else {
                if(PROBABILISTIC) { output.set_molecular_decoding_confidence(0); output.set_molecular_distance(0); }
                else { output.set_cellular_distance(0); }       /* mdd.cpp:136 clears the CELLULAR distance: reproduced */
            }
        }
};
class BatchedNaive : public NaiveMolecularDecoder, public Scatter {
    public:
        BatchedNaive(const Value& ontology) : NaiveMolecularDecoder(ontology) {}
        void apply(const Read& input, Read& output, const phq_result*) override {
            this->observation.clear();
            this->rule.apply(input, this->observation);
            output.append_to_raw_molecular_barcode(this->observation);
        }
};
class BatchedPassthrough : public Scatter {
    public:
        void apply(const Read&, Read&, const phq_result*) override {}
};

/* the factory of transcode.cpp:66-161 over the batched subclasses */
Scatter* make_scatter(const Value& v, Topic topic) {
    const string algorithm(decode_value_by_key< string >("algorithm", v));
    if(algorithm == "pamld") {
        switch(topic) {
            case T_SAMPLE:      return new Batched< PamlSampleDecoder, T_SAMPLE, true >(v);
            case T_MOLECULAR:   return new Batched< PamlMolecularDecoder, T_MOLECULAR, true >(v);
            case T_CELLULAR:    return new Batched< PamlCellularDecoder, T_CELLULAR, true >(v);
        }
    } else if(algorithm == "mdd") {
        switch(topic) {
            case T_SAMPLE:      return new Batched< MdSampleDecoder, T_SAMPLE, false >(v);
            case T_MOLECULAR:   return new Batched< MdMolecularDecoder, T_MOLECULAR, false >(v);
            case T_CELLULAR:    return new Batched< MdCellularDecoder, T_CELLULAR, false >(v);
        }
    } else if(algorithm == "naive" && topic == T_MOLECULAR) {
        return new BatchedNaive(v);
    } else if(algorithm == "passthrough") {
        return new BatchedPassthrough();
    }
    throw ConfigurationError("unsupported decoder algorithm " + algorithm);
}

void collect_topic(const Value& job, const char* key, Topic topic, vector< std::unique_ptr< Scatter > >& chain) {
    Value::ConstMemberIterator r = job.FindMember(key);
    if(r == job.MemberEnd()) { return; }
    if(r->value.IsObject()) { chain.emplace_back(make_scatter(r->value, topic)); }
    else if(r->value.IsArray()) { for(const auto& e : r->value.GetArray()) { chain.emplace_back(make_scatter(e, topic)); } }
}

string last_error;

}   /* namespace */

extern "C" {

const char* phq_binding_last_error() { return last_error.c_str(); }

/*  One feed batch through the binding. Inputs as the oracle driver takes them (flat BAM code / Phred arrays per input
    segment); outputs as phq_ref_tags returns them: text[r][t] (stride bytes, NUL terminated) for t = RG BC QT RX QX OX
    BZ CB CR CY, XB XM XC as floats (0 = absent), the final QC fail flags, and the job report of the device
    accumulators (malloc'd, caller frees with free()). `batch_reads` is the feed's buffer capacity: reads are
    classified that many per phq_decode_batch_bam call (configuration.json:369). */
int32_t phq_binding_run(const char* compiled_job_json, int32_t device, int64_t n_reads, int32_t n_segments, int64_t batch_reads,
                        const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                        const uint8_t* qcfail_in, int32_t stride, char* text, float* error_probability, uint8_t* out_qcfail, char** report_json) {
    try {
        Document job;
        if(job.Parse< rapidjson::kParseFullPrecisionFlag >(compiled_job_json).HasParseError()) { throw ConfigurationError("JSON parse error"); }
        vector< std::unique_ptr< Scatter > > chain;       /* sample, then molecular[], then cellular[] : transcode.h:51-60 */
        collect_topic(job, "sample", T_SAMPLE, chain);
        collect_topic(job, "molecular", T_MOLECULAR, chain);
        collect_topic(job, "cellular", T_CELLULAR, chain);

        phq::BatchDecoder decoder(compiled_job_json, device);
        if(decoder.decoder_cardinality() != chain.size()) { throw InternalError("decoder chains disagree"); }
        const size_t n_decoder(chain.size());
        if(batch_reads < 1) { batch_reads = 2048; }

        /* the feed's Read pool (transcode.h:178-179: the reference reuses one pair; a batch needs `batch_reads` of them) */
        vector< std::unique_ptr< Read > > pool;
        for(int64_t i(0); i < batch_reads; ++i) {
            pool.emplace_back(new Read(n_segments, Platform::ILLUMINA, 0));
            pool.back()->clear();
        }
        Read output(1, Platform::ILLUMINA, 0);
        output.clear();

        vector< vector< uint8_t > > staged_code(static_cast< size_t >(n_segments)), staged_quality(static_cast< size_t >(n_segments));
        vector< vector< int64_t > > staged_offset(static_cast< size_t >(n_segments));
        vector< vector< phq_result > > verdict(n_decoder);
        vector< phq_result* > verdict_pointer(n_decoder, static_cast< phq_result* >(NULL));
        vector< uint8_t > flag_in(static_cast< size_t >(batch_reads)), flag_out(static_cast< size_t >(batch_reads));
        for(size_t k(0); k < n_decoder; ++k) {
            verdict[k].resize(static_cast< size_t >(batch_reads));
            verdict_pointer[k] = verdict[k].data();
        }

        for(int64_t begin(0); begin < n_reads; begin += batch_reads) {
            const int64_t count(std::min(batch_reads, n_reads - begin));
            /* the feed: fill the Reads of the batch */
            for(int64_t i(0); i < count; ++i) {
                Read& input(*pool[static_cast< size_t >(i)]);
                for(int32_t s(0); s < n_segments; ++s) {
                    const int64_t from(offset[s][begin + i]);
                    const int64_t to(offset[s][begin + i + 1]);
                    input[s].fill(code[s] + from, quality[s] + from, static_cast< int32_t >(to - from));
                }
                input.set_qcfail(qcfail_in != NULL && qcfail_in[begin + i]);
                flag_in[static_cast< size_t >(i)] = input.qcfail() ? 1 : 0;
            }
            /* the seam: the Segment buffers of the batch, end to end per input segment, as phq_raw_segment */
            vector< phq_raw_segment > segments(static_cast< size_t >(n_segments));
            for(int32_t s(0); s < n_segments; ++s) {
                staged_code[s].clear(); staged_quality[s].clear(); staged_offset[s].assign(1, 0);
                for(int64_t i(0); i < count; ++i) {
                    const Segment& segment((*pool[static_cast< size_t >(i)])[s]);
                    staged_code[s].insert(staged_code[s].end(), segment.code, segment.code + segment.length);
                    staged_quality[s].insert(staged_quality[s].end(), segment.quality, segment.quality + segment.length);
                    staged_offset[s].push_back(static_cast< int64_t >(staged_code[s].size()));
                }
                segments[s].sequence = staged_code[s].data();
                segments[s].quality = staged_quality[s].data();
                segments[s].offset = staged_offset[s].data();
                segments[s].length = 0;
            }
            decoder.classify_bam(count, segments, flag_in.data(), verdict_pointer, flag_out.data());
            /* the scatter: verdicts back into the reference's Read, then Read::flush */
            for(int64_t i(0); i < count; ++i) {
                const Read& input(*pool[static_cast< size_t >(i)]);
                const int64_t r(begin + i);
                for(auto& segment : output) { segment.set_qcfail(input.qcfail()); }
                for(size_t k(0); k < n_decoder; ++k) { chain[k]->apply(input, output, &verdict[k][static_cast< size_t >(i)]); }
                output.set_qcfail(flag_out[static_cast< size_t >(i)] != 0);
                output.flush();
                const Auxiliary& a(output.auxiliary());
                const kstring_t* field[10] = { &a.RG, &a.BC, &a.QT, &a.RX, &a.QX, &a.OX, &a.BZ, &a.CB, &a.CR, &a.CY };
                for(int32_t t(0); t < 10; ++t) {
                    char* const to(text + (static_cast< size_t >(r) * 10 + t) * stride);
                    memset(to, 0, stride);
                    if(field[t]->l > 0 && field[t]->s != NULL) {
                        if(static_cast< int32_t >(field[t]->l) >= stride) { throw InternalError("tag longer than the stride"); }
                        memcpy(to, field[t]->s, field[t]->l);
                    }
                }
                error_probability[r * 3 + 0] = a.XB;
                error_probability[r * 3 + 1] = a.XM;
                error_probability[r * 3 + 2] = a.XC;
                if(out_qcfail != NULL) { out_qcfail[r] = output.qcfail() ? 1 : 0; }
                output.clear();
            }
            for(int64_t i(0); i < count; ++i) { pool[static_cast< size_t >(i)]->clear(); }
        }
        if(report_json != NULL) {
            const string report(decoder.report(static_cast< uint64_t >(n_reads), 0));
            *report_json = static_cast< char* >(malloc(report.size() + 1));
            memcpy(*report_json, report.c_str(), report.size() + 1);
        }
        return 0;
    } catch(std::exception& e) {
        last_error.assign(e.what());
        return -1;
    }
}

}   /* extern "C" */
