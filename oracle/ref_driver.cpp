/*  oracle/_ref driver — TEST INFRASTRUCTURE, not product code.

    Links the reference's OWN, UNMODIFIED decoder classes (compiled from the
    sources where they lie under /root/reference by oracle/Makefile) and drives
    them exactly the way the reference's per-thread loop does:

        TranscodingThread::run        /root/reference/transcode.h:202-225
        TranscodingDecoder::classify  /root/reference/transcode.h:51-65
        decoder factory by algorithm  /root/reference/transcode.cpp:66-161
        collect / finalize            /root/reference/transcode.cpp:162-195

    Nothing of the reference is copied here: this file only instantiates the
    reference classes (through thin "probe" subclasses that expose protected
    members read-only) and marshals flat arrays in and out over a C ABI so the
    Python test-suite / bench.py cpu_baseline leg can call it through ctypes.

    Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference legs may load the resulting oracle/_ref/libpheniqs_ref.so.
*/

#include "include.h"
#include "pamld.h"
#include "mdd.h"
#include "naive.h"

#include <chrono>
#include <memory>

namespace {

/* read-only window on the protected state of a reference decoder */
template < class D > class Probe : public D {
    public:
        Probe(const Value& ontology) : D(ontology) {}
        int32_t probe_index() const { return this->decoded->index; }
        int32_t probe_distance() const { return this->edit_distance; }
        const vector< Barcode >& probe_tags() const { return this->tag_array; }
        const Barcode& probe_unclassified() const { return this->unclassified; }
};
template < class D > class PamlProbe : public Probe< D > {
    public:
        PamlProbe(const Value& ontology) : Probe< D >(ontology) {}
        double probe_confidence() const { return this->decoding_confidence; }
};
class PassthroughProbe : public Classifier< Barcode > {
    public:
        PassthroughProbe(const Value& ontology) : Classifier< Barcode >(ontology) {}
        int32_t probe_index() const { return this->decoded->index; }
        const vector< Barcode >& probe_tags() const { return this->tag_array; }
        const Barcode& probe_unclassified() const { return this->unclassified; }
};

enum Kind { K_PAMLD = 0, K_MDD = 1, K_NAIVE = 2, K_PASSTHROUGH = 3 };
enum Topic { T_SAMPLE = 0, T_MOLECULAR = 1, T_CELLULAR = 2 };

struct Slot {
    Classifier< Barcode >* classifier;
    Kind kind;
    Topic topic;
    std::function< void(int32_t&, int32_t&, double&) > probe;
    std::function< const vector< Barcode >&() > tags;
    std::function< const Barcode&() > unclassified;
};

template < class P > void bind_common(Slot& slot, P* p) {
    slot.classifier = p;
    slot.tags = [p]() -> const vector< Barcode >& { return p->probe_tags(); };
    slot.unclassified = [p]() -> const Barcode& { return p->probe_unclassified(); };
}
template < class D > void make_paml(Slot& slot, const Value& v) {
    auto* p = new PamlProbe< D >(v);
    bind_common(slot, p);
    slot.kind = K_PAMLD;
    slot.probe = [p](int32_t& i, int32_t& d, double& c) { i = p->probe_index(); d = p->probe_distance(); c = p->probe_confidence(); };
}
template < class D > void make_md(Slot& slot, const Value& v) {
    auto* p = new Probe< D >(v);
    bind_common(slot, p);
    slot.kind = K_MDD;
    slot.probe = [p](int32_t& i, int32_t& d, double& c) { i = p->probe_index(); d = p->probe_distance(); c = 0; };
}
void make_naive(Slot& slot, const Value& v) {
    auto* p = new Probe< NaiveMolecularDecoder >(v);
    bind_common(slot, p);
    slot.kind = K_NAIVE;
    slot.probe = [p](int32_t& i, int32_t& d, double& c) { i = p->probe_index(); d = 0; c = 0; };
}
void make_passthrough(Slot& slot, const Value& v) {
    auto* p = new PassthroughProbe(v);
    bind_common(slot, p);
    slot.kind = K_PASSTHROUGH;
    slot.probe = [p](int32_t& i, int32_t& d, double& c) { i = p->probe_index(); d = 0; c = 0; };
}

/* the factory of transcode.cpp:66-161, expressed over the probe subclasses */
Slot make_slot(const Value& v, Topic topic) {
    Slot slot;
    slot.topic = topic;
    string algorithm(decode_value_by_key< string >("algorithm", v));
    if(algorithm == "pamld") {
        switch(topic) {
            case T_SAMPLE:      make_paml< PamlSampleDecoder >(slot, v); break;
            case T_MOLECULAR:   make_paml< PamlMolecularDecoder >(slot, v); break;
            case T_CELLULAR:    make_paml< PamlCellularDecoder >(slot, v); break;
        }
    } else if(algorithm == "mdd") {
        switch(topic) {
            case T_SAMPLE:      make_md< MdSampleDecoder >(slot, v); break;
            case T_MOLECULAR:   make_md< MdMolecularDecoder >(slot, v); break;
            case T_CELLULAR:    make_md< MdCellularDecoder >(slot, v); break;
        }
    } else if(algorithm == "naive" && topic == T_MOLECULAR) {
        make_naive(slot, v);
    } else if(algorithm == "passthrough") {
        make_passthrough(slot, v);
    } else {
        throw ConfigurationError("unsupported decoder algorithm " + algorithm);
    }
    return slot;
}

/* one private decoder set, as each TranscodingThread owns (transcode.cpp:2296) */
struct DecoderSet {
    vector< Slot > chain;   /* sample, then molecular[], then cellular[] : transcode.h:51-60 */
    uint64_t count;
    uint64_t pf_count;
    DecoderSet(const Value& job) : count(0), pf_count(0) {
        Value::ConstMemberIterator r = job.FindMember("sample");
        if(r != job.MemberEnd() && r->value.IsObject()) {
            chain.push_back(make_slot(r->value, T_SAMPLE));
        }
        r = job.FindMember("molecular");
        if(r != job.MemberEnd()) {
            if(r->value.IsObject()) { chain.push_back(make_slot(r->value, T_MOLECULAR)); }
            else if(r->value.IsArray()) { for(const auto& e : r->value.GetArray()) { chain.push_back(make_slot(e, T_MOLECULAR)); } }
        }
        r = job.FindMember("cellular");
        if(r != job.MemberEnd()) {
            if(r->value.IsObject()) { chain.push_back(make_slot(r->value, T_CELLULAR)); }
            else if(r->value.IsArray()) { for(const auto& e : r->value.GetArray()) { chain.push_back(make_slot(e, T_CELLULAR)); } }
        }
    }
    ~DecoderSet() { for(auto& s : chain) { delete s.classifier; } }
    void collect(const DecoderSet& other) {
        count += other.count;
        pf_count += other.pf_count;
        for(size_t i(0); i < chain.size(); ++i) { chain[i].classifier->collect(*other.chain[i].classifier); }
    }
};

struct Handle {
    Document job;
    int32_t input_segment_cardinality;
    std::unique_ptr< DecoderSet > total;    /* the job-level collect target (transcode.cpp:1770,317-320) */
    bool finalized;
    string report;
    string error;
};

struct Batch {
    int64_t n_reads;
    int32_t n_segments;
    const uint8_t* const* code;      /* [n_segments] flat BAM-code bytes */
    const uint8_t* const* quality;   /* [n_segments] flat phred bytes */
    const int64_t* const* offset;    /* [n_segments][n_reads + 1] */
    const uint8_t* qcfail_in;        /* [n_reads] or NULL */
};

/* the body of TranscodingThread::run (transcode.h:202-225) over reads [begin, end) */
void run_slice(DecoderSet& set, const Batch& b, int64_t begin, int64_t end,
               int32_t* out_index, int32_t* out_distance, double* out_confidence,
               uint8_t* out_qcfail, uint32_t* out_read_distance, double* out_read_confidence, int32_t* out_channel) {

    const size_t n_decoder(set.chain.size());
    Read input(b.n_segments, Platform::ILLUMINA, 0);
    Read output(1, Platform::ILLUMINA, 0);
    input.clear();
    output.clear();
    for(int64_t r(begin); r < end; ++r) {
        for(int32_t s(0); s < b.n_segments; ++s) {
            const int64_t from(b.offset[s][r]);
            const int64_t to(b.offset[s][r + 1]);
            input[s].fill(b.code[s] + from, b.quality[s] + from, static_cast< int32_t >(to - from));
        }
        const bool qcfail(b.qcfail_in != NULL && b.qcfail_in[r]);
        input.set_qcfail(qcfail);
        for(auto& segment : output) { segment.set_qcfail(qcfail); }

        for(size_t k(0); k < n_decoder; ++k) {
            Slot& slot(set.chain[k]);
            slot.classifier->classify(input, output);
            if(out_index != NULL) {
                int32_t i, d; double c;
                slot.probe(i, d, c);
                out_index[r * n_decoder + k] = i;
                out_distance[r * n_decoder + k] = d;
                out_confidence[r * n_decoder + k] = c;
            }
        }
        ++set.count;
        if(!output.qcfail()) { ++set.pf_count; }

        if(out_qcfail != NULL) { out_qcfail[r] = output.qcfail() ? 1 : 0; }
        if(out_read_distance != NULL) {
            out_read_distance[r * 3 + 0] = output.sample_distance;
            out_read_distance[r * 3 + 1] = output.molecular_distance;
            out_read_distance[r * 3 + 2] = output.cellular_distance;
        }
        if(out_read_confidence != NULL) {
            out_read_confidence[r * 3 + 0] = output.sample_decoding_confidence;
            out_read_confidence[r * 3 + 1] = output.molecular_decoding_confidence;
            out_read_confidence[r * 3 + 2] = output.cellular_decoding_confidence;
        }
        if(out_channel != NULL) { out_channel[r] = output.channel_index; }
        input.clear();
        output.clear();
    }
}

void copy_option(const AccumulatingOption& o, uint64_t* u, double* f) {
    /* field order of AccumulatingOption, selector.h:34-41 */
    u[0] = o.count;
    u[1] = o.pf_count;
    u[2] = o.accumulated_distance;
    u[3] = o.low_conditional_confidence_count;
    u[4] = o.low_confidence_count;
    u[5] = o.accumulated_pf_distance;
    f[0] = o.accumulated_confidence;
    f[1] = o.accumulated_pf_confidence;
}

}   /* namespace */

extern "C" {

void* phq_ref_create(const char* job_json, int32_t input_segment_cardinality, char* error, size_t error_capacity) {
    Handle* h(new Handle());
    try {
        h->input_segment_cardinality = input_segment_cardinality;
        h->finalized = false;
        /* kParseFullPrecisionFlag: rapidjson's default number parser is off by up to a few ulp, which the
           reference never sees for compiled values (it compiles and constructs decoders in one process
           from the same in-memory Document); correctly rounded parsing makes the text round trip exact. */
        if(h->job.Parse< rapidjson::kParseFullPrecisionFlag >(job_json).HasParseError()) {
            throw ConfigurationError(string("JSON parse error: ") + GetParseError_En(h->job.GetParseError()));
        }
        h->total.reset(new DecoderSet(h->job));
        return h;
    } catch(Error& e) {
        if(error != NULL && error_capacity > 0) {
            string m(e.what());
            for(auto& s : e.stack) { m.append(" <- "); m.append(s); }
            strncpy(error, m.c_str(), error_capacity - 1);
            error[error_capacity - 1] = '\0';
        }
    } catch(std::exception& e) {
        if(error != NULL && error_capacity > 0) {
            strncpy(error, e.what(), error_capacity - 1);
            error[error_capacity - 1] = '\0';
        }
    }
    delete h;
    return NULL;
}
void phq_ref_destroy(void* handle) {
    delete static_cast< Handle* >(handle);
}
int32_t phq_ref_decoder_count(void* handle) {
    return static_cast< int32_t >(static_cast< Handle* >(handle)->total->chain.size());
}
int32_t phq_ref_barcode_count(void* handle, int32_t decoder) {
    return static_cast< int32_t >(static_cast< Handle* >(handle)->total->chain[decoder].tags().size());
}

/*  Decode n_reads with n_threads private decoder sets (contiguous slices) and
    collect them into the handle's job-level set. Returns the wall-clock seconds
    spent inside the per-thread loops (thread spawn + join included), or -1. */
double phq_ref_decode(void* handle, int64_t n_reads, int32_t n_segments,
                      const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                      const uint8_t* qcfail_in, int32_t n_threads,
                      int32_t* out_index, int32_t* out_distance, double* out_confidence,
                      uint8_t* out_qcfail, uint32_t* out_read_distance, double* out_read_confidence, int32_t* out_channel) {
    Handle* h(static_cast< Handle* >(handle));
    try {
        Batch b = { n_reads, n_segments, code, quality, offset, qcfail_in };
        if(n_threads < 1) { n_threads = 1; }
        vector< std::unique_ptr< DecoderSet > > part;
        for(int32_t t(0); t < n_threads; ++t) { part.emplace_back(new DecoderSet(h->job)); }

        auto t0(std::chrono::steady_clock::now());
        if(n_threads == 1) {
            run_slice(*part[0], b, 0, n_reads, out_index, out_distance, out_confidence, out_qcfail, out_read_distance, out_read_confidence, out_channel);
        } else {
            vector< thread > pool;
            for(int32_t t(0); t < n_threads; ++t) {
                const int64_t begin(n_reads * t / n_threads);
                const int64_t end(n_reads * (t + 1) / n_threads);
                pool.emplace_back([&, t, begin, end]() {
                    run_slice(*part[t], b, begin, end, out_index, out_distance, out_confidence, out_qcfail, out_read_distance, out_read_confidence, out_channel);
                });
            }
            for(auto& t : pool) { t.join(); }
        }
        auto t1(std::chrono::steady_clock::now());
        for(auto& p : part) { h->total->collect(*p); }
        return std::chrono::duration< double >(t1 - t0).count();
    } catch(std::exception& e) {
        h->error.assign(e.what());
        return -1;
    }
}

/*  The tags the reference's OUTPUT carries for every read: one reference thread classifies the
    reads in order, Read::flush (read.h:187-237) assembles the auxiliary fields, and the strings Auxiliary::encode
    (auxiliary.cpp:320-361) would append are copied out: text[r][t] (stride bytes each, NUL terminated) for t =
    RG BC QT RX QX OX BZ CB CR CY, and XB XM XC as floats (0 = absent). Accumulates like phq_ref_decode. */
int32_t phq_ref_tags(void* handle, int64_t n_reads, int32_t n_segments,
                     const uint8_t* const* code, const uint8_t* const* quality, const int64_t* const* offset,
                     const uint8_t* qcfail_in, int32_t stride, char* text, float* error_probability, uint8_t* out_qcfail) {
    Handle* h(static_cast< Handle* >(handle));
    try {
        DecoderSet set(h->job);
        Read input(n_segments, Platform::ILLUMINA, 0);
        Read output(1, Platform::ILLUMINA, 0);
        input.clear();
        output.clear();
        for(int64_t r(0); r < n_reads; ++r) {
            for(int32_t s(0); s < n_segments; ++s) {
                const int64_t from(offset[s][r]);
                const int64_t to(offset[s][r + 1]);
                input[s].fill(code[s] + from, quality[s] + from, static_cast< int32_t >(to - from));
            }
            const bool qcfail(qcfail_in != NULL && qcfail_in[r]);
            input.set_qcfail(qcfail);
            for(auto& segment : output) { segment.set_qcfail(qcfail); }
            for(auto& slot : set.chain) { slot.classifier->classify(input, output); }
            ++set.count;
            if(!output.qcfail()) { ++set.pf_count; }
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
            input.clear();
            output.clear();
        }
        h->total->collect(set);
        return 0;
    } catch(std::exception& e) {
        h->error.assign(e.what());
        return -1;
    }
}

/* raw accumulator tables of decoder k: row 0 = undetermined, rows 1..NB = codec order */
void phq_ref_accumulators(void* handle, int32_t decoder, uint64_t* u64_table /* [(NB+1)*6] */, double* f64_table /* [(NB+1)*2] */) {
    Handle* h(static_cast< Handle* >(handle));
    const Slot& slot(h->total->chain[decoder]);
    copy_option(slot.unclassified(), u64_table, f64_table);
    size_t row(1);
    for(const auto& tag : slot.tags()) {
        copy_option(tag, u64_table + row * 6, f64_table + row * 2);
        ++row;
    }
}
void phq_ref_totals(void* handle, uint64_t* count, uint64_t* pf_count) {
    Handle* h(static_cast< Handle* >(handle));
    *count = h->total->count;
    *pf_count = h->total->pf_count;
}

/*  Run the reference finalize() chain (pamld.h:40-48 / decoder.h:77-83 /
    classifier.h:94-124) once, and return estimated noise prior and the
    per-barcode estimated concentration priors of decoder k. */
void phq_ref_finalize(void* handle) {
    Handle* h(static_cast< Handle* >(handle));
    if(!h->finalized) {
        for(auto& slot : h->total->chain) { slot.classifier->finalize(); }
        h->finalized = true;
    }
}
void phq_ref_estimated_priors(void* handle, int32_t decoder, double* noise, double* concentration /* [NB] */) {
    Handle* h(static_cast< Handle* >(handle));
    phq_ref_finalize(handle);
    const Slot& slot(h->total->chain[decoder]);
    *noise = slot.classifier->estimated_noise_prior;
    size_t i(0);
    for(const auto& tag : slot.tags()) { concentration[i++] = tag.estimated_concentration_prior; }
}
/* the reference's own report encoding of decoder k (classifier.h:161-177), as a JSON string */
const char* phq_ref_report(void* handle, int32_t decoder, int32_t precision) {
    Handle* h(static_cast< Handle* >(handle));
    phq_ref_finalize(handle);
    Document document;
    document.SetObject();
    h->total->chain[decoder].classifier->encode(document, document);
    StringBuffer buffer;
    PrettyWriter< StringBuffer > writer(buffer);
    writer.SetMaxDecimalPlaces(precision);
    document.Accept(writer);
    h->report.assign(buffer.GetString(), buffer.GetSize());
    return h->report.c_str();
}
const char* phq_ref_last_error(void* handle) {
    return static_cast< Handle* >(handle)->error.c_str();
}

}   /* extern "C" */
