/*  report.hpp — the decoder sections of the job report and the prior adjusted job.

    Host C++ over the accumulator tables the kernels fill. Restates, for the decoder chain:

      AccumulatingOption::finalize / encode        selector.cpp:78-135
      AccumulatingSelector::finalize / encode      selector.cpp:185-247
      PamlDecoder::finalize                        pamld.h:40-48
      Decoder::finalize                            decoder.h:77-83
      Classifier::finalize / encode                classifier.h:94-124, 161-177
      Barcode::encode                              barcode.cpp:53-67
      TranscodingDecoder::finalize / encode        transcode.cpp:180-229
      Transcode::finalize (report assembly)        transcode.cpp:1811-1863  (read group metadata, clean, sort)
      apply_prior (the prior tool)                 tool/pheniqs-prior-api.py:39-56, adjusted: 186-215
      Classifier::adjust_prior                     classifier.h:125-160

    Feed statistics ("incoming") are the caller's: they are encoded when given.
*/
#ifndef PHQ_REPORT_HPP
#define PHQ_REPORT_HPP

#include "spec.hpp"

#include <map>

namespace phq {

/* the tables of one decoder: [(N + 1)][ACC_U64_COLUMNS] and [(N + 1)][ACC_F64_COLUMNS], row 0 = unclassified */
struct AccumulatorTables {
    const uint64_t* u64;
    const double* f64;
};

namespace report_detail {

enum { COUNT = 0, PF_COUNT = 1, DISTANCE = 2, LOW_CONDITIONAL = 3, LOW_CONFIDENCE = 4, PF_DISTANCE = 5, U64_COLUMNS = 6 };
enum { CONFIDENCE = 0, PF_CONFIDENCE = 1, F64_COLUMNS = 2 };

/* AccumulatingSelector after finalize (selector.h:62-92) */
struct Selector {
    uint64_t count, pf_count, classified_count, pf_classified_count;
    uint64_t accumulated_classified_distance, accumulated_pf_classified_distance;
    uint64_t low_conditional_confidence_count, low_confidence_count;
    double accumulated_classified_confidence, accumulated_pf_classified_confidence;
    double estimated_noise_prior;
    Selector() : count(0), pf_count(0), classified_count(0), pf_classified_count(0), accumulated_classified_distance(0),
        accumulated_pf_classified_distance(0), low_conditional_confidence_count(0), low_confidence_count(0),
        accumulated_classified_confidence(0), accumulated_pf_classified_confidence(0), estimated_noise_prior(0) {}
};

/* AccumulatingOption::finalize + encode for one row (selector.cpp:78-135) */
inline Json encode_option(const uint64_t* u, const double* f, const Selector& parent, double estimated_concentration_prior) {
    const uint64_t count(u[COUNT]), pf_count(u[PF_COUNT]);
    double average_distance(0), average_confidence(0), pooled_fraction(0), pooled_classified_fraction(0);
    double pf_fraction(0), average_pf_distance(0), average_pf_confidence(0), pf_pooled_fraction(0), pf_pooled_classified_fraction(0);
    if(count > 0) {
        average_distance = u[DISTANCE] / double(count);
        average_confidence = f[CONFIDENCE] / double(count);
        if(parent.count > 0) { pooled_fraction = double(count) / double(parent.count); }
        if(parent.classified_count > 0) { pooled_classified_fraction = double(count) / double(parent.classified_count); }
    }
    if(pf_count > 0) {
        pf_fraction = double(pf_count) / double(count);
        average_pf_distance = u[PF_DISTANCE] / double(pf_count);
        average_pf_confidence = f[PF_CONFIDENCE] / double(pf_count);
        if(parent.pf_count > 0) { pf_pooled_fraction = double(pf_count) / double(parent.pf_count); }
        if(parent.pf_classified_count > 0) { pf_pooled_classified_fraction = double(pf_count) / double(parent.pf_classified_count); }
    }
    Json o(Json::object());
    o.set("count", Json::integer(static_cast< int64_t >(count)));
    if(average_distance > 0) { o.set("average distance", Json::number(average_distance)); }
    if(average_confidence > 0) { o.set("average confidence", Json::number(average_confidence)); }
    if(u[LOW_CONDITIONAL] > 0) { o.set("low conditional confidence count", Json::integer(static_cast< int64_t >(u[LOW_CONDITIONAL]))); }
    if(u[LOW_CONFIDENCE] > 0) { o.set("low confidence count", Json::integer(static_cast< int64_t >(u[LOW_CONFIDENCE]))); }
    o.set("pooled fraction", Json::number(pooled_fraction));
    if(pooled_classified_fraction > 0) { o.set("pooled classified fraction", Json::number(pooled_classified_fraction)); }
    o.set("pf count", Json::integer(static_cast< int64_t >(pf_count)));
    if(average_pf_distance > 0) { o.set("average pf distance", Json::number(average_pf_distance)); }
    if(average_pf_confidence > 0) { o.set("average pf confidence", Json::number(average_pf_confidence)); }
    o.set("pf fraction", Json::number(pf_fraction));
    o.set("pf pooled fraction", Json::number(pf_pooled_fraction));
    if(pf_pooled_classified_fraction > 0) { o.set("pf pooled classified fraction", Json::number(pf_pooled_classified_fraction)); }
    if(estimated_concentration_prior > 0) { o.set("estimated concentration", Json::number(estimated_concentration_prior)); }
    return o;
}

/* the read group tags encode_value(HeadRGAtom) writes (atom.cpp:1104-1122), taken from a codec / undetermined record */
inline void encode_read_group(const Json& record, Json& element) {
    static const char* const TAG[] = { "ID", "BC", "CN", "DS", "DT", "FO", "KS", "LB", "PG", "PI", "PL", "PM", "PU", "SM" };
    for(const char* tag : TAG) {
        const Json* v(record.find(tag));
        if(v != NULL && v->is_string() && !v->as_string().empty()) { element.set(tag, *v); }
    }
}

/* json.cpp:834-874 */
inline void clean(Json& v) {
    switch(v.type()) {
        case Json::Bool: if(!v.as_bool()) { v = Json(); } break;
        case Json::String: if(v.as_string().empty()) { v = Json(); } break;
        case Json::Object: {
            Json kept(Json::object());
            for(auto& m : v.members()) {
                clean(m.second);
                if(!m.second.is_null()) { kept.set(m.first, m.second); }
            }
            v = kept.members().empty() ? Json() : kept;
            break;
        }
        case Json::Array: {
            Json kept(Json::array());
            for(auto& e : v.items()) {
                clean(e);
                if(!e.is_null()) { kept.push(e); }
            }
            v = kept.items().empty() ? Json() : kept;
            break;
        }
        default: break;
    }
}

}   /* namespace report_detail */

/*  One classifier's report element: finalize in the reference's order (PamlDecoder, Decoder, Classifier,
    AccumulatingSelector), then Classifier::encode. `element` is the decoder's compiled JSON (codec records carry
    the barcode segments), `d` its parsed form with the live priors. */
inline Json encode_classifier_report(const DecoderSpec& d, const Json& element, const AccumulatorTables& tables,
                                     double* estimated_noise = NULL, std::vector< double >* estimated_concentration = NULL) {
    using namespace report_detail;
    const int32_t N(d.barcode_cardinality);
    const uint64_t* const u(tables.u64);
    const double* const f(tables.f64);
    Selector s;
    /* pamld.h:40-48 */
    if(d.algorithm == PHQ_PAMLD) {
        for(int32_t b(1); b <= N; ++b) {
            s.accumulated_classified_confidence += f[b * F64_COLUMNS + CONFIDENCE];
            s.accumulated_pf_classified_confidence += f[b * F64_COLUMNS + PF_CONFIDENCE];
            s.low_conditional_confidence_count += u[b * U64_COLUMNS + LOW_CONDITIONAL];
            s.low_confidence_count += u[b * U64_COLUMNS + LOW_CONFIDENCE];
        }
    }
    /* decoder.h:77-83 */
    if(d.algorithm == PHQ_PAMLD || d.algorithm == PHQ_MDD) {
        for(int32_t b(1); b <= N; ++b) {
            s.accumulated_classified_distance += u[b * U64_COLUMNS + DISTANCE];
            s.accumulated_pf_classified_distance += u[b * U64_COLUMNS + PF_DISTANCE];
        }
    }
    /* classifier.h:94-124 */
    for(int32_t b(1); b <= N; ++b) {
        s.classified_count += u[b * U64_COLUMNS + COUNT];
        s.pf_classified_count += u[b * U64_COLUMNS + PF_COUNT];
    }
    s.count = s.classified_count + u[COUNT];
    s.pf_count = s.pf_classified_count + u[PF_COUNT];
    double estimated_noise_count(static_cast< double >(s.low_conditional_confidence_count));
    const double confident_noise_ratio(estimated_noise_count / (estimated_noise_count + s.pf_classified_count));
    if(s.low_confidence_count > 0) { estimated_noise_count += double(s.low_confidence_count) * confident_noise_ratio; }
    s.estimated_noise_prior = estimated_noise_count / double(s.count);
    const double estimated_not_noise_prior(1.0 - s.estimated_noise_prior);
    if(estimated_noise != NULL) { *estimated_noise = s.estimated_noise_prior; }
    if(estimated_concentration != NULL) { estimated_concentration->assign(static_cast< size_t >(N), 0.0); }

    /* selector.cpp:185-247 */
    double pf_fraction(0), classified_fraction(0), pf_classified_fraction(0), average_classified_distance(0), average_classified_confidence(0);
    double classified_pf_fraction(0), average_pf_classified_distance(0), average_pf_classified_confidence(0);
    if(s.count > 0) {
        pf_fraction = double(s.pf_count) / double(s.count);
        classified_fraction = double(s.classified_count) / double(s.count);
    }
    if(s.pf_count > 0) { pf_classified_fraction = double(s.pf_classified_count) / double(s.pf_count); }
    if(s.classified_count > 0) {
        average_classified_distance = s.accumulated_classified_distance / double(s.classified_count);
        average_classified_confidence = s.accumulated_classified_confidence / double(s.classified_count);
        classified_pf_fraction = double(s.pf_classified_count) / double(s.classified_count);
    }
    if(s.pf_classified_count > 0) {
        average_pf_classified_distance = s.accumulated_pf_classified_distance / double(s.pf_classified_count);
        average_pf_classified_confidence = s.accumulated_pf_classified_confidence / double(s.pf_classified_count);
    }
    Json o(Json::object());
    o.set("index", Json::integer(d.index));
    o.set("count", Json::integer(static_cast< int64_t >(s.count)));
    o.set("pf count", Json::integer(static_cast< int64_t >(s.pf_count)));
    o.set("classified count", Json::integer(static_cast< int64_t >(s.classified_count)));
    if(s.low_conditional_confidence_count > 0) { o.set("low conditional confidence count", Json::integer(static_cast< int64_t >(s.low_conditional_confidence_count))); }
    if(s.low_confidence_count > 0) { o.set("low confidence count", Json::integer(static_cast< int64_t >(s.low_confidence_count))); }
    o.set("pf classified count", Json::integer(static_cast< int64_t >(s.pf_classified_count)));
    o.set("pf fraction", Json::number(pf_fraction));
    o.set("classified fraction", Json::number(classified_fraction));
    if(average_classified_distance > 0) { o.set("average classified distance", Json::number(average_classified_distance)); }
    if(average_classified_confidence > 0) { o.set("average classified confidence", Json::number(average_classified_confidence)); }
    o.set("pf classified fraction", Json::number(pf_classified_fraction));
    o.set("classified pf fraction", Json::number(classified_pf_fraction));
    if(average_pf_classified_distance > 0) { o.set("average pf classified distance", Json::number(average_pf_classified_distance)); }
    if(average_pf_classified_confidence > 0) { o.set("average pf classified confidence", Json::number(average_pf_classified_confidence)); }
    if(s.estimated_noise_prior > 0) { o.set("estimated noise", Json::number(s.estimated_noise_prior)); }

    /* classifier.h:161-177, barcode.cpp:53-67 */
    Json unclassified(encode_option(u, f, s, 0.0));
    unclassified.set("index", Json::integer(0));
    o.set("unclassified", unclassified);
    if(N > 0) {
        std::vector< const Json* > record(static_cast< size_t >(N), NULL);
        const Json* const codec(element.find("codec"));
        if(codec != NULL && codec->is_object()) {
            for(const auto& m : codec->members()) {
                const int32_t index(get_int(m.second, "index"));
                if(index >= 1 && index <= N) { record[static_cast< size_t >(index - 1)] = &m.second; }
            }
        }
        Json classified(Json::array());
        for(int32_t b(1); b <= N; ++b) {
            /* element.finalize(*this) gives pf_pooled_classified_fraction; the estimate follows (classifier.h:118-120) */
            double pf_pooled_classified_fraction(0);
            if(u[b * U64_COLUMNS + PF_COUNT] > 0 && s.pf_classified_count > 0) {
                pf_pooled_classified_fraction = double(u[b * U64_COLUMNS + PF_COUNT]) / double(s.pf_classified_count);
            }
            const double estimate(estimated_not_noise_prior * pf_pooled_classified_fraction);
            if(estimated_concentration != NULL) { (*estimated_concentration)[static_cast< size_t >(b - 1)] = estimate; }
            Json e(encode_option(u + b * U64_COLUMNS, f + b * F64_COLUMNS, s, estimate));
            e.set("index", Json::integer(b));
            e.set("concentration", Json::number(d.concentration[static_cast< size_t >(b - 1)]));
            const Json* const source(record[static_cast< size_t >(b - 1)]);
            if(source != NULL && source->find("barcode") != NULL) { e.set("barcode", source->at("barcode")); }
            classified.push(e);
        }
        o.set("classified", classified);
    }
    return o;
}

/*  The decoder sections of Transcode::finalize's report (transcode.cpp:1811-1863): "outgoing", "sample",
    "molecular", "cellular" (+ "incoming" when the caller knows the feed counts), read group tags on the sample
    elements, cleaned and key sorted. `tables[k]` belongs to chain[k]. */
inline Json encode_job_report(const Json& job, const std::vector< DecoderSpec >& chain, const std::vector< AccumulatorTables >& tables,
                              uint64_t count, uint64_t pf_count, uint64_t incoming_count, uint64_t incoming_pf_count) {
    Json report(Json::object());
    if(incoming_count > 0) {
        Json e(Json::object());
        e.set("count", Json::integer(static_cast< int64_t >(incoming_count)));
        e.set("pf count", Json::integer(static_cast< int64_t >(incoming_pf_count)));
        e.set("pf fraction", Json::number(double(incoming_pf_count) / double(incoming_count)));
        report.set("incoming", e);
    }
    if(count > 0) {
        Json e(Json::object());
        e.set("count", Json::integer(static_cast< int64_t >(count)));
        e.set("pf count", Json::integer(static_cast< int64_t >(pf_count)));
        e.set("pf fraction", Json::number(double(pf_count) / double(count)));
        report.set("outgoing", e);
    }
    static const char* const TOPIC[] = { "sample", "molecular", "cellular" };
    for(int32_t topic : { PHQ_SAMPLE, PHQ_MOLECULAR, PHQ_CELLULAR }) {
        const std::vector< const Json* > elements(topic_elements(job, TOPIC[topic]));
        Json array(Json::array());
        for(size_t k(0); k < chain.size(); ++k) {
            if(chain[k].topic != topic) { continue; }
            const Json& element(*elements.at(static_cast< size_t >(topic == PHQ_SAMPLE ? 0 : chain[k].index)));
            Json section(encode_classifier_report(chain[k], element, tables[k]));
            if(topic == PHQ_SAMPLE) {
                /* read group metadata (transcode.cpp:1840-1859) */
                const Json* const undetermined(element.find("undetermined"));
                if(undetermined != NULL && undetermined->is_object()) { report_detail::encode_read_group(*undetermined, *section.find("unclassified")); }
                const Json* const codec(element.find("codec"));
                Json* const classified(section.find("classified"));
                if(codec != NULL && codec->is_object() && classified != NULL) {
                    for(const auto& m : codec->members()) {
                        const int32_t index(get_int(m.second, "index"));
                        if(index >= 1 && static_cast< size_t >(index) <= classified->items().size()) {
                            report_detail::encode_read_group(m.second, classified->items()[static_cast< size_t >(index - 1)]);
                        }
                    }
                }
                report.set("sample", section);
            } else {
                array.push(section);
            }
        }
        if(topic != PHQ_SAMPLE && !array.items().empty()) { report.set(TOPIC[topic], array); }
    }
    report_detail::clean(report);
    if(report.is_null()) { report = Json::object(); }
    report.sort_keys();
    return report;
}

/*  The prior adjusted job: apply_prior of the reference's prior tool (tool/pheniqs-prior-api.py:39-56) for the
    sample / molecular / cellular decoders of a job (elements matched by "index" when both sides are arrays, lines
    186-215), which is what Classifier::adjust_prior (classifier.h:125-160) writes for --prior: `noise` from
    "estimated noise", every codec record's `concentration` from the "estimated concentration" of the report record
    with the same barcode segments (0 when the report has the barcode without an estimate). */
inline void apply_prior(Json& decoder, const Json& section) {
    const Json* const noise(section.find("estimated noise"));
    if(noise != NULL && noise->is_number()) { decoder.set("noise", *noise); }
    Json* const codec(decoder.find("codec"));
    const Json* const classified(section.find("classified"));
    if(codec == NULL || !codec->is_object() || classified == NULL || !classified->is_array()) { return; }
    auto key_of = [](const Json& record) {
        std::string key;
        const Json* const segments(record.find("barcode"));
        if(segments != NULL && segments->is_array()) { for(const auto& s : segments->items()) { key += s.as_string(); } }
        return key;
    };
    std::map< std::string, const Json* > by_barcode;
    for(const auto& record : classified->items()) { by_barcode[key_of(record)] = &record; }
    for(auto& m : codec->members()) {
        if(!m.second.is_object() || m.second.find("barcode") == NULL) { continue; }
        const auto found(by_barcode.find(key_of(m.second)));
        if(found == by_barcode.end()) { continue; }
        const Json* const estimate(found->second->find("estimated concentration"));
        m.second.set("concentration", (estimate != NULL && estimate->is_number()) ? *estimate : Json::integer(0));
    }
}
inline Json adjust_job(const Json& job, const Json& report) {
    Json adjusted(job);
    for(const char* topic : { "sample", "cellular", "molecular" }) {
        Json* const model(adjusted.find(topic));
        const Json* const section(report.find(topic));
        if(model == NULL || section == NULL) { continue; }
        if(model->is_object() && section->is_object()) {
            apply_prior(*model, *section);
        } else if(model->is_array() && section->is_array()) {
            /* both sides carry "index"; elements without one are matched by position */
            for(size_t i(0); i < model->items().size(); ++i) {
                Json& item(model->items()[i]);
                const int32_t index(item.is_object() && item.find("index") != NULL ? get_int(item, "index") : static_cast< int32_t >(i));
                for(size_t j(0); j < section->items().size(); ++j) {
                    const Json& candidate(section->items()[j]);
                    const int32_t other(candidate.find("index") != NULL ? get_int(candidate, "index") : static_cast< int32_t >(j));
                    if(other == index) { apply_prior(item, candidate); break; }
                }
            }
        }
    }
    return adjusted;
}

}   /* namespace phq */
#endif
