/*  spec.hpp — decoder ontology: from JSON to the flat specification the device uses.

    Two entry points:

      compile_job()            the decoder part of the reference's job compiler
                               (transcode.cpp:735-768 default knit, :824-940 cardinalities, random
                               barcode probability and barcode validation, :941-1039 index order,
                               undetermined, concentration normalisation; metric.h:87-111,216-242
                               Shannon bound and distance tolerance; configuration.json:368-376,
                               423-501 defaults).
      parse_compiled_job()     what the decoder constructors read from a compiled ontology
                               (classifier.h:54-60, decoder.h:44-52, pamld.cpp:24-31, mdd.cpp:24-27,
                               barcode.cpp:24-46, transform.cpp:100-331; factory transcode.cpp:66-161).

    File:line citations are relative to the reference tree.
*/
#ifndef PHQ_SPEC_HPP
#define PHQ_SPEC_HPP

#include "json.hpp"
#include "../../include/pheniqs_b200.h"

#include <cstdlib>
#include <climits>
#include <map>
#include <set>

namespace phq {

/* the reference's Error hierarchy (error.h:46-136) reduced to code + message */
struct Error : public std::runtime_error {
    int code;
    Error(int code, const std::string& message) : std::runtime_error(message), code(code) {}
};
struct ConfigurationError : public Error {
    explicit ConfigurationError(const std::string& message) : Error(PHQ_CONFIGURATION_ERROR, "Configuration error : " + message) {}
};
struct InternalError : public Error {
    explicit InternalError(const std::string& message) : Error(PHQ_INTERNAL_ERROR, "Internal error : " + message) {}
};
struct SequenceError : public Error {
    explicit SequenceError(const std::string& message) : Error(PHQ_SEQUENCE_ERROR, "Sequence error : " + message) {}
};
struct OverflowError : public Error {
    explicit OverflowError(const std::string& message) : Error(PHQ_OVERFLOW_ERROR, "Overflow error : " + message) {}
};

/* iupac.h:153-171 AsciiToAmbiguousBam */
inline uint8_t ascii_to_bam(unsigned char c) {
    switch(c) {
        case '=': return 0;
        case 'A': case 'a': case '0': return 1;
        case 'C': case 'c': case '1': return 2;
        case 'M': case 'm': return 3;
        case 'G': case 'g': case '2': return 4;
        case 'R': case 'r': return 5;
        case 'S': case 's': return 6;
        case 'V': case 'v': return 7;
        case 'T': case 't': case '3': return 8;
        case 'W': case 'w': return 9;
        case 'Y': case 'y': return 10;
        case 'H': case 'h': return 11;
        case 'K': case 'k': return 12;
        case 'D': case 'd': return 13;
        case 'B': case 'b': return 14;
        default: return 15;
    }
}

/* Token + Transform (transform.h:36-121) */
struct TransformSpec {
    int32_t token_index;
    int32_t input_segment_index;
    int32_t start;
    int32_t end;
    bool end_terminated;
    int32_t output_segment_index;
    bool reverse_complement;

    bool constant() const {                 /* transform.h:59-64 */
        if(end_terminated) { return (start >= 0 && end >= 0) || (start < 0 && end < 0); }
        return start < 0;
    }
    bool empty() const {                    /* transform.h:47-49 */
        return (end_terminated && start >= end) && ((start >= 0 && end >= 0) || (start < 0 && end < 0));
    }
    int32_t length() const {                /* transform.h:50-58 */
        if(constant()) {
            if(end_terminated) { return empty() ? 0 : end - start; }
            return -start;
        }
        return -1;
    }
    int32_t absolute_end(int32_t n) const {     /* transform.h:65-72 */
        if(end_terminated) {
            if(end < 0) { int32_t v(n + end); return v < 0 ? 0 : v; }
            return end > n ? n : end;
        }
        return n;
    }
    int32_t absolute_start(int32_t n) const {   /* transform.h:73-80 */
        if(start < 0) { int32_t v(n + start); return v < 0 ? 0 : v; }
        return start > n ? 0 : start;
    }
};

struct DecoderSpec {
    int32_t algorithm;
    int32_t topic;
    int32_t index;
    int32_t barcode_cardinality;                /* N, undetermined excluded */
    int32_t segment_cardinality;
    int32_t nucleotide_cardinality;
    std::vector< int32_t > segment_length;
    std::vector< int32_t > segment_offset;
    std::vector< TransformSpec > transform;
    std::vector< uint8_t > barcode;             /* [N][nucleotide_cardinality] BAM codes, segments concatenated */
    std::vector< double > concentration;        /* [N] */
    std::vector< std::string > barcode_key;     /* [N] codec keys, index order */
    double noise;
    double confidence_threshold;
    double random_barcode_probability;
    int32_t high_quality_threshold;
    int32_t high_quality_distance_threshold;
    int32_t quality_masking_threshold;
    std::vector< int32_t > distance_tolerance;
    bool multiplexing_classifier;
    int32_t corrected_quality;

    bool tiled() const { return algorithm == PHQ_PAMLD || algorithm == PHQ_MDD; }
    int32_t word_cardinality() const { return (nucleotide_cardinality + 15) / 16; }
    int32_t quality_word_cardinality() const { return (nucleotide_cardinality + 3) / 4; }
};

/* ------------------------------------------------------------------ small typed getters
   decode_value_by_key semantics (json.cpp:320-437): missing or null -> zero value */
inline double get_double(const Json& o, const char* key, double fallback = 0) {
    const Json* v(o.find(key));
    if(v == NULL || v->is_null()) { return fallback; }
    if(!v->is_number()) { throw ConfigurationError(std::string(key) + " element is not a number"); }
    return v->as_double();
}
inline int32_t get_int(const Json& o, const char* key, int32_t fallback = 0) {
    const Json* v(o.find(key));
    if(v == NULL || v->is_null()) { return fallback; }
    if(!v->is_number() || v->as_double() != std::floor(v->as_double())) { throw ConfigurationError(std::string(key) + " element is not a 32 bit integer"); }
    return static_cast< int32_t >(v->as_double());
}
inline bool get_bool(const Json& o, const char* key) {
    const Json* v(o.find(key));
    if(v == NULL || v->is_null()) { return false; }
    if(!v->is_bool()) { throw ConfigurationError(std::string(key) + " element is not a boolean"); }
    return v->as_bool();
}
inline std::string get_string(const Json& o, const char* key) {
    const Json* v(o.find(key));
    if(v == NULL || v->is_null()) { return std::string(); }
    if(!v->is_string()) { throw ConfigurationError(std::string(key) + " element is not a string"); }
    return v->as_string();
}

inline int32_t algorithm_from_string(const std::string& name) {
    if(name == "pamld") { return PHQ_PAMLD; }
    if(name == "mdd") { return PHQ_MDD; }
    if(name == "naive") { return PHQ_NAIVE; }
    if(name == "passthrough") { return PHQ_PASSTHROUGH; }
    throw ConfigurationError("unsupported decoder algorithm " + name);
}

/* token "segment:start:end" (transform.cpp:100-127; pattern configuration.json:1427) */
inline void parse_token(const std::string& pattern, int32_t& segment, int32_t& start, int32_t& end, bool& end_terminated) {
    size_t a(pattern.find(':'));
    size_t b(a == std::string::npos ? std::string::npos : pattern.find(':', a + 1));
    if(a == std::string::npos || b == std::string::npos || pattern.find(':', b + 1) != std::string::npos) {
        throw ConfigurationError("illegal token syntax " + pattern);
    }
    auto integer = [&](const std::string& text, bool negative_allowed) -> int32_t {
        if(text.empty()) { throw ConfigurationError("illegal token syntax " + pattern); }
        size_t i(0);
        if(text[0] == '-') { if(!negative_allowed) { throw ConfigurationError("illegal token syntax " + pattern); } i = 1; }
        if(i >= text.size()) { throw ConfigurationError("illegal token syntax " + pattern); }
        for(size_t k(i); k < text.size(); ++k) { if(text[k] < '0' || text[k] > '9') { throw ConfigurationError("illegal token syntax " + pattern); } }
        return static_cast< int32_t >(std::stol(text));
    };
    std::string s0(pattern.substr(0, a)), s1(pattern.substr(a + 1, b - a - 1)), s2(pattern.substr(b + 1));
    segment = integer(s0, false);
    start = s1.empty() ? 0 : integer(s1, true);
    end_terminated = !s2.empty();
    end = s2.empty() ? 0 : integer(s2, true);
}

/* Rule (transform.cpp:252-331): knit elements are ':' separated token references, '~' prefix = reverse complement */
inline void parse_rule(const Json& transform, std::vector< TransformSpec >& out, int32_t& output_segment_cardinality) {
    if(!transform.is_object()) { throw ConfigurationError("element transform must be a dictionary"); }
    const Json* token(transform.find("token"));
    if(token == NULL || !token->is_array()) { throw ConfigurationError("transform element is missing a token array"); }
    std::vector< TransformSpec > tokens;
    int32_t token_index(0);
    for(const auto& element : token->items()) {
        if(!element.is_string()) { throw ConfigurationError("token element must be a string"); }
        TransformSpec t;
        t.token_index = token_index++;
        parse_token(element.as_string(), t.input_segment_index, t.start, t.end, t.end_terminated);
        t.output_segment_index = 0;
        t.reverse_complement = false;
        tokens.push_back(t);
    }
    std::vector< std::string > knit;
    const Json* k(transform.find("knit"));
    if(k == NULL || k->is_null() || (k->is_array() && k->items().empty())) {
        for(size_t i(0); i < tokens.size(); ++i) { knit.push_back(std::to_string(i)); }      /* transcode.cpp:735-768 */
    } else {
        if(!k->is_array()) { throw ConfigurationError("rule observation element must be an array"); }
        for(const auto& element : k->items()) {
            if(!element.is_string()) { throw ConfigurationError("transform element must be a string"); }
            knit.push_back(element.as_string());
        }
    }
    out.clear();
    output_segment_cardinality = 0;
    for(const auto& pattern : knit) {
        size_t at(0);
        while(true) {
            size_t colon(pattern.find(':', at));
            std::string part(pattern.substr(at, colon == std::string::npos ? std::string::npos : colon - at));
            bool reverse(false);
            if(!part.empty() && part[0] == '~') { reverse = true; part.erase(0, 1); }
            if(part.empty()) { throw ConfigurationError("transform must explicitly specify a token reference"); }
            for(char c : part) { if(c < '0' || c > '9') { throw ConfigurationError(std::string("illegal character in transform ") + c); } }
            int32_t reference(static_cast< int32_t >(std::stol(part)));
            if(reference >= static_cast< int32_t >(tokens.size())) { throw ConfigurationError("invalid token reference " + std::to_string(reference) + " in transform"); }
            TransformSpec t(tokens[reference]);
            t.output_segment_index = output_segment_cardinality;
            t.reverse_complement = reverse;
            out.push_back(t);
            if(colon == std::string::npos) { break; }
            at = colon + 1;
        }
        ++output_segment_cardinality;
    }
}

inline std::vector< const Json* > topic_elements(const Json& job, const char* topic) {
    std::vector< const Json* > out;
    const Json* e(job.find(topic));
    if(e == NULL || e->is_null()) { return out; }
    if(e->is_object()) { out.push_back(e); }
    else if(e->is_array()) { for(const auto& d : e->items()) { out.push_back(&d); } }
    else { throw ConfigurationError(std::string(topic) + " decoder element must be a dictionary or an array"); }
    return out;
}

/* ------------------------------------------------------------------ compiled ontology -> DecoderSpec */
inline DecoderSpec parse_compiled_decoder(const Json& v, int32_t topic) {
    DecoderSpec d;
    d.topic = topic;
    d.algorithm = algorithm_from_string(get_string(v, "algorithm"));
    if(d.algorithm == PHQ_NAIVE && topic != PHQ_MOLECULAR) { throw ConfigurationError("unsupported decoder algorithm naive"); }
    d.index = get_int(v, "index");
    if(v.find("undetermined") == NULL) { throw ConfigurationError("classifier must declare an undetermined element"); }
    d.multiplexing_classifier = get_bool(v, "multiplexing classifier");
    d.corrected_quality = get_int(v, "corrected quality");
    d.noise = get_double(v, "noise");
    d.confidence_threshold = get_double(v, "confidence threshold");
    d.random_barcode_probability = get_double(v, "random barcode probability");
    d.high_quality_threshold = get_int(v, "high quality threshold");
    d.high_quality_distance_threshold = get_int(v, "high quality distance threshold");
    d.quality_masking_threshold = get_int(v, "quality masking threshold");
    d.segment_cardinality = 0;
    d.nucleotide_cardinality = 0;
    d.barcode_cardinality = 0;

    if(d.algorithm != PHQ_PASSTHROUGH) {
        const Json* transform(v.find("transform"));
        if(transform == NULL || transform->is_null()) { throw ConfigurationError("no element transform found"); }
        int32_t cardinality(0);
        parse_rule(*transform, d.transform, cardinality);
        d.segment_cardinality = get_int(v, "segment cardinality", cardinality);
        if(d.segment_cardinality != cardinality) { throw ConfigurationError("segment cardinality inconsistent with transform"); }
        if(d.segment_cardinality > PHQ_MAX_SEGMENTS) { throw ConfigurationError("more than " + std::to_string(PHQ_MAX_SEGMENTS) + " barcode segments are not supported on this path"); }
        d.segment_length.assign(d.segment_cardinality, 0);
        for(const auto& t : d.transform) {
            if(t.empty()) { throw ConfigurationError("token " + std::to_string(t.token_index) + " is empty"); }
            if(!t.constant()) { throw ConfigurationError("token " + std::to_string(t.token_index) + " is not fixed width"); }
            d.segment_length[t.output_segment_index] += t.length();
        }
        d.segment_offset.assign(d.segment_cardinality + 1, 0);
        for(int32_t i(0); i < d.segment_cardinality; ++i) { d.segment_offset[i + 1] = d.segment_offset[i] + d.segment_length[i]; }
        d.nucleotide_cardinality = d.segment_offset[d.segment_cardinality];
        int32_t declared(get_int(v, "nucleotide cardinality", d.nucleotide_cardinality));
        if(declared != d.nucleotide_cardinality) { throw ConfigurationError("nucleotide cardinality inconsistent with transform"); }
    }

    const Json* codec(v.find("codec"));
    if(codec != NULL && !codec->is_null()) {
        if(!codec->is_object()) { throw ConfigurationError("codec element must be a dictionary"); }
        /* barcode rows in "index" order; the reference iterates the (key sorted) codec and trusts it */
        std::vector< std::pair< int32_t, const Json::Member* > > order;
        for(const auto& record : codec->members()) {
            order.emplace_back(get_int(record.second, "index"), &record);
        }
        std::stable_sort(order.begin(), order.end(), [](const std::pair< int32_t, const Json::Member* >& a, const std::pair< int32_t, const Json::Member* >& b) { return a.first < b.first; });
        d.barcode_cardinality = static_cast< int32_t >(order.size());
        d.barcode.assign(static_cast< size_t >(d.barcode_cardinality) * d.nucleotide_cardinality, 0);
        d.concentration.assign(d.barcode_cardinality, 0);
        std::set< std::string > unique;
        for(int32_t i(0); i < d.barcode_cardinality; ++i) {
            const Json& record(order[i].second->second);
            if(order[i].first != i + 1) { throw ConfigurationError("barcode index must enumerate the codec from 1"); }
            d.barcode_key.push_back(order[i].second->first);
            d.concentration[i] = get_double(record, "concentration");
            const Json* barcode(record.find("barcode"));
            if(barcode == NULL || !barcode->is_array() || static_cast< int32_t >(barcode->items().size()) != d.segment_cardinality) {
                throw ConfigurationError("barcode must have exactly " + std::to_string(d.segment_cardinality) + " segments");
            }
            std::string flat;
            for(int32_t s(0); s < d.segment_cardinality; ++s) {
                const Json& segment(barcode->items()[s]);
                if(!segment.is_string()) { throw ConfigurationError("barcode segment " + std::to_string(s) + " must be a string"); }
                const std::string& text(segment.as_string());
                if(static_cast< int32_t >(text.size()) != d.segment_length[s]) {
                    throw ConfigurationError("expected " + std::to_string(d.segment_length[s]) + " but found " + std::to_string(text.size()) + " nucleotides in segment " + std::to_string(s) + " of barcode " + order[i].second->first);
                }
                for(size_t j(0); j < text.size(); ++j) {
                    uint8_t code(ascii_to_bam(static_cast< unsigned char >(text[j])));
                    if(d.tiled() && code != 1 && code != 2 && code != 4 && code != 8) {
                        /* configuration.json:546 allows [ATCG=]; '=' only ever appears in the undetermined barcode */
                        throw ConfigurationError("barcode " + order[i].second->first + " holds a degenerate nucleotide; only A, C, G and T are supported on this path");
                    }
                    d.barcode[static_cast< size_t >(i) * d.nucleotide_cardinality + d.segment_offset[s] + j] = code;
                }
                flat += text;
            }
            if(d.tiled() && !unique.insert(flat).second) { throw ConfigurationError("duplicate barcode sequence " + flat); }
        }
    }

    if(d.tiled()) {
        if(d.nucleotide_cardinality < 1) { throw ConfigurationError("decoder has no nucleotides to decode"); }
        if(d.nucleotide_cardinality > PHQ_MAX_NUCLEOTIDES) {
            throw ConfigurationError("nucleotide cardinality " + std::to_string(d.nucleotide_cardinality) + " exceeds the " + std::to_string(PHQ_MAX_NUCLEOTIDES) + " supported on this path");
        }
        if(d.barcode_cardinality < 1) { throw ConfigurationError("decoder has an empty codec"); }
    }
    if(d.algorithm == PHQ_PAMLD) {
        /* transcode.cpp:1540-1565 */
        if(d.confidence_threshold < 0 || d.confidence_threshold > 1) { throw ConfigurationError("confidence threshold value " + std::to_string(d.confidence_threshold) + " not between 0 and 1"); }
        if(d.noise < 0 || d.noise > 1) { throw ConfigurationError("noise value " + std::to_string(d.noise) + " not between 0 and 1"); }
    }
    if(d.algorithm == PHQ_MDD) {
        const Json* tolerance(v.find("distance tolerance"));
        if(tolerance == NULL || !tolerance->is_array() || static_cast< int32_t >(tolerance->items().size()) != d.segment_cardinality) {
            throw ConfigurationError("distance tolerance cardinality inconsistant with " + std::to_string(d.segment_cardinality) + " barcode segment cardinality");
        }
        for(const auto& t : tolerance->items()) {
            /* json.cpp:391 narrows every element through uint8_t */
            d.distance_tolerance.push_back(static_cast< int32_t >(static_cast< uint8_t >(t.as_int())));
        }
        if(d.quality_masking_threshold < 0 || d.quality_masking_threshold > 255) { throw ConfigurationError("quality masking threshold element is not an 8 bit unsigned integer"); }
    }
    return d;
}

/* chain order: sample, molecular[], cellular[] (transcode.h:51-60) */
inline std::vector< DecoderSpec > parse_compiled_job(const Json& job) {
    if(!job.is_object()) { throw ConfigurationError("job element must be a dictionary"); }
    std::vector< DecoderSpec > chain;
    const char* name[3] = { "sample", "molecular", "cellular" };
    const int32_t topic[3] = { PHQ_SAMPLE, PHQ_MOLECULAR, PHQ_CELLULAR };
    for(int t(0); t < 3; ++t) {
        std::vector< const Json* > elements(topic_elements(job, name[t]));
        if(t == 0 && elements.size() > 1) { throw ConfigurationError("only one sample decoder is allowed"); }
        for(const Json* e : elements) {
            try {
                chain.push_back(parse_compiled_decoder(*e, topic[t]));
            } catch(ConfigurationError& error) {
                throw ConfigurationError(std::string(name[t]) + " decoder : " + (error.what() + strlen("Configuration error : ")));
            }
        }
    }
    if(chain.empty()) { throw ConfigurationError("job declares no decoder"); }
    return chain;
}

/* ------------------------------------------------------------------ job level JSON semantics of the reference */

/* merge_json_value (json.cpp:780-803): members of `base` the ontology lacks are copied in, common objects recurse, the ontology wins */
inline void merge_json(const Json& base, Json& ontology) {
    if(base.is_null()) { return; }
    if(ontology.is_null()) { ontology = base; return; }
    if(!base.is_object()) { return; }
    if(!ontology.is_object()) { throw ConfigurationError("element is not a dictionary"); }
    for(const auto& m : base.members()) {
        Json* found(ontology.find(m.first));
        if(found != NULL) {
            try { merge_json(m.second, *found); }
            catch(ConfigurationError& e) { throw ConfigurationError(m.first + " " + (e.what() + strlen("Configuration error : "))); }
        } else {
            ontology.set(m.first, m.second);
        }
    }
}
/* project_json_value (json.cpp:804-832): the keys of `base`, valued from the ontology where it has them */
inline Json project_json(const Json& base, const Json& ontology) {
    Json container;
    if(!base.is_null() && !ontology.is_null() && base.is_object()) {
        if(ontology.is_object()) {
            container = Json::object();
            for(const auto& m : base.members()) {
                const Json* element(ontology.find(m.first));
                container.set(m.first, element != NULL ? project_json(m.second, *element) : m.second);
            }
        } else if(ontology.is_array()) {
            container = Json::array();
            for(const auto& e : ontology.items()) { container.push(project_json(base, e)); }
        }
    }
    if(!ontology.is_null() && container.is_null()) { container = ontology; }
    return container;
}
/* clean_json_value (json.cpp:833-874): false, empty strings, empty containers and nulls disappear (in place) */
inline void clean_json(Json& v) {
    if(v.is_bool()) { if(!v.as_bool()) { v = Json(); } }
    else if(v.is_string()) { if(v.as_string().empty()) { v = Json(); } }
    else if(v.is_object()) {
        std::vector< Json::Member >& members(v.members());
        for(auto& m : members) { clean_json(m.second); }
        members.erase(std::remove_if(members.begin(), members.end(), [](const Json::Member& m) { return m.second.is_null(); }), members.end());
        if(members.empty()) { v = Json(); }
    } else if(v.is_array()) {
        std::vector< Json >& items(v.items());
        for(auto& e : items) { clean_json(e); }
        items.erase(std::remove_if(items.begin(), items.end(), [](const Json& e) { return e.is_null(); }), items.end());
        if(items.empty()) { v = Json(); }
    }
}

/*  Transcode::apply_inheritance (transcode.cpp:328-442): `base` references between the decoders of the job's
    `decoder` repository are resolved from the roots down, then the sample / molecular / cellular decoders inherit
    from the repository, and the repository is dropped. */
inline int32_t inheritance_depth(const std::string& key, Json& repository, std::map< std::string, int32_t >& depth) {
    auto known(depth.find(key));
    if(known != depth.end()) { return known->second; }
    Json* value(repository.find(key));
    if(value == NULL || value->is_null()) { throw ConfigurationError("referencing an unknown parent " + key); }
    int32_t d(0);
    const std::string base(value->is_object() ? get_string(*value, "base") : std::string());
    if(!base.empty()) {
        if(base == key) { throw ConfigurationError(key + " references itself as parent"); }
        d = inheritance_depth(base, repository, depth) + 1;
    }
    depth[key] = d;
    return d;
}
inline void apply_decoder_inheritance(Json& value, const Json* repository) {
    if(!value.is_object()) { return; }
    const std::string base(get_string(value, "base"));
    if(!base.empty() && repository != NULL) {
        const Json* parent(repository->find(base));
        if(parent == NULL) { throw ConfigurationError("reference to an unknown base " + base); }
        merge_json(*parent, value);
    }
    value.erase("base");
    clean_json(value);
}
inline void apply_inheritance(Json& job) {
    Json* repository(job.find("decoder"));
    if(repository != NULL && repository->is_object()) {
        std::map< std::string, int32_t > depth;
        int32_t deepest(0);
        for(const auto& m : repository->members()) {
            if(!m.second.is_null()) { deepest = std::max(deepest, inheritance_depth(m.first, *repository, depth)); }
        }
        for(int32_t level(1); level <= deepest; ++level) {
            for(const auto& record : depth) {
                if(record.second != level) { continue; }
                Json* value(repository->find(record.first));
                const std::string base(get_string(*value, "base"));
                if(!base.empty()) {
                    const Json parent(repository->at(base));
                    merge_json(parent, *value);
                    value->erase("base");
                }
            }
        }
    }
    const char* name[3] = { "sample", "molecular", "cellular" };
    for(int t(0); t < 3; ++t) {
        Json* e(job.find(name[t]));
        if(e == NULL || e->is_null()) { continue; }
        try {
            if(e->is_object()) { apply_decoder_inheritance(*e, repository); }
            else if(e->is_array()) { for(auto& d : e->items()) { apply_decoder_inheritance(d, repository); } }
        } catch(ConfigurationError& error) {
            throw ConfigurationError(std::string(name[t]) + " decoder : " + (error.what() + strlen("Configuration error : ")));
        }
    }
    if(repository != NULL) { job.erase("decoder"); }
}

/* the "<topic>:decoder" and "<topic>:barcode" projections (configuration.json:423-501) */
inline Json decoder_template(const char* topic) {
    Json t(Json::object());
    const bool sample(strcmp(topic, "sample") == 0);
    if(sample) { for(const char* k : { "CN", "DT", "LB", "PG", "PI", "PL", "PM", "SM" }) { t.set(k, Json()); } }
    t.set("algorithm", Json::string(strcmp(topic, "molecular") == 0 ? "naive" : "pamld"));
    t.set("codec", Json());
    t.set("confidence threshold", Json::number(0.95));
    t.set("corrected quality", Json());
    t.set("distance tolerance", Json());
    if(sample) { t.set("flowcell id", Json()); t.set("flowcell lane number", Json()); }
    t.set("high quality distance threshold", Json::integer(0));
    t.set("high quality threshold", Json::integer(30));
    t.set("noise", Json::number(0.01));
    t.set("quality masking threshold", Json::integer(0));
    t.set("segment cardinality", Json::integer(0));
    t.set("undetermined", Json());
    return t;
}
inline Json barcode_template(const char* topic) {
    Json t(Json::object());
    const bool sample(strcmp(topic, "sample") == 0);
    if(sample) { for(const char* k : { "CN", "DT", "LB", "PG", "PI", "PL", "PM", "SM" }) { t.set(k, Json()); } }
    t.set("algorithm", Json());
    t.set("concentration", Json::integer(1));
    if(sample) { t.set("flowcell id", Json()); t.set("flowcell lane number", Json()); }
    t.set("segment cardinality", Json());
    return t;
}
/* Transcode::infer_PU / infer_ID (transcode.cpp:1224-1260): PU = [flowcell id:[lane:]]barcode, ID = PU, unless given */
inline void infer_platform_unit(Json& container, bool undetermined) {
    if(get_string(container, "PU").empty()) {
        std::string suffix;
        if(!undetermined) {
            const Json* barcode(container.find("barcode"));
            if(barcode != NULL && barcode->is_array()) { for(const auto& segment : barcode->items()) { suffix += segment.as_string(); } }
        } else { suffix = "undetermined"; }
        if(!suffix.empty()) {
            std::string buffer(get_string(container, "flowcell id"));
            if(!buffer.empty()) {
                buffer.push_back(':');
                const Json* lane(container.find("flowcell lane number"));
                if(lane != NULL && lane->is_number()) { buffer += std::to_string(lane->as_int()); buffer.push_back(':'); }
            }
            buffer += suffix;
            container.set("PU", Json::string(buffer));
        }
    }
    if(get_string(container, "ID").empty()) {
        const std::string unit(get_string(container, "PU"));
        if(!unit.empty()) { container.set("ID", Json::string(unit)); }
    }
}

/* ------------------------------------------------------------------ job compile (decoder sections only) */

/* WordMetric::find_shannon_bound over the distinct words of one segment (metric.h:87-111) */
inline int32_t shannon_bound(const std::set< std::string >& words, int32_t length) {
    if(words.empty()) { return 0; }
    std::vector< const std::string* > w;
    for(const auto& s : words) { w.push_back(&s); }
    int32_t minimum(length);
    for(size_t i(0); i < w.size(); ++i) {
        for(size_t j(i + 1); j < w.size(); ++j) {
            int32_t distance(0);
            const std::string& a(*w[i]);
            const std::string& b(*w[j]);
            for(size_t k(0); k < a.size(); ++k) {
                if(a[k] != b[k]) { if(++distance >= minimum) { break; } }
            }
            if(distance < minimum) { minimum = distance; }
        }
    }
    return (minimum - 1) / 2;
}

inline Json compile_decoder(Json value, const char* topic, int32_t index, const Json& default_decoder, const Json& default_barcode) {
    if(!value.is_object()) { throw ConfigurationError("decoder element must be a dictionary"); }
    /* Transcode::compile_decoder (transcode.cpp:937-966): overlay on the default decoder (the "<topic>:decoder"
       projection valued from the job root, transcode.cpp:769-790), then clean */
    value.set("index", Json::integer(index));
    merge_json(default_decoder, value);
    clean_json(value);
    /* the default barcode of this codec: the "<topic>:barcode" projection valued from the decoder */
    Json default_codec_barcode(project_json(default_barcode, value));
    clean_json(default_codec_barcode);
    if(default_codec_barcode.is_null()) { default_codec_barcode = Json::object(); }
    const int32_t algorithm(algorithm_from_string(get_string(value, "algorithm")));

    std::vector< int32_t > barcode_length;
    int32_t segment_cardinality(0);
    int32_t nucleotide_cardinality(0);
    if(value.has("transform")) {
        std::vector< TransformSpec > transform;
        parse_rule(value.at("transform"), transform, segment_cardinality);
        Json compiled_transform(Json::object());
        compiled_transform.set("token", value.at("transform").at("token"));
        {
            /* re-encode the knit the way encode_key_value(list<Transform>) does (transform.cpp:222-248) */
            Json knit(Json::array());
            std::string current;
            int32_t segment(0);
            for(const auto& t : transform) {
                if(t.output_segment_index != segment) { knit.push(Json::string(current)); current.clear(); ++segment; }
                if(!current.empty()) { current.push_back(':'); }
                if(t.reverse_complement) { current.push_back('~'); }
                current += std::to_string(t.token_index);
            }
            knit.push(Json::string(current));
            compiled_transform.set("knit", knit);
        }
        value.set("transform", compiled_transform);
        barcode_length.assign(segment_cardinality, 0);
        for(const auto& t : transform) {
            if(t.empty()) { throw ConfigurationError("token " + std::to_string(t.token_index) + " is empty"); }
            if(!t.constant()) { throw ConfigurationError("token " + std::to_string(t.token_index) + " is not fixed width"); }
            barcode_length[t.output_segment_index] += t.length();
            nucleotide_cardinality += t.length();
        }
        value.set("segment cardinality", Json::integer(segment_cardinality));
        value.set("nucleotide cardinality", Json::integer(nucleotide_cardinality));
        Json lengths(Json::array());
        for(int32_t n : barcode_length) { lengths.push(Json::integer(n)); }
        value.set("barcode length", lengths);

        const double lower_bound(1.0 / double(pow(4, (nucleotide_cardinality))));       /* transcode.cpp:855-863 */
        if(value.has("random barcode probability")) {
            if(get_double(value, "random barcode probability") < lower_bound) { throw ConfigurationError("random barcode probability is smaller than lower bound"); }
        } else {
            value.set("random barcode probability", Json::number(lower_bound));
        }
    } else if(algorithm != PHQ_PASSTHROUGH) {
        throw ConfigurationError("no element transform found");
    }

    const double noise(get_double(value, "noise"));
    {
        Json undetermined(value.has("undetermined") ? value.at("undetermined") : Json::object());
        merge_json(default_codec_barcode, undetermined);
        Json barcode(Json::array());
        for(int32_t n : barcode_length) { barcode.push(Json::string(std::string(static_cast< size_t >(n), '='))); }
        undetermined.set("barcode", barcode);
        undetermined.set("segment cardinality", Json::integer(segment_cardinality));
        undetermined.set("index", Json::integer(0));
        infer_platform_unit(undetermined, true);
        undetermined.set("concentration", Json::number(noise));
        value.set("undetermined", undetermined);
    }

    if(value.has("codec")) {
        Json& codec(*value.find("codec"));     /* compiled in place: a whitelist codec is hundreds of megabytes of JSON */
        if(!codec.is_object()) { throw ConfigurationError("codec element must be a dictionary"); }
        codec.sort_keys();                                                              /* transcode.cpp:350, json.cpp:875-893 */
        int32_t barcode_index(1);
        double total_concentration(0);
        std::set< std::string > unique;
        std::vector< std::set< std::string > > segment_words(static_cast< size_t >(segment_cardinality));
        for(auto& record : codec.members()) {
            Json& element(record.second);
            if(!element.is_object()) { throw ConfigurationError("codec element " + record.first + " must be a dictionary"); }
            merge_json(default_codec_barcode, element);
            const Json* barcode(element.find("barcode"));
            if(barcode == NULL || !barcode->is_array()) { throw ConfigurationError("barcode " + record.first + " has no barcode array"); }
            if(static_cast< int32_t >(barcode->items().size()) != segment_cardinality) {
                throw ConfigurationError("expected " + std::to_string(segment_cardinality) + " segments but found " + std::to_string(barcode->items().size()) + " in barcode " + record.first);
            }
            std::string flat, hyphenated;
            for(int32_t s(0); s < segment_cardinality; ++s) {
                const std::string& text(barcode->items()[s].as_string());
                if(static_cast< int32_t >(text.size()) != barcode_length[s]) {
                    throw ConfigurationError("expected " + std::to_string(barcode_length[s]) + " but found " + std::to_string(text.size()) + " nucleotides in segment " + std::to_string(s) + " of barcode " + record.first);
                }
                flat += text;
                if(s) { hyphenated += "-"; }
                hyphenated += text;
                segment_words[s].insert(text);
            }
            if(!unique.insert(flat).second) { throw ConfigurationError("duplicate barcode sequence " + flat); }
            element.set("index", Json::integer(barcode_index++));
            element.set("segment cardinality", Json::integer(segment_cardinality));
            element.set("BC", Json::string(hyphenated));
            infer_platform_unit(element, false);
            double concentration(element.has("concentration") ? get_double(element, "concentration") : 1.0);    /* <topic>:barcode projection */
            if(!(concentration >= 0)) { throw ConfigurationError("barcode concentration must be a positive number"); }
            element.set("concentration", Json::number(concentration));
            total_concentration += concentration;
        }
        value.set("barcode cardinality", Json::integer(barcode_index));
        if(!(total_concentration > 0)) { throw ConfigurationError("total pool concentration is not a positive number"); }
        const double factor((1.0 - noise) / total_concentration);                       /* transcode.cpp:1024-1029 */
        for(auto& record : codec.members()) {
            record.second.set("concentration", Json::number(get_double(record.second, "concentration") * factor));
        }

        /* CodecMetric::compile_barcode_tolerance (metric.h:216-242). The pairwise scan is quadratic in the
           distinct words of a segment; it is what the reference does, and only MDD consumes the result. */
        const bool wanted(algorithm == PHQ_MDD || barcode_index <= 4097);
        if(wanted) {
            Json bound(Json::array());
            std::vector< int32_t > bounds;
            for(int32_t s(0); s < segment_cardinality; ++s) {
                bounds.push_back(shannon_bound(segment_words[s], barcode_length[s]));
                bound.push(Json::integer(bounds.back()));
            }
            value.set("shannon bound", bound);
            if(value.has("distance tolerance")) {
                const Json& tolerance(value.at("distance tolerance"));
                if(!tolerance.is_array() || static_cast< int32_t >(tolerance.items().size()) != segment_cardinality) {
                    throw ConfigurationError(std::to_string(tolerance.is_array() ? tolerance.items().size() : 0) + " distance tolerance cardinality inconsistant with " + std::to_string(segment_cardinality) + " barcode segment cardinality");
                }
                for(int32_t s(0); s < segment_cardinality; ++s) {
                    if(tolerance.items()[s].as_int() > bounds[s]) {
                        throw ConfigurationError("barcode tolerance for segment " + std::to_string(s) + " is higher than shannon bound " + std::to_string(bounds[s]));
                    }
                }
            } else {
                value.set("distance tolerance", bound);
            }
        }
    }
    /* validate_decoder (transcode.cpp:1540-1565) */
    const double threshold(get_double(value, "confidence threshold"));
    if(threshold < 0 || threshold > 1) { throw ConfigurationError("confidence threshold value " + std::to_string(threshold) + " not between 0 and 1"); }
    if(noise < 0 || noise > 1) { throw ConfigurationError("noise value " + std::to_string(noise) + " not between 0 and 1"); }
    value.sort_keys();
    return value;
}

inline Json compile_job(const Json& directive) {
    if(!directive.is_object()) { throw ConfigurationError("job element must be a dictionary"); }
    Json job(directive);
    apply_inheritance(job);
    /* the job level defaults this path consumes (configuration.json "default") */
    if(!job.has("corrected quality")) { job.set("corrected quality", Json::integer(30)); }
    Json out(Json::object());
    const char* name[3] = { "sample", "molecular", "cellular" };
    for(int t(0); t < 3; ++t) {
        Json* e(job.find(name[t]));
        if(e == NULL || e->is_null()) { continue; }
        /* Transcode::compile_topic (transcode.cpp:769-823): defaults are the projections valued from the job root. The
           projection's keys do not include the decoders themselves, so they can be moved out of the (local) job. */
        const Json default_decoder(project_json(decoder_template(name[t]), job));
        const Json default_barcode(barcode_template(name[t]));
        try {
            if(e->is_object()) {
                out.set(name[t], compile_decoder(std::move(*e), name[t], 0, default_decoder, default_barcode));
            } else if(e->is_array()) {
                Json list(Json::array());
                int32_t index(0);
                for(auto& d : e->items()) { list.push(compile_decoder(std::move(d), name[t], index++, default_decoder, default_barcode)); }
                out.set(name[t], std::move(list));
            } else { throw ConfigurationError("decoder element must be a dictionary or an array"); }
        } catch(ConfigurationError& error) {
            throw ConfigurationError(std::string(name[t]) + " decoder : " + (error.what() + strlen("Configuration error : ")));
        }
    }
    /*  Transcode::find_multiplexing_decoder (transcode.cpp:1120-1222): the decoder that mentions `output` (on itself, its
        undetermined element or a codec element) routes reads to channels; when none does, the sample decoder. */
    {
        auto mentions_output = [](const Json& d) {
            if(d.find("output") != NULL) { return true; }
            const Json* u(d.find("undetermined"));
            if(u != NULL && u->is_object() && u->find("output") != NULL) { return true; }
            const Json* codec(d.find("codec"));
            if(codec != NULL && codec->is_object()) {
                for(const auto& record : codec->members()) { if(record.second.is_object() && record.second.find("output") != NULL) { return true; } }
            }
            return false;
        };
        std::vector< Json* > candidate;
        for(int t(0); t < 3; ++t) {
            Json* e(out.find(name[t]));
            if(e == NULL) { continue; }
            if(e->is_object()) { if(mentions_output(*e)) { candidate.push_back(e); } }
            else { for(auto& d : e->items()) { if(mentions_output(d)) { candidate.push_back(&d); } } }
        }
        if(candidate.size() > 1) { throw ConfigurationError("multiple multiplexing classifier candidates found"); }
        Json* chosen(candidate.empty() ? out.find("sample") : candidate.front());
        if(chosen != NULL && chosen->is_object()) { chosen->set("multiplexing classifier", Json::boolean(true)); }
    }
    return out;
}

/*  Job::load_instruction_with_import (job.cpp:160-224): the document at `path` with the documents its `import`
    list names merged underneath it, depth first; an import path is relative to the importing document and a
    document is only visited once. */
inline std::string canonical_path(const std::string& path) {
    char resolved[4096];
    return realpath(path.c_str(), resolved) != NULL ? std::string(resolved) : path;
}
inline Json load_job_with_import(const std::string& given, std::set< std::string >& visited) {
    const std::string path(canonical_path(given));      /* the same document reached through another spelling is the same document */
    FILE* const file(fopen(path.c_str(), "rb"));
    if(file == NULL) { throw ConfigurationError("unable to read instruction file from " + path); }
    std::string text;
    char buffer[1 << 16];
    size_t got;
    while((got = fread(buffer, 1, sizeof(buffer), file)) > 0) { text.append(buffer, got); }
    fclose(file);
    Json document(Json::parse(text));
    if(!document.is_object()) { throw ConfigurationError("instruction " + path + " is not a dictionary"); }
    visited.insert(path);
    const Json* import(document.find("import"));
    if(import != NULL && import->is_array()) {
        const size_t slash(path.find_last_of('/'));
        const std::string directory(slash == std::string::npos ? std::string() : path.substr(0, slash + 1));
        Json aggregated;
        for(const auto& record : import->items()) {
            std::string target(record.as_string());
            if(target.empty()) { continue; }
            if(target[0] != '/') { target = directory + target; }
            target = canonical_path(target);
            if(visited.count(target)) { continue; }
            Json imported(load_job_with_import(target, visited));
            merge_json(aggregated, imported);
            aggregated = imported;
        }
        merge_json(aggregated, document);
    }
    document.erase("import");
    return document;
}

}   /* namespace phq */
#endif
