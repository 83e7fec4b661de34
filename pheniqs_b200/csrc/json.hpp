/*  json.hpp — a small JSON document model for the decoder configuration.

    The reference reads its configuration with rapidjson (json.h / json.cpp); the hot path
    only needs to read compiled decoder ontologies and to write them back, so this is a
    self-contained reader / writer with the two behaviours of the reference that matter
    here: object members keep insertion order and can be key-sorted (json.cpp:875-893), and
    numbers are written with at most `precision` decimal places.
*/
#ifndef PHQ_JSON_HPP
#define PHQ_JSON_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace phq {

struct JsonError : public std::runtime_error {
    explicit JsonError(const std::string& what) : std::runtime_error(what) {}
};

class Json {
    public:
        enum Type { Null, Bool, Number, String, Array, Object };
        typedef std::pair< std::string, Json > Member;

        Json() : type_(Null), boolean_(false), number_(0), integral_(false) {}
        /* the key index is a cache: never copied; moves are noexcept so that containers of values move on growth */
        Json(const Json& o) : type_(o.type_), boolean_(o.boolean_), number_(o.number_), integral_(o.integral_), string_(o.string_), items_(o.items_), members_(o.members_) {}
        Json(Json&& o) noexcept : type_(o.type_), boolean_(o.boolean_), number_(o.number_), integral_(o.integral_), string_(std::move(o.string_)),
            items_(std::move(o.items_)), members_(std::move(o.members_)), index_(std::move(o.index_)) { o.type_ = Null; }
        Json& operator=(const Json& o) { if(this != &o) { Json copy(o); *this = std::move(copy); } return *this; }
        Json& operator=(Json&& o) noexcept {
            if(this != &o) {
                type_ = o.type_; boolean_ = o.boolean_; number_ = o.number_; integral_ = o.integral_;
                string_ = std::move(o.string_); items_ = std::move(o.items_); members_ = std::move(o.members_); index_ = std::move(o.index_);
                o.type_ = Null;
            }
            return *this;
        }
        static Json boolean(bool v) { Json j; j.type_ = Bool; j.boolean_ = v; return j; }
        static Json number(double v) { Json j; j.type_ = Number; j.number_ = v; j.integral_ = false; return j; }
        static Json integer(int64_t v) { Json j; j.type_ = Number; j.number_ = static_cast< double >(v); j.integral_ = true; return j; }
        static Json string(const std::string& v) { Json j; j.type_ = String; j.string_ = v; return j; }
        static Json array() { Json j; j.type_ = Array; return j; }
        static Json object() { Json j; j.type_ = Object; return j; }

        Type type() const { return type_; }
        bool is_null() const { return type_ == Null; }
        bool is_bool() const { return type_ == Bool; }
        bool is_number() const { return type_ == Number; }
        bool is_string() const { return type_ == String; }
        bool is_array() const { return type_ == Array; }
        bool is_object() const { return type_ == Object; }

        bool as_bool() const { expect(Bool, "boolean"); return boolean_; }
        double as_double() const { expect(Number, "number"); return number_; }
        int64_t as_int() const {
            expect(Number, "number");
            if(number_ != std::floor(number_)) { throw JsonError("expected an integer"); }
            return static_cast< int64_t >(number_);
        }
        const std::string& as_string() const { expect(String, "string"); return string_; }
        const std::vector< Json >& items() const { expect(Array, "array"); return items_; }
        std::vector< Json >& items() { expect(Array, "array"); return items_; }
        const std::vector< Member >& members() const { expect(Object, "object"); return members_; }
        std::vector< Member >& members() { expect(Object, "object"); index_.reset(); return members_; }

        const Json* find(const std::string& key) const {
            if(type_ != Object) { return NULL; }
            const long at(position(key));
            return at < 0 ? NULL : &members_[static_cast< size_t >(at)].second;
        }
        Json* find(const std::string& key) {
            if(type_ != Object) { return NULL; }
            const long at(position(key));
            return at < 0 ? NULL : &members_[static_cast< size_t >(at)].second;
        }
        bool has(const std::string& key) const { const Json* v(find(key)); return v != NULL && !v->is_null(); }
        const Json& at(const std::string& key) const {
            const Json* v(find(key));
            if(v == NULL) { throw JsonError("element " + key + " not found"); }
            return *v;
        }
        /* replace or append, like encode_key_value (RemoveMember + AddMember) */
        void set(const std::string& key, const Json& value) {
            expect(Object, "object");
            const long at(position(key));
            if(at >= 0) { members_[static_cast< size_t >(at)].second = value; return; }
            members_.emplace_back(key, value);
            if(index_ != nullptr) { index_->emplace(key, members_.size() - 1); }
        }
        void erase(const std::string& key) {
            expect(Object, "object");
            index_.reset();
            members_.erase(std::remove_if(members_.begin(), members_.end(), [&](const Member& m) { return m.first == key; }), members_.end());
        }
        void set(const std::string& key, Json&& value) {
            expect(Object, "object");
            const long at(position(key));
            if(at >= 0) { members_[static_cast< size_t >(at)].second = std::move(value); return; }
            members_.emplace_back(key, std::move(value));
            if(index_ != nullptr) { index_->emplace(key, members_.size() - 1); }
        }
        void push(const Json& value) { expect(Array, "array"); items_.push_back(value); }
        void push(Json&& value) { expect(Array, "array"); items_.push_back(std::move(value)); }

        /* json.cpp:875-893: recursive key sort, byte order */
        void sort_keys() {
            if(type_ == Object) {
                for(auto& m : members_) { m.second.sort_keys(); }
                index_.reset();
                std::stable_sort(members_.begin(), members_.end(), [](const Member& a, const Member& b) { return a.first < b.first; });
            } else if(type_ == Array) {
                for(auto& e : items_) { e.sort_keys(); }
            }
        }

        static Json parse(const std::string& text) {
            Parser p(text);
            Json value(p.parse_value());
            p.skip();
            if(!p.done()) { throw JsonError("trailing characters after JSON document at offset " + std::to_string(p.at)); }
            return value;
        }
        std::string dump(int precision = 17, int indent = 4) const {
            std::string out;
            write(out, precision, indent, 0);
            return out;
        }

    private:
        Type type_;
        bool boolean_;
        double number_;
        bool integral_;
        std::string string_;
        std::vector< Json > items_;
        std::vector< Member > members_;
        /* key -> position, built on the first lookup of a large object (a 737 K barcode codec is one object); a cache
           that is allocated on demand and dropped whenever positions may change */
        mutable std::unique_ptr< std::unordered_map< std::string, size_t > > index_;
        long position(const std::string& key) const {
            if(members_.size() < 32) {
                for(size_t i(0); i < members_.size(); ++i) { if(members_[i].first == key) { return static_cast< long >(i); } }
                return -1;
            }
            if(index_ == nullptr) {
                index_.reset(new std::unordered_map< std::string, size_t >());
                index_->reserve(members_.size() * 2);
                for(size_t i(0); i < members_.size(); ++i) { index_->emplace(members_[i].first, i); }     /* first occurrence wins */
            }
            const auto found(index_->find(key));
            return found == index_->end() ? -1 : static_cast< long >(found->second);
        }

        void expect(Type t, const char* name) const {
            if(type_ != t) { throw JsonError(std::string("expected a JSON ") + name); }
        }

        struct Parser {
            const std::string& s;
            size_t at;
            explicit Parser(const std::string& text) : s(text), at(0) {}
            bool done() const { return at >= s.size(); }
            void skip() { while(at < s.size() && (s[at] == ' ' || s[at] == '\t' || s[at] == '\n' || s[at] == '\r')) { ++at; } }
            char peek() { skip(); if(done()) { throw JsonError("unexpected end of JSON"); } return s[at]; }
            void consume(char c) {
                if(peek() != c) { throw JsonError(std::string("expected '") + c + "' at offset " + std::to_string(at)); }
                ++at;
            }
            bool literal(const char* word) {
                size_t n(strlen(word));
                if(s.compare(at, n, word) == 0) { at += n; return true; }
                return false;
            }
            Json parse_value() {
                char c(peek());
                if(c == '{') { return parse_object(); }
                if(c == '[') { return parse_array(); }
                if(c == '"') { return Json::string(parse_string()); }
                if(literal("true")) { return Json::boolean(true); }
                if(literal("false")) { return Json::boolean(false); }
                if(literal("null")) { return Json(); }
                return parse_number();
            }
            Json parse_object() {
                Json o(Json::object());
                consume('{');
                if(peek() == '}') { ++at; return o; }
                while(true) {
                    if(peek() != '"') { throw JsonError("expected a member name at offset " + std::to_string(at)); }
                    std::string key(parse_string());
                    consume(':');
                    o.set(key, parse_value());
                    char c(peek());
                    ++at;
                    if(c == '}') { break; }
                    if(c != ',') { throw JsonError("expected ',' or '}' at offset " + std::to_string(at - 1)); }
                }
                return o;
            }
            Json parse_array() {
                Json a(Json::array());
                consume('[');
                if(peek() == ']') { ++at; return a; }
                while(true) {
                    a.push(parse_value());
                    char c(peek());
                    ++at;
                    if(c == ']') { break; }
                    if(c != ',') { throw JsonError("expected ',' or ']' at offset " + std::to_string(at - 1)); }
                }
                return a;
            }
            static void append_utf8(std::string& out, uint32_t cp) {
                if(cp < 0x80) { out.push_back(static_cast< char >(cp)); }
                else if(cp < 0x800) { out.push_back(static_cast< char >(0xC0 | (cp >> 6))); out.push_back(static_cast< char >(0x80 | (cp & 0x3F))); }
                else if(cp < 0x10000) { out.push_back(static_cast< char >(0xE0 | (cp >> 12))); out.push_back(static_cast< char >(0x80 | ((cp >> 6) & 0x3F))); out.push_back(static_cast< char >(0x80 | (cp & 0x3F))); }
                else { out.push_back(static_cast< char >(0xF0 | (cp >> 18))); out.push_back(static_cast< char >(0x80 | ((cp >> 12) & 0x3F))); out.push_back(static_cast< char >(0x80 | ((cp >> 6) & 0x3F))); out.push_back(static_cast< char >(0x80 | (cp & 0x3F))); }
            }
            uint32_t parse_hex4() {
                if(at + 4 > s.size()) { throw JsonError("truncated \\u escape"); }
                uint32_t v(0);
                for(int i(0); i < 4; ++i) {
                    char c(s[at++]);
                    v <<= 4;
                    if(c >= '0' && c <= '9') { v |= c - '0'; }
                    else if(c >= 'a' && c <= 'f') { v |= c - 'a' + 10; }
                    else if(c >= 'A' && c <= 'F') { v |= c - 'A' + 10; }
                    else { throw JsonError("illegal \\u escape"); }
                }
                return v;
            }
            std::string parse_string() {
                consume('"');
                std::string out;
                while(true) {
                    if(done()) { throw JsonError("unterminated string"); }
                    char c(s[at++]);
                    if(c == '"') { break; }
                    if(c == '\\') {
                        if(done()) { throw JsonError("unterminated escape"); }
                        char e(s[at++]);
                        switch(e) {
                            case '"': out.push_back('"'); break;
                            case '\\': out.push_back('\\'); break;
                            case '/': out.push_back('/'); break;
                            case 'b': out.push_back('\b'); break;
                            case 'f': out.push_back('\f'); break;
                            case 'n': out.push_back('\n'); break;
                            case 'r': out.push_back('\r'); break;
                            case 't': out.push_back('\t'); break;
                            case 'u': {
                                uint32_t cp(parse_hex4());
                                if(cp >= 0xD800 && cp < 0xDC00 && at + 1 < s.size() && s[at] == '\\' && s[at + 1] == 'u') {
                                    at += 2;
                                    uint32_t low(parse_hex4());
                                    cp = 0x10000 + ((cp - 0xD800) << 10) + (low - 0xDC00);
                                }
                                append_utf8(out, cp);
                                break;
                            }
                            default: throw JsonError("illegal escape in string");
                        }
                    } else { out.push_back(c); }
                }
                return out;
            }
            Json parse_number() {
                skip();
                const char* begin(s.c_str() + at);
                char* end(NULL);
                double v(strtod(begin, &end));
                if(end == begin) { throw JsonError("illegal JSON value at offset " + std::to_string(at)); }
                bool integral(true);
                for(const char* p(begin); p < end; ++p) { if(*p == '.' || *p == 'e' || *p == 'E') { integral = false; } }
                at += static_cast< size_t >(end - begin);
                Json j(Json::number(v));
                j.integral_ = integral;
                return j;
            }
        };

        static void write_string(std::string& out, const std::string& s) {
            out.push_back('"');
            for(unsigned char c : s) {
                switch(c) {
                    case '"': out += "\\\""; break;
                    case '\\': out += "\\\\"; break;
                    case '\n': out += "\\n"; break;
                    case '\r': out += "\\r"; break;
                    case '\t': out += "\\t"; break;
                    default:
                        if(c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; }
                        else { out.push_back(static_cast< char >(c)); }
                }
            }
            out.push_back('"');
        }
        static void write_number(std::string& out, double v, bool integral, int precision) {
            char b[64];
            if(integral && std::fabs(v) < 9.0e15) {
                snprintf(b, sizeof(b), "%lld", static_cast< long long >(v));
                out += b;
                return;
            }
            if(!std::isfinite(v)) { out += "null"; return; }
            /* shortest representation that round-trips, then capped at `precision` decimal places */
            snprintf(b, sizeof(b), "%.17g", v);
            for(int digits(1); digits < 17; ++digits) {
                char t[64];
                snprintf(t, sizeof(t), "%.*g", digits, v);
                if(strtod(t, NULL) == v) { memcpy(b, t, sizeof(t)); break; }
            }
            std::string text(b);
            if(text.find('e') == std::string::npos && text.find('E') == std::string::npos) {
                size_t dot(text.find('.'));
                if(dot == std::string::npos) { text += ".0"; }
                else if(precision >= 0 && text.size() - dot - 1 > static_cast< size_t >(precision)) {
                    text.resize(dot + 1 + static_cast< size_t >(precision));
                    while(text.size() > dot + 2 && text.back() == '0') { text.pop_back(); }
                }
            }
            out += text;
        }
        void write(std::string& out, int precision, int indent, int depth) const {
            switch(type_) {
                case Null: out += "null"; break;
                case Bool: out += boolean_ ? "true" : "false"; break;
                case Number: write_number(out, number_, integral_, precision); break;
                case String: write_string(out, string_); break;
                case Array: {
                    if(items_.empty()) { out += "[]"; break; }
                    out += "[";
                    for(size_t i(0); i < items_.size(); ++i) {
                        if(i) { out += ","; }
                        newline(out, indent, depth + 1);
                        items_[i].write(out, precision, indent, depth + 1);
                    }
                    newline(out, indent, depth);
                    out += "]";
                    break;
                }
                case Object: {
                    if(members_.empty()) { out += "{}"; break; }
                    out += "{";
                    for(size_t i(0); i < members_.size(); ++i) {
                        if(i) { out += ","; }
                        newline(out, indent, depth + 1);
                        write_string(out, members_[i].first);
                        out += indent > 0 ? ": " : ":";
                        members_[i].second.write(out, precision, indent, depth + 1);
                    }
                    newline(out, indent, depth);
                    out += "}";
                    break;
                }
            }
        }
        static void newline(std::string& out, int indent, int depth) {
            if(indent > 0) { out.push_back('\n'); out.append(static_cast< size_t >(indent) * depth, ' '); }
        }
};

}   /* namespace phq */
#endif
