// value.h — instruction payload type and JSON reader for the host side.
//
// Plays the role of the reference's elem::js::Value + parseJSON (runtime/elem/Value.h:37-197, JSON.h:17-156)
// for the one thing this library needs them for: decoding instruction batches
//   [[0,id,"type"],[2,parent,child,chan],[3,id,"key",value],[4,[roots]],[5]]       
// Numbers are doubles, ids are cast to int32 exactly like Runtime.h:299,321,341-343.  Written from scratch
// (a ~150-line recursive-descent parser) instead of vendoring a JSON library.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace eb {

class Value {
public:
    enum class Type { Undefined, Null, Bool, Number, String, Array, Object };

    Value() = default;
    static Value null() { Value v; v.type_ = Type::Null; return v; }
    static Value boolean(bool b) { Value v; v.type_ = Type::Bool; v.b_ = b; return v; }
    static Value number(double d) { Value v; v.type_ = Type::Number; v.n_ = d; return v; }
    static Value string(std::string s) { Value v; v.type_ = Type::String; v.s_ = std::move(s); return v; }
    static Value array() { Value v; v.type_ = Type::Array; return v; }
    static Value object() { Value v; v.type_ = Type::Object; return v; }

    Type type() const { return type_; }
    bool isUndefined() const { return type_ == Type::Undefined; }
    bool isNull() const { return type_ == Type::Null; }
    bool isBool() const { return type_ == Type::Bool; }
    bool isNumber() const { return type_ == Type::Number; }
    bool isString() const { return type_ == Type::String; }
    bool isArray() const { return type_ == Type::Array; }
    bool isObject() const { return type_ == Type::Object; }

    bool asBool() const { return b_; }
    double asNumber() const { return n_; }
    const std::string& asString() const { return s_; }
    const std::vector<Value>& asArray() const { return a_; }
    std::vector<Value>& asArray() { return a_; }
    const std::map<std::string, Value>& asObject() const { return o_; }
    std::map<std::string, Value>& asObject() { return o_; }

    bool operator==(const Value& o) const {
        if (type_ != o.type_) return false;
        switch (type_) {
            case Type::Bool: return b_ == o.b_;
            case Type::Number: return n_ == o.n_;
            case Type::String: return s_ == o.s_;
            case Type::Array: return a_ == o.a_;
            case Type::Object: return o_ == o.o_;
            default: return true;
        }
    }

private:
    Type type_ = Type::Undefined;
    bool b_ = false;
    double n_ = 0.0;
    std::string s_;
    std::vector<Value> a_;
    std::map<std::string, Value> o_;
};

struct JsonError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class JsonParser {
public:
    JsonParser(const char* s, size_t n) : p_(s), end_(s + n) {}

    Value parseDocument() {
        Value v = parseValue(0);
        skipWs();
        if (p_ != end_) fail("trailing characters");
        return v;
    }

private:
    const char* p_;
    const char* end_;

    [[noreturn]] void fail(const char* what) const { throw JsonError(std::string("JSON: ") + what); }

    void skipWs() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
    }

    bool consume(const char* lit) {
        size_t n = 0;
        while (lit[n]) ++n;
        if ((size_t) (end_ - p_) < n) return false;
        for (size_t i = 0; i < n; ++i) if (p_[i] != lit[i]) return false;
        p_ += n;
        return true;
    }

    Value parseValue(int depth) {
        if (depth > 256) fail("nesting too deep");
        skipWs();
        if (p_ >= end_) fail("unexpected end");
        const char ch = *p_;
        if (ch == '[') {
            ++p_;
            Value v = Value::array();
            skipWs();
            if (p_ < end_ && *p_ == ']') { ++p_; return v; }
            for (;;) {
                v.asArray().push_back(parseValue(depth + 1));
                skipWs();
                if (p_ >= end_) fail("unterminated array");
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == ']') { ++p_; break; }
                fail("expected , or ]");
            }
            return v;
        }
        if (ch == '{') {
            ++p_;
            Value v = Value::object();
            skipWs();
            if (p_ < end_ && *p_ == '}') { ++p_; return v; }
            for (;;) {
                skipWs();
                if (p_ >= end_ || *p_ != '"') fail("expected object key");
                std::string key = parseString();
                skipWs();
                if (p_ >= end_ || *p_ != ':') fail("expected :");
                ++p_;
                v.asObject()[key] = parseValue(depth + 1);
                skipWs();
                if (p_ >= end_) fail("unterminated object");
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; break; }
                fail("expected , or }");
            }
            return v;
        }
        if (ch == '"') return Value::string(parseString());
        if (consume("true")) return Value::boolean(true);
        if (consume("false")) return Value::boolean(false);
        if (consume("null")) return Value::null();
        if (ch == '-' || (ch >= '0' && ch <= '9')) return Value::number(parseNumber());
        // JSON.stringify never emits these, but Python's json module does; accept them as numbers.
        if (consume("NaN")) return Value::number(std::nan(""));
        if (consume("Infinity")) return Value::number(HUGE_VAL);
        fail("unexpected character");
    }

    double parseNumber() {
        const char* start = p_;
        if (p_ < end_ && *p_ == '-') {
            ++p_;
            if (consume("Infinity")) return -HUGE_VAL;
        }
        while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) ++p_;
        std::string tmp(start, p_);
        char* endp = nullptr;
        const double d = std::strtod(tmp.c_str(), &endp);
        if (endp == tmp.c_str() || *endp != '\0') fail("bad number");
        return d;
    }

    static void appendUtf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) out.push_back((char) cp);
        else if (cp < 0x800) { out.push_back((char) (0xC0 | (cp >> 6))); out.push_back((char) (0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) {
            out.push_back((char) (0xE0 | (cp >> 12))); out.push_back((char) (0x80 | ((cp >> 6) & 0x3F))); out.push_back((char) (0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char) (0xF0 | (cp >> 18))); out.push_back((char) (0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char) (0x80 | ((cp >> 6) & 0x3F))); out.push_back((char) (0x80 | (cp & 0x3F)));
        }
    }

    uint32_t parseHex4() {
        if (end_ - p_ < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            const char ch = *p_++;
            v <<= 4;
            if (ch >= '0' && ch <= '9') v |= (uint32_t) (ch - '0');
            else if (ch >= 'a' && ch <= 'f') v |= (uint32_t) (ch - 'a' + 10);
            else if (ch >= 'A' && ch <= 'F') v |= (uint32_t) (ch - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }

    std::string parseString() {
        ++p_;   // opening quote
        std::string out;
        for (;;) {
            if (p_ >= end_) fail("unterminated string");
            const char ch = *p_++;
            if (ch == '"') break;
            if (ch != '\\') { out.push_back(ch); continue; }
            if (p_ >= end_) fail("bad escape");
            const char e = *p_++;
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp = parseHex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        const uint32_t lo = parseHex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    appendUtf8(out, cp);
                } break;
                default: fail("bad escape");
            }
        }
        return out;
    }
};

inline Value parseJson(const char* s, size_t n) { return JsonParser(s, n).parseDocument(); }


// JSON text of a Value (what the reference's js::serialize does for the types an instruction batch can carry, JSON.h:160-248):
// numbers in shortest round-trip form, undefined as null.
inline void writeJson(std::string& o, const Value& v) {
    switch (v.type()) {
        case Value::Type::Undefined: case Value::Type::Null: o += "null"; break;
        case Value::Type::Bool: o += v.asBool() ? "true" : "false"; break;
        case Value::Type::Number: {
            const double d = v.asNumber();
            if (!std::isfinite(d)) { o += "null"; break; }
            char b[40];
            if (d == std::floor(d) && std::fabs(d) < 1e15) std::snprintf(b, sizeof b, "%.0f", d);
            else std::snprintf(b, sizeof b, "%.17g", d);
            o += b;
        } break;
        case Value::Type::String: {
            o += '"';
            for (unsigned char ch : v.asString()) {
                if (ch == '"' || ch == '\\') { o += '\\'; o += (char) ch; }
                else if (ch == '\n') o += "\\n";
                else if (ch == '\t') o += "\\t";
                else if (ch < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
                else o += (char) ch;
            }
            o += '"';
        } break;
        case Value::Type::Array: {
            o += '[';
            bool first = true;
            for (auto& e : v.asArray()) { if (!first) o += ','; first = false; writeJson(o, e); }
            o += ']';
        } break;
        case Value::Type::Object: {
            o += '{';
            bool first = true;
            for (auto& kv : v.asObject()) { if (!first) o += ','; first = false; writeJson(o, Value::string(kv.first)); o += ':'; writeJson(o, kv.second); }
            o += '}';
        } break;
    }
}

} // namespace eb
