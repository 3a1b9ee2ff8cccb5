// json_min.hpp -- a small recursive-descent JSON reader for the scene.json schema.
//
// The reference parses scenes with nlohmann::json (Code/json.hpp, third-party). The only
// semantics the scene path depends on are: numbers are read as IEEE doubles (strtod) or
// 64-bit integers and narrowed with static_cast by get<float>() / get<int>(); objects keep
// key lookup; arrays keep order. This reader provides exactly that and nothing more.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace jsonmin {

struct Value;
using Member = std::pair<std::string, Value>;

struct Value {
    enum Kind : uint8_t { Null, Bool, Int, Float, String, Array, Object } kind = Null;
    bool b = false;
    int64_t i = 0;
    double d = 0.0;
    std::string s;
    std::vector<Value> arr;
    std::vector<Member> obj;

    bool is_object() const { return kind == Object; }
    bool is_array() const { return kind == Array; }
    bool is_number() const { return kind == Int || kind == Float; }
    bool is_string() const { return kind == String; }

    const Value* find(const char* key) const {
        if (kind != Object) return nullptr;
        for (const Member& m : obj)
            if (m.first == key) return &m.second;
        return nullptr;
    }
    bool contains(const char* key) const { return find(key) != nullptr; }
    const Value& at(const char* key) const {
        const Value* v = find(key);
        if (!v) throw std::runtime_error(std::string("json: missing key '") + key + "'");
        return *v;
    }
    // nlohmann get<float>(): static_cast<float> of the stored double or integer.
    float as_float() const {
        if (kind == Float) return static_cast<float>(d);
        if (kind == Int) return static_cast<float>(i);
        if (kind == Bool) return b ? 1.0f : 0.0f;
        throw std::runtime_error("json: value is not a number");
    }
    // nlohmann get<int>(): static_cast<int> (truncation toward zero for doubles).
    int as_int() const {
        if (kind == Int) return static_cast<int>(i);
        if (kind == Float) return static_cast<int>(d);
        if (kind == Bool) return b ? 1 : 0;
        throw std::runtime_error("json: value is not a number");
    }
    float value_float(const char* key, float dflt) const {
        const Value* v = find(key);
        return v ? v->as_float() : dflt;
    }
    void as_float3(float out[3]) const {
        if (kind != Array || arr.size() != 3) throw std::runtime_error("json: expected array of 3 numbers");
        for (int k = 0; k < 3; ++k) out[k] = arr[k].as_float();
    }
};

class Parser {
public:
    Parser(const char* begin, const char* end) : p_(begin), end_(end) {}

    Value parse_document() {
        Value v = parse_value();
        skip_ws();
        if (p_ != end_) fail("trailing characters after JSON document");
        return v;
    }

private:
    const char* p_;
    const char* end_;

    [[noreturn]] void fail(const char* msg) const { throw std::runtime_error(std::string("json parse error: ") + msg); }

    void skip_ws() {
        while (p_ != end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }

    Value parse_value() {
        skip_ws();
        if (p_ == end_) fail("unexpected end of input");
        switch (*p_) {
            case '{': return parse_object();
            case '[': return parse_array();
            case '"': { Value v; v.kind = Value::String; v.s = parse_string(); return v; }
            case 't': expect_word("true"); { Value v; v.kind = Value::Bool; v.b = true; return v; }
            case 'f': expect_word("false"); { Value v; v.kind = Value::Bool; v.b = false; return v; }
            case 'n': expect_word("null"); return Value();
            default: return parse_number();
        }
    }

    void expect_word(const char* w) {
        size_t n = std::strlen(w);
        if ((size_t)(end_ - p_) < n || std::memcmp(p_, w, n) != 0) fail("invalid literal");
        p_ += n;
    }

    Value parse_number() {
        const char* start = p_;
        bool is_float = false;
        if (p_ != end_ && *p_ == '-') ++p_;
        if (p_ == end_ || *p_ < '0' || *p_ > '9') fail("invalid number");
        while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        if (p_ != end_ && *p_ == '.') { is_float = true; ++p_; while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
        if (p_ != end_ && (*p_ == 'e' || *p_ == 'E')) {
            is_float = true; ++p_;
            if (p_ != end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        char buf[64];
        size_t n = (size_t)(p_ - start);
        std::string big;
        const char* text;
        if (n < sizeof(buf)) { std::memcpy(buf, start, n); buf[n] = 0; text = buf; }
        else { big.assign(start, n); text = big.c_str(); }
        Value v;
        if (!is_float && n < 19) { v.kind = Value::Int; v.i = std::strtoll(text, nullptr, 10); }
        else { v.kind = Value::Float; v.d = std::strtod(text, nullptr); }
        return v;
    }

    std::string parse_string() {
        ++p_;  // opening quote
        std::string out;
        while (true) {
            if (p_ == end_) fail("unterminated string");
            char c = *p_++;
            if (c == '"') break;
            if (c == '\\') {
                if (p_ == end_) fail("bad escape");
                char e = *p_++;
                switch (e) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        if (end_ - p_ < 4) fail("bad \\u escape");
                        unsigned cp = 0;
                        for (int k = 0; k < 4; ++k) {
                            char h = *p_++;
                            cp <<= 4;
                            if (h >= '0' && h <= '9') cp |= (unsigned)(h - '0');
                            else if (h >= 'a' && h <= 'f') cp |= (unsigned)(h - 'a' + 10);
                            else if (h >= 'A' && h <= 'F') cp |= (unsigned)(h - 'A' + 10);
                            else fail("bad \\u escape");
                        }
                        // UTF-8 encode (surrogate pairs are passed through individually; file names
                        // in scene files are ASCII in practice).
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: fail("bad escape");
                }
            } else {
                out += c;
            }
        }
        return out;
    }

    Value parse_array() {
        ++p_;
        Value v;
        v.kind = Value::Array;
        skip_ws();
        if (p_ != end_ && *p_ == ']') { ++p_; return v; }
        while (true) {
            v.arr.push_back(parse_value());
            skip_ws();
            if (p_ == end_) fail("unterminated array");
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; break; }
            fail("expected ',' or ']'");
        }
        return v;
    }

    Value parse_object() {
        ++p_;
        Value v;
        v.kind = Value::Object;
        skip_ws();
        if (p_ != end_ && *p_ == '}') { ++p_; return v; }
        while (true) {
            skip_ws();
            if (p_ == end_ || *p_ != '"') fail("expected string key");
            std::string key = parse_string();
            skip_ws();
            if (p_ == end_ || *p_ != ':') fail("expected ':'");
            ++p_;
            Value child = parse_value();
            // nlohmann keeps the LAST duplicate key; emulate by overwriting.
            bool replaced = false;
            for (Member& m : v.obj)
                if (m.first == key) { m.second = std::move(child); replaced = true; break; }
            if (!replaced) v.obj.emplace_back(std::move(key), std::move(child));
            skip_ws();
            if (p_ == end_) fail("unterminated object");
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == '}') { ++p_; break; }
            fail("expected ',' or '}'");
        }
        return v;
    }
};

inline Value parse(const std::string& text) {
    Parser p(text.data(), text.data() + text.size());
    return p.parse_document();
}

}  // namespace jsonmin
