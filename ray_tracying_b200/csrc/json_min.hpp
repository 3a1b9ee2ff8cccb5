// json_min.hpp -- a small recursive-descent JSON reader for the scene.json schema.
//
// The reference parses scenes with nlohmann::json (Code/json.hpp, third-party). The only
// semantics the scene path depends on are: numbers are read as IEEE doubles (strtod) or
// 64-bit integers and narrowed with static_cast by get<float>() / get<int>(); objects keep
// key lookup; arrays keep order. This reader provides exactly that and nothing more.
#pragma once

#include <cstdint>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace jsonmin {

struct Value;
using Member = std::pair<std::string, Value>;

// 24 bytes per value: scene files are millions of numbers, so scalars carry no containers; strings,
// arrays and objects keep theirs in a heap box.
struct Value {
    enum Kind : uint8_t { Null, Bool, Int, Float, String, Array, Object } kind = Null;
    union {
        bool b;
        int64_t i;
        double d;
    };
    struct Box;
    std::unique_ptr<Box> box;  // String / Array / Object only

    Value() : i(0) {}
    Value(Value&&) noexcept = default;
    Value& operator=(Value&&) noexcept = default;
    Value(const Value&) = delete;
    Value& operator=(const Value&) = delete;

    bool is_object() const { return kind == Object; }
    bool is_array() const { return kind == Array; }
    bool is_number() const { return kind == Int || kind == Float; }
    bool is_string() const { return kind == String; }

    inline const std::string& str() const;
    inline const std::vector<Value>& array() const;
    inline const Value* find(const char* key) const;
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
    inline void as_float3(float out[3]) const;
};

struct Value::Box {
    std::string s;
    std::vector<Value> arr;
    std::vector<Member> obj;
};

inline const std::string& Value::str() const {
    static const std::string empty;
    return box ? box->s : empty;
}
inline const std::vector<Value>& Value::array() const {
    static const std::vector<Value> empty;
    return box ? box->arr : empty;
}
inline const Value* Value::find(const char* key) const {
    if (kind != Object || !box) return nullptr;
    for (const Member& m : box->obj)
        if (m.first == key) return &m.second;
    return nullptr;
}
inline void Value::as_float3(float out[3]) const {
    if (kind != Array || array().size() != 3) throw std::runtime_error("json: expected array of 3 numbers");
    for (int k = 0; k < 3; ++k) out[k] = box->arr[k].as_float();
}

// The text ranges of the elements of one array of the document that was NOT turned into Values (see
// Parser::defer): big scene files are mostly four arrays of shapes, whose elements the loader parses
// and converts on several threads, each element with its own small Parser.
struct DeferredArray {
    std::string key;
    std::vector<std::pair<const char*, const char*>> elements;  // [begin, end) of each element's text
};

class Parser {
public:
    Parser(const char* begin, const char* end) : p_(begin), end_(end) {}

    // Members of the ROOT object with one of these keys whose value is an array are not parsed: the value
    // becomes an empty Array and the element ranges are recorded in `out` (last duplicate key wins, like
    // every other member).
    void defer(std::vector<std::string> keys, std::vector<DeferredArray>* out) { defer_keys_ = std::move(keys); deferred_ = out; }

    // The value of this key of the ROOT object is not parsed either: it becomes Null and its text range is
    // reported (the loader memoises material blocks by their text: a scene repeats a handful of them millions
    // of times).
    void skip_key(const char* key, std::pair<const char*, const char*>* range) { skip_key_ = key; skip_range_ = range; }

    // Deferred arrays of more than a few MB are scanned by this many threads (1 = always the serial scan).
    void set_threads(unsigned n) { threads_ = n ? n : 1; }

    Value parse_document() {
        depth_ = 0;
        Value v = parse_value();
        skip_ws();
        if (p_ != end_) fail("trailing characters after JSON document");
        return v;
    }

private:
    std::vector<std::string> defer_keys_;
    std::vector<DeferredArray>* deferred_ = nullptr;
    const char* skip_key_ = nullptr;
    std::pair<const char*, const char*>* skip_range_ = nullptr;
    int depth_ = 0;
    unsigned threads_ = 1;

    static bool is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

    // first byte in [q, e) that is one of " { } [ ] (and , if `comma`), or e
    static const char* next_structural(const char* q, const char* e, bool comma) {
#if defined(__SSE2__)
        while (e - q >= 16) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(q));
            const __m128i f = _mm_or_si128(v, _mm_set1_epi8(0x20));
            __m128i hit = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(f, _mm_set1_epi8('{')), _mm_cmpeq_epi8(f, _mm_set1_epi8('}'))),
                                       _mm_cmpeq_epi8(v, _mm_set1_epi8('"')));
            if (comma) hit = _mm_or_si128(hit, _mm_cmpeq_epi8(v, _mm_set1_epi8(',')));
            const int mask = _mm_movemask_epi8(hit);
            if (mask) return q + __builtin_ctz((unsigned)mask);
            q += 16;
        }
#endif
        for (; q < e; ++q) {
            const char c = *q;
            if (c == '"' || c == '{' || c == '}' || c == '[' || c == ']' || (comma && c == ',')) return q;
        }
        return e;
    }

    // A '"' is a string delimiter unless an odd number of backslashes stands before it (only meaningful inside
    // strings; outside, a backslash is malformed and the element's own parse will say so).
    static bool real_quote(const char* q, const char* region_begin) {
        int n = 0;
        while (q - n - 1 >= region_begin && *(q - n - 1) == '\\') ++n;
        return (n & 1) == 0;
    }

    // The REST of an array -- p_ stands at the first byte of an element, bracket depth 1 -- scanned by several threads:
    //   pass A: string delimiters per chunk            -> is each chunk's first byte inside a string?
    //   pass B: bracket depth change per chunk          -> depth at each chunk's first byte, chunk in which the array closes
    //   pass C: commas at depth 1 and the closing ']'   -> element boundaries
    // Appends the elements to d and leaves p_ behind the ']'. Returns false (nothing consumed, nothing appended) whenever
    // the outcome is not a plain list of non-empty elements: the serial scan then continues and reports malformed input
    // exactly as before.
    bool defer_rest_parallel(DeferredArray* d) {
        const char* const rb = p_;
        const size_t len = (size_t)(end_ - rb);
        const size_t chunks = std::min<size_t>((size_t)threads_ * 4, len >> 20);
        if (threads_ < 2 || chunks < 2) return false;
        auto cb = [&](size_t c) { return rb + len * c / chunks; };
        auto run = [&](auto fn) {
            std::vector<std::thread> pool;
            const unsigned t_n = (unsigned)std::min<size_t>(threads_, chunks);
            for (unsigned t = 1; t < t_n; ++t) pool.emplace_back([=] { for (size_t c = t; c < chunks; c += t_n) fn(c); });
            for (size_t c = 0; c < chunks; c += t_n) fn(c);
            for (std::thread& th : pool) th.join();
        };
        std::vector<size_t> quotes(chunks, 0);
        run([&](size_t c) {
            size_t n = 0;
            const char* e = cb(c + 1);
            for (const char* q = cb(c); q < e; ++q) {
                q = static_cast<const char*>(std::memchr(q, '"', (size_t)(e - q)));
                if (!q) break;
                if (real_quote(q, rb)) ++n;
            }
            quotes[c] = n;
        });
        std::vector<char> in_str(chunks + 1, 0);
        for (size_t c = 0; c < chunks; ++c) in_str[c + 1] = (char)((in_str[c] + quotes[c]) & 1);
        // per chunk: net change of the bracket depth, and the lowest depth reached right after a closing bracket
        // (relative to the chunk's first byte; "none" if the chunk closes nothing)
        const long long NONE = (long long)1 << 60;
        std::vector<long long> delta(chunks, 0), low(chunks, NONE);
        run([&](size_t c) {
            bool s = in_str[c] != 0;
            long long dep = 0, mn = NONE;
            const char* e = cb(c + 1);
            for (const char* q = next_structural(cb(c), e, false); q < e; q = next_structural(q + 1, e, false)) {
                const char ch = *q;
                if (ch == '"') { if (real_quote(q, rb)) s = !s; }
                else if (!s) {
                    if (ch == '{' || ch == '[') ++dep;
                    else { --dep; mn = std::min(mn, dep); }
                }
            }
            delta[c] = dep;
            low[c] = mn;
        });
        // the region starts at depth 1: the array closes in the first chunk in which a closing bracket brings the depth
        // to 0 (pass C finds the byte)
        std::vector<long long> dep0(chunks + 1, 1);
        size_t last = chunks;
        for (size_t c = 0; c < chunks; ++c) {
            if (last == chunks && low[c] != NONE && dep0[c] + low[c] <= 0) last = c;
            dep0[c + 1] = dep0[c] + delta[c];
        }
        if (last == chunks) return false;  // unterminated: let the serial scan report it
        std::vector<std::vector<const char*>> commas(last + 1);
        std::vector<const char*> close(last + 1, nullptr);
        run([&](size_t c) {
            if (c > last) return;
            bool s = in_str[c] != 0;
            long long dep = dep0[c];
            const char* e = cb(c + 1);
            for (const char* q = next_structural(cb(c), e, true); q < e; q = next_structural(q + 1, e, true)) {
                const char ch = *q;
                if (ch == '"') { if (real_quote(q, rb)) s = !s; }
                else if (!s) {
                    if (ch == '{' || ch == '[') ++dep;
                    else if (ch == ',') { if (dep == 1) commas[c].push_back(q); }
                    else if (--dep == 0) { close[c] = q; return; }
                }
            }
        });
        const char* end_pos = nullptr;
        size_t end_chunk = 0;
        for (size_t c = 0; c <= last; ++c) if (close[c]) { end_pos = close[c]; end_chunk = c; break; }
        if (!end_pos || *end_pos != ']') return false;
        std::vector<std::pair<const char*, const char*>> elems;
        const char* b = rb;
        auto add = [&](const char* e) -> bool {
            const char* x = b;
            const char* y = e;
            while (x < y && is_ws(*x)) ++x;
            while (y > x && is_ws(*(y - 1))) --y;
            if (x == y) return false;
            elems.emplace_back(x, y);
            return true;
        };
        for (size_t c = 0; c <= end_chunk; ++c)
            for (const char* q : commas[c]) {
                if (q > end_pos) break;
                if (!add(q)) return false;
                b = q + 1;
            }
        if (!add(end_pos)) return false;  // "[x, ]" is for the serial scan to reject
        d->elements.insert(d->elements.end(), elems.begin(), elems.end());
        p_ = end_pos + 1;
        return true;
    }

    // Skips one value without building it; the text must be well formed as far as brackets and strings go
    // (anything else is caught when the element itself is parsed).
    void skip_value() {
        skip_ws();
        if (p_ == end_) fail("unexpected end of input");
        if (*p_ == '"') { skip_string(); return; }
        if (*p_ != '{' && *p_ != '[') {
            while (p_ != end_ && *p_ != ',' && *p_ != ']' && *p_ != '}' && *p_ != ' ' && *p_ != '\n' && *p_ != '\t' && *p_ != '\r') ++p_;
            return;
        }
        int depth = 0;
        while (p_ != end_) {
#if defined(__SSE2__)
            // 16 bytes at a time until one of " { } [ ] shows up (scene files are mostly digits and commas)
            while (end_ - p_ >= 16) {
                const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p_));
                // { } differ from [ ] by bit 5 (0x20): fold them together, then two compares cover all four
                const __m128i f = _mm_or_si128(v, _mm_set1_epi8(0x20));
                const __m128i hit = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(f, _mm_set1_epi8('{')), _mm_cmpeq_epi8(f, _mm_set1_epi8('}'))),
                                                 _mm_cmpeq_epi8(v, _mm_set1_epi8('"')));
                const int mask = _mm_movemask_epi8(hit);
                if (mask) { p_ += __builtin_ctz((unsigned)mask); break; }
                p_ += 16;
            }
            if (p_ == end_) break;
#endif
            const char c = *p_;
            if (c == '"') { skip_string(); continue; }
            ++p_;
            if (c == '{' || c == '[') ++depth;
            else if (c == '}' || c == ']') { if (--depth == 0) return; }
        }
        fail("unterminated array or object");
    }
    void skip_string() {
        ++p_;
        while (p_ != end_) {
            const char c = *p_++;
            if (c == '"') return;
            if (c == '\\') { if (p_ == end_) break; ++p_; }
        }
        fail("unterminated string");
    }
    void defer_array(const std::string& key) {
        DeferredArray* d = nullptr;
        for (DeferredArray& e : *deferred_) if (e.key == key) d = &e;
        if (!d) { deferred_->emplace_back(); d = &deferred_->back(); d->key = key; }
        d->elements.clear();
        const char* const begin = p_;
        ++p_;  // '['
        skip_ws();
        if (p_ != end_ && *p_ == ']') { ++p_; return; }
        bool tried = false;
        while (true) {
            skip_ws();
            // an array that is still open after 1 MB is a big one: the rest is scanned by all threads
            if (!tried && p_ - begin > (1 << 20) && (size_t)(end_ - p_) > ((size_t)4 << 20)) {
                tried = true;
                if (defer_rest_parallel(d)) return;
            }
            const char* b = p_;
            skip_value();
            d->elements.emplace_back(b, p_);
            skip_ws();
            if (p_ == end_) fail("unterminated array");
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; break; }
            fail("expected ',' or ']'");
        }
    }

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
            case '"': { Value v; v.kind = Value::String; v.box.reset(new Value::Box()); v.box->s = parse_string(); return v; }
            case 't': expect_word("true"); { Value v; v.kind = Value::Bool; v.b = true; return v; }
            case 'f': expect_word("false"); { Value v; v.kind = Value::Bool; v.b = false; return v; }
            case 'n': expect_word("null"); { Value v; return v; }
            default: return parse_number();
        }
    }

    void expect_word(const char* w) {
        size_t n = std::strlen(w);
        if ((size_t)(end_ - p_) < n || std::memcmp(p_, w, n) != 0) fail("invalid literal");
        p_ += n;
    }

    // Numbers as nlohmann reads them: integers without fraction / exponent as int64, everything else as
    // the correctly rounded IEEE double (strtod). Fast path (Clinger): a decimal significand below 2^53
    // times / over an exactly representable power of ten (<= 1e22) is ONE correctly rounded IEEE
    // operation on exact operands, i.e. the same double strtod returns; anything else goes to strtod.
    Value parse_number() {
        static const double pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                         1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        const char* start = p_;
        bool is_float = false, negative = false;
        if (p_ != end_ && *p_ == '-') { negative = true; ++p_; }
        if (p_ == end_ || *p_ < '0' || *p_ > '9') fail("invalid number");
        uint64_t mant = 0;
        int digits = 0, exp10 = 0;  // significant digits accumulated (leading zeros do not count)
        while (p_ != end_ && *p_ >= '0' && *p_ <= '9') {
            if (digits < 19) { mant = mant * 10u + (uint64_t)(*p_ - '0'); if (mant != 0) ++digits; }
            else { ++exp10; ++digits; }
            ++p_;
        }
        if (p_ != end_ && *p_ == '.') {
            is_float = true;
            ++p_;
            while (p_ != end_ && *p_ >= '0' && *p_ <= '9') {
                if (digits < 19) { mant = mant * 10u + (uint64_t)(*p_ - '0'); if (mant != 0) ++digits; --exp10; }
                else ++digits;
                ++p_;
            }
        }
        bool has_exp = false;
        if (p_ != end_ && (*p_ == 'e' || *p_ == 'E')) {
            is_float = true; has_exp = true; ++p_;
            bool eneg = false;
            if (p_ != end_ && (*p_ == '+' || *p_ == '-')) { eneg = *p_ == '-'; ++p_; }
            int e = 0;
            while (p_ != end_ && *p_ >= '0' && *p_ <= '9') { if (e < 100000) e = e * 10 + (*p_ - '0'); ++p_; }
            exp10 += eneg ? -e : e;
        }
        (void)has_exp;
        const size_t n = (size_t)(p_ - start);
        Value v;
        if (!is_float && n < 19) {
            v.kind = Value::Int;
            v.i = negative ? -(int64_t)mant : (int64_t)mant;
            return v;
        }
        v.kind = Value::Float;
        if (digits <= 19 && mant < (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
            double x = (double)mant;
            x = exp10 < 0 ? x / pow10[-exp10] : x * pow10[exp10];
            v.d = negative ? -x : x;
            return v;
        }
        char buf[64];
        std::string big;
        const char* text;
        if (n < sizeof(buf)) { std::memcpy(buf, start, n); buf[n] = 0; text = buf; }
        else { big.assign(start, n); text = big.c_str(); }
        v.d = std::strtod(text, nullptr);
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
        ++depth_;
        Value v;
        v.kind = Value::Array;
        v.box.reset(new Value::Box());
        v.box->arr.reserve(4);  // most arrays of a scene file are 3-vectors
        skip_ws();
        if (p_ != end_ && *p_ == ']') { ++p_; --depth_; return v; }
        while (true) {
            v.box->arr.push_back(parse_value());
            skip_ws();
            if (p_ == end_) fail("unterminated array");
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; break; }
            fail("expected ',' or ']'");
        }
        --depth_;
        return v;
    }

    Value parse_object() {
        ++p_;
        const bool root = depth_ == 0;
        ++depth_;
        Value v;
        v.kind = Value::Object;
        v.box.reset(new Value::Box());
        v.box->obj.reserve(8);
        skip_ws();
        if (p_ != end_ && *p_ == '}') { ++p_; --depth_; return v; }
        while (true) {
            skip_ws();
            if (p_ == end_ || *p_ != '"') fail("expected string key");
            std::string key = parse_string();
            skip_ws();
            if (p_ == end_ || *p_ != ':') fail("expected ':'");
            ++p_;
            Value child;
            bool deferred = false;
            if (root && deferred_) {
                skip_ws();
                if (p_ != end_ && *p_ == '[')
                    for (const std::string& k : defer_keys_) if (k == key) deferred = true;
            }
            if (root && skip_key_ && key == skip_key_) {
                skip_ws();
                const char* b = p_;
                skip_value();
                *skip_range_ = std::make_pair(b, p_);  // the last duplicate wins, like every other member
                child.kind = Value::Null;
            } else if (deferred) {
                defer_array(key);
                child.kind = Value::Array;
                child.box.reset(new Value::Box());
            } else {
                if (root && deferred_)  // a later duplicate that is not an array replaces an earlier deferred one
                    for (size_t i = 0; i < deferred_->size(); ++i) if ((*deferred_)[i].key == key) { deferred_->erase(deferred_->begin() + (long)i); break; }
                child = parse_value();
            }
            // nlohmann keeps the LAST duplicate key; emulate by overwriting.
            bool replaced = false;
            for (Member& m : v.box->obj)
                if (m.first == key) { m.second = std::move(child); replaced = true; break; }
            if (!replaced) v.box->obj.emplace_back(std::move(key), std::move(child));
            skip_ws();
            if (p_ == end_) fail("unterminated object");
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == '}') { ++p_; break; }
            fail("expected ',' or '}'");
        }
        --depth_;
        return v;
    }
};

inline Value parse(const std::string& text) {
    Parser p(text.data(), text.data() + text.size());
    return p.parse_document();
}

// One element of a deferred array.
inline Value parse_range(const char* begin, const char* end) {
    Parser p(begin, end);
    return p.parse_document();
}
// ... with the value of root key `key` left unparsed (Null) and its text range in `range` ({nullptr, nullptr} if absent).
inline Value parse_range_skipping(const char* begin, const char* end, const char* key, std::pair<const char*, const char*>& range) {
    range = std::make_pair((const char*)nullptr, (const char*)nullptr);
    Parser p(begin, end);
    p.skip_key(key, &range);
    return p.parse_document();
}

}  // namespace jsonmin
