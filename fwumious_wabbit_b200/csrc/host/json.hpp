// Minimal JSON DOM: enough to read and write the two blobs the reference embeds in its files
// (VwNamespaceMapSource and ModelInstance, serde_json pretty-printed; persistence.rs:20-52).
// Object key order is preserved (the reference writes fields in declaration order; readers accept any).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fwhost {

struct JValue {
    enum Type { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    double num = 0.0;
    bool is_int = false;   // print without a fraction
    bool is_f32 = false;   // print as the shortest string that round-trips an f32 (serde_json / ryu behaviour)
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;

    static JValue null() { return JValue(); }
    static JValue boolean(bool v) { JValue j; j.t = Bool; j.b = v; return j; }
    static JValue integer(int64_t v) { JValue j; j.t = Num; j.num = (double)v; j.is_int = true; return j; }
    static JValue f32(float v) { JValue j; j.t = Num; j.num = v; j.is_f32 = true; return j; }
    static JValue string(const std::string &s) { JValue j; j.t = Str; j.str = s; return j; }
    static JValue array() { JValue j; j.t = Arr; return j; }
    static JValue object() { JValue j; j.t = Obj; return j; }

    void set(const std::string &k, JValue v) { obj.emplace_back(k, std::move(v)); }
    const JValue *get(const std::string &k) const
    {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const JValue &at(const std::string &k) const
    {
        const JValue *v = get(k);
        if (!v) throw std::runtime_error("missing field `" + k + "`");
        return *v;
    }
    double as_num() const { if (t != Num) throw std::runtime_error("expected a number"); return num; }
    bool as_bool() const { if (t != Bool) throw std::runtime_error("expected a boolean"); return b; }
    const std::string &as_str() const { if (t != Str) throw std::runtime_error("expected a string"); return str; }
};

class JParser {
    const char *p, *e;
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("JSON: ") + m); }
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    std::string parse_string()
    {
        if (p >= e || *p != '"') fail("expected string");
        p++;
        std::string s;
        while (p < e && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= e) fail("bad escape");
                switch (*p) {
                case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                case 'u': {
                    if (e - p < 5) fail("bad \\u");
                    unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                    p += 4;
                    if (cp < 0x80) s += (char)cp;
                    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
                    else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: s += *p;
                }
                p++;
            } else s += *p++;
        }
        if (p >= e) fail("unterminated string");
        p++;
        return s;
    }
    JValue value()
    {
        ws();
        if (p >= e) fail("unexpected end");
        JValue v;
        if (*p == '{') {
            v.t = JValue::Obj; p++; ws();
            if (p < e && *p == '}') { p++; return v; }
            for (;;) {
                ws();
                std::string k = parse_string();
                ws();
                if (p >= e || *p != ':') fail("expected ':'");
                p++;
                v.obj.emplace_back(k, value());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (*p == '[') {
            v.t = JValue::Arr; p++; ws();
            if (p < e && *p == ']') { p++; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (*p == '"') {
            v.t = JValue::Str; v.str = parse_string();
        } else if (!strncmp(p, "true", 4) && e - p >= 4) { v = JValue::boolean(true); p += 4; }
        else if (!strncmp(p, "false", 5) && e - p >= 5) { v = JValue::boolean(false); p += 5; }
        else if (!strncmp(p, "null", 4) && e - p >= 4) { p += 4; }
        else {
            char *end = nullptr;
            std::string tmp(p, std::min<size_t>(e - p, 64));
            double d = strtod(tmp.c_str(), &end);
            if (end == tmp.c_str()) fail("unexpected token");
            v.t = JValue::Num; v.num = d;
            std::string tok(tmp.c_str(), (size_t)(end - tmp.c_str()));
            v.is_int = tok.find_first_of(".eE") == std::string::npos;
            p += end - tmp.c_str();
        }
        return v;
    }
public:
    JParser(const char *s, size_t n) : p(s), e(s + n) {}
    JValue parse() { JValue v = value(); ws(); return v; }
};

inline JValue json_parse(const std::string &s) { return JParser(s.data(), s.size()).parse(); }

inline std::string f32_shortest(float f)
{
    if (std::isnan(f)) return "null"; // serde_json writes non-finite floats as null
    if (std::isinf(f)) return "null";
    char buf[64];
    for (int prec = 1; prec <= 9; prec++) {
        snprintf(buf, sizeof(buf), "%.*g", prec, (double)f);
        if (strtof(buf, nullptr) == f) break;
    }
    std::string s(buf);
    // ryu / serde style: plain decimals for moderate exponents, always with a fraction
    if (s.find('e') != std::string::npos) {
        // expand small exponents like 1e-05 the way ryu does for |exp| < 5..16 : keep it simple, use %f when exact
        double d = f;
        if (std::fabs(d) >= 1e-5 && std::fabs(d) < 1e16) {
            for (int prec = 1; prec <= 20; prec++) {
                snprintf(buf, sizeof(buf), "%.*f", prec, d);
                if (strtof(buf, nullptr) == f) { s = buf; break; }
            }
        } else {
            // normalise exponent form: 1e-07 -> 1e-7
            size_t pos = s.find('e');
            std::string mant = s.substr(0, pos), ex = s.substr(pos + 1);
            int exv = atoi(ex.c_str());
            if (mant.find('.') == std::string::npos) mant += ".0";
            s = mant + "e" + std::to_string(exv);
            return s;
        }
    }
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos) s += ".0";
    return s;
}

inline void json_escape(const std::string &s, std::string &out)
{
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
        case '"': out += "\\\""; break; case '\\': out += "\\\\"; break; case '\n': out += "\\n"; break;
        case '\t': out += "\\t"; break; case '\r': out += "\\r"; break;
        default:
            if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; }
            else out += (char)c;
        }
    }
    out += '"';
}

// serde_json::to_vec_pretty layout: two-space indentation, "key": value, [] and {} for empties
inline void json_dump(const JValue &v, std::string &out, int indent = 0)
{
    auto pad = [&](int n) { out.append((size_t)n * 2, ' '); };
    switch (v.t) {
    case JValue::Null: out += "null"; break;
    case JValue::Bool: out += v.b ? "true" : "false"; break;
    case JValue::Num:
        if (v.is_int) { char b[32]; snprintf(b, sizeof(b), "%lld", (long long)v.num); out += b; }
        else if (v.is_f32) out += f32_shortest((float)v.num);
        else { char b[40]; snprintf(b, sizeof(b), "%.17g", v.num); std::string s(b); if (s.find_first_of(".e") == std::string::npos) s += ".0"; out += s; }
        break;
    case JValue::Str: json_escape(v.str, out); break;
    case JValue::Arr:
        if (v.arr.empty()) { out += "[]"; break; }
        out += "[\n";
        for (size_t i = 0; i < v.arr.size(); i++) { pad(indent + 1); json_dump(v.arr[i], out, indent + 1); out += (i + 1 < v.arr.size()) ? ",\n" : "\n"; }
        pad(indent); out += "]";
        break;
    case JValue::Obj:
        if (v.obj.empty()) { out += "{}"; break; }
        out += "{\n";
        for (size_t i = 0; i < v.obj.size(); i++) {
            pad(indent + 1); json_escape(v.obj[i].first, out); out += ": "; json_dump(v.obj[i].second, out, indent + 1);
            out += (i + 1 < v.obj.size()) ? ",\n" : "\n";
        }
        pad(indent); out += "}";
        break;
    }
}

inline std::string json_to_string(const JValue &v) { std::string s; json_dump(v, s, 0); return s; }

} // namespace fwhost
