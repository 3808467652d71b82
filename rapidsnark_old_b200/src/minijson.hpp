// Minimal JSON value: parse + compact dump with object keys in sorted order, i.e. what the reference gets from
// `json j = json::parse(text); file << j;` with nlohmann::json (src/fullprover.cpp:108-113; nlohmann's default object
// type is a std::map, its operator<< dumps without whitespace).  Only what FullProver needs: circom inputs are
// objects / arrays of strings and integers.  Numbers keep their source spelling (an integer token is printed back
// unchanged, which is also what nlohmann does for values that fit 64 bits).  Malformed text throws
// std::runtime_error like nlohmann's parse_error (which derives from it through json::exception -> std::exception;
// the reference's worker only catches std::runtime_error, fullprover.cpp:163 - here the request fails cleanly).
#ifndef B200_MINIJSON_HPP
#define B200_MINIJSON_HPP
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace minijson {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    std::string text;                                  // Number: source token; String: decoded bytes (UTF-8)
    std::vector<Value> items;
    std::map<std::string, Value> members;
};

class Parser {
    const std::string &s;
    size_t i = 0;
    int depth = 0;

    [[noreturn]] void fail(const char *what) const {
        throw std::runtime_error("[json.exception.parse_error.101] parse error at byte " + std::to_string(i + 1) + ": " + what);
    }
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
    bool lit(const char *w) {
        size_t n = 0;
        while (w[n]) n++;
        if (s.compare(i, n, w) != 0) return false;
        i += n;
        return true;
    }
    static void utf8(std::string &o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (i + 4 > s.size()) fail("truncated \\u escape");
        unsigned v = 0;
        for (int k = 0; k < 4; k++) {
            char c = s[i++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string str() {
        std::string o;
        i++;   // opening quote
        for (;;) {
            if (i >= s.size()) fail("unterminated string");
            unsigned char c = (unsigned char)s[i++];
            if (c == '"') return o;
            if (c < 0x20) fail("control character in string");
            if (c != '\\') { o += (char)c; continue; }
            if (i >= s.size()) fail("unterminated escape");
            char e = s[i++];
            switch (e) {
                case '"': o += '"'; break;
                case '\\': o += '\\'; break;
                case '/': o += '/'; break;
                case 'b': o += '\b'; break;
                case 'f': o += '\f'; break;
                case 'n': o += '\n'; break;
                case 'r': o += '\r'; break;
                case 't': o += '\t'; break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00) {
                        if (!(i + 1 < s.size() && s[i] == '\\' && s[i + 1] == 'u')) fail("lone surrogate");
                        i += 2;
                        unsigned lo = hex4();
                        if (lo < 0xDC00 || lo > 0xDFFF) fail("bad surrogate pair");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    } else if (cp >= 0xDC00 && cp < 0xE000) fail("lone surrogate");
                    utf8(o, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
    std::string number() {
        size_t b = i;
        if (s[i] == '-') i++;
        if (i >= s.size()) fail("bad number");
        if (s[i] == '0') i++;
        else if (s[i] >= '1' && s[i] <= '9') { while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++; }
        else fail("bad number");
        if (i < s.size() && s[i] == '.') {
            i++;
            if (i >= s.size() || s[i] < '0' || s[i] > '9') fail("bad fraction");
            while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++;
        }
        if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
            i++;
            if (i < s.size() && (s[i] == '+' || s[i] == '-')) i++;
            if (i >= s.size() || s[i] < '0' || s[i] > '9') fail("bad exponent");
            while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++;
        }
        return s.substr(b, i - b);
    }
    Value value() {
        if (++depth > 512) fail("nesting too deep");
        ws();
        if (i >= s.size()) fail("unexpected end of input");
        Value v;
        char c = s[i];
        if (c == '{') {
            v.kind = Value::Object;
            i++;
            ws();
            if (i < s.size() && s[i] == '}') { i++; depth--; return v; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("object key expected");
                std::string k = str();
                ws();
                if (i >= s.size() || s[i] != ':') fail("':' expected");
                i++;
                v.members[k] = value();          // a repeated key keeps the last value, as nlohmann does
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == '}') { i++; break; }
                fail("',' or '}' expected");
            }
        } else if (c == '[') {
            v.kind = Value::Array;
            i++;
            ws();
            if (i < s.size() && s[i] == ']') { i++; depth--; return v; }
            for (;;) {
                v.items.push_back(value());
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == ']') { i++; break; }
                fail("',' or ']' expected");
            }
        } else if (c == '"') {
            v.kind = Value::String;
            v.text = str();
        } else if (c == '-' || (c >= '0' && c <= '9')) {
            v.kind = Value::Number;
            v.text = number();
        } else if (lit("true")) { v.kind = Value::Bool; v.b = true; }
        else if (lit("false")) { v.kind = Value::Bool; v.b = false; }
        else if (lit("null")) { v.kind = Value::Null; }
        else fail("unexpected character");
        depth--;
        return v;
    }

public:
    explicit Parser(const std::string &text) : s(text) {}
    Value parse() {
        Value v = value();
        ws();
        if (i != s.size()) fail("trailing characters");
        return v;
    }
};

inline Value parse(const std::string &text) { return Parser(text).parse(); }

inline void escape(std::string &o, const std::string &s) {
    for (unsigned char ch : s) {
        switch (ch) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default:
                if (ch < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
                else o += (char)ch;
        }
    }
}

inline void dump(std::string &o, const Value &v) {
    switch (v.kind) {
        case Value::Null: o += "null"; break;
        case Value::Bool: o += v.b ? "true" : "false"; break;
        case Value::Number: o += v.text; break;
        case Value::String: o += '"'; escape(o, v.text); o += '"'; break;
        case Value::Array: {
            o += '[';
            for (size_t k = 0; k < v.items.size(); k++) { if (k) o += ','; dump(o, v.items[k]); }
            o += ']';
            break;
        }
        case Value::Object: {
            o += '{';
            bool first = true;
            for (auto &kv : v.members) {
                if (!first) o += ',';
                first = false;
                o += '"'; escape(o, kv.first); o += "\":";
                dump(o, kv.second);
            }
            o += '}';
            break;
        }
    }
}

inline std::string dump(const Value &v) { std::string o; dump(o, v); return o; }

}  // namespace minijson
#endif
