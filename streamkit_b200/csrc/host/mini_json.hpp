// mini_json.hpp -- the small subset of JSON the node parameter blobs use (objects, numbers, bools, null,
// strings, nested objects). The reference parses params with serde_json (gain.rs:29-36, resampler.rs:21-38,
// mixer.rs:21-79); this is the host-side equivalent for the C++ node mirror.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>

namespace skhost {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    bool is_integer = false;
    std::string str;
    std::map<std::string, std::shared_ptr<JsonValue>> obj;

    const JsonValue *get(const std::string &k) const {
        auto it = obj.find(k);
        return it == obj.end() ? nullptr : it->second.get();
    }
};

class JsonParser {
  public:
    explicit JsonParser(const char *s) : p_(s ? s : "") {}
    bool parse(JsonValue &out, std::string &err) {
        skip();
        if (!value(out, err)) return false;
        skip();
        if (*p_) { err = "trailing characters after JSON value"; return false; }
        return true;
    }

  private:
    const char *p_;
    void skip() { while (*p_ && std::isspace((unsigned char)*p_)) ++p_; }
    bool lit(const char *w) {
        size_t n = 0;
        while (w[n]) ++n;
        for (size_t i = 0; i < n; ++i) if (p_[i] != w[i]) return false;
        p_ += n;
        return true;
    }
    bool string(std::string &out, std::string &err) {
        if (*p_ != '"') { err = "expected string"; return false; }
        ++p_;
        out.clear();
        while (*p_ && *p_ != '"') {
            if (*p_ == '\\') {
                ++p_;
                switch (*p_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': { for (int i = 0; i < 4 && p_[1]; ++i) ++p_; out += '?'; break; }
                    case 0: err = "unterminated escape"; return false;
                    default: out += *p_;
                }
                ++p_;
            } else {
                out += *p_++;
            }
        }
        if (*p_ != '"') { err = "unterminated string"; return false; }
        ++p_;
        return true;
    }
    bool value(JsonValue &v, std::string &err) {
        skip();
        if (*p_ == '{') {
            ++p_;
            v.kind = JsonValue::Object;
            skip();
            if (*p_ == '}') { ++p_; return true; }
            for (;;) {
                skip();
                std::string key;
                if (!string(key, err)) return false;
                skip();
                if (*p_ != ':') { err = "expected ':'"; return false; }
                ++p_;
                auto child = std::make_shared<JsonValue>();
                if (!value(*child, err)) return false;
                v.obj[key] = child;
                skip();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; return true; }
                err = "expected ',' or '}'";
                return false;
            }
        }
        if (*p_ == '"') { v.kind = JsonValue::String; return string(v.str, err); }
        if (lit("null")) { v.kind = JsonValue::Null; return true; }
        if (lit("true")) { v.kind = JsonValue::Bool; v.b = true; return true; }
        if (lit("false")) { v.kind = JsonValue::Bool; v.b = false; return true; }
        if (*p_ == '[') { err = "arrays are not used in node parameters"; return false; }
        char *end = nullptr;
        double d = std::strtod(p_, &end);
        if (end == p_) { err = "unexpected character in JSON"; return false; }
        v.kind = JsonValue::Number;
        v.num = d;
        v.is_integer = true;
        for (const char *q = p_; q < end; ++q) if (*q == '.' || *q == 'e' || *q == 'E') v.is_integer = false;
        p_ = end;
        return true;
    }
};

}  // namespace skhost
