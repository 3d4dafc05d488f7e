// kb2_json.h — minimal JSON object reader for the flat config objects Knowhere passes
// (reference: include/knowhere/config.h:587-700 — every index parameter is a top-level scalar).
// Accepts {"key": number|string|true|false|null, ...}; nested values are skipped.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <string>

namespace kb2 {

class JsonObj {
 public:
    std::map<std::string, std::string> kv;  // raw token text (strings unquoted)
    bool ok = true;

    static JsonObj
    parse(const char* s) {
        JsonObj o;
        if (!s) return o;
        const char* p = s;
        skip_ws(p);
        if (*p == 0) return o;
        if (*p != '{') { o.ok = false; return o; }
        p++;
        for (;;) {
            skip_ws(p);
            if (*p == '}') break;
            if (*p != '"') { o.ok = false; return o; }
            std::string key = read_string(p, o.ok);
            if (!o.ok) return o;
            skip_ws(p);
            if (*p != ':') { o.ok = false; return o; }
            p++;
            skip_ws(p);
            std::string val;
            if (*p == '"') {
                val = read_string(p, o.ok);
                if (!o.ok) return o;
            } else if (*p == '{' || *p == '[') {
                skip_nested(p, o.ok);
                if (!o.ok) return o;
                val = "";
            } else {
                const char* b = p;
                while (*p && *p != ',' && *p != '}' && !isspace((unsigned char)*p)) p++;
                val.assign(b, p);
            }
            o.kv[key] = val;
            skip_ws(p);
            if (*p == ',') { p++; continue; }
            if (*p == '}') break;
            o.ok = false;
            return o;
        }
        return o;
    }
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    long long
    get_int(const std::string& k, long long dflt) const {
        auto it = kv.find(k);
        if (it == kv.end() || it->second.empty()) return dflt;
        return (long long)strtod(it->second.c_str(), nullptr);
    }
    double
    get_num(const std::string& k, double dflt) const {
        auto it = kv.find(k);
        if (it == kv.end() || it->second.empty()) return dflt;
        return strtod(it->second.c_str(), nullptr);
    }
    bool
    get_bool(const std::string& k, bool dflt) const {
        auto it = kv.find(k);
        if (it == kv.end()) return dflt;
        const std::string& v = it->second;
        if (v == "true" || v == "True" || v == "1") return true;
        if (v == "false" || v == "False" || v == "0") return false;
        return dflt;
    }
    std::string
    get_str(const std::string& k, const std::string& dflt) const {
        auto it = kv.find(k);
        return it == kv.end() ? dflt : it->second;
    }

 private:
    static void skip_ws(const char*& p) { while (*p && isspace((unsigned char)*p)) p++; }
    static std::string
    read_string(const char*& p, bool& ok) {
        std::string out;
        p++;  // opening quote
        while (*p && *p != '"') {
            if (*p == '\\' && p[1]) { p++; }
            out.push_back(*p++);
        }
        if (*p != '"') { ok = false; return out; }
        p++;
        return out;
    }
    static void
    skip_nested(const char*& p, bool& ok) {
        int depth = 0;
        bool in_str = false;
        while (*p) {
            char c = *p++;
            if (in_str) {
                if (c == '\\' && *p) p++;
                else if (c == '"') in_str = false;
                continue;
            }
            if (c == '"') in_str = true;
            else if (c == '{' || c == '[') depth++;
            else if (c == '}' || c == ']') { depth--; if (depth == 0) return; }
        }
        ok = false;
    }
};

}  // namespace kb2
