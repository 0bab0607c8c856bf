// json.h -- a small JSON reader/writer for the scene / technique files.
// Replaces the reference's use of nlohmann::json 2.1.1 (reflectcuts/json/json.hpp) for the
// few operations the path needs: parse, find(key), operator[], array size, typed get.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace evplp_host {

class Json {
public:
    enum Type { Null, Bool, Number, String, Array, Object };
    Json() : mType(Null) {}
    Type type() const { return mType; }
    bool is_null() const { return mType == Null; }
    bool is_object() const { return mType == Object; }
    bool is_array() const { return mType == Array; }
    bool contains(const std::string& key) const { return mType == Object && mObject.find(key) != mObject.end(); }
    size_t size() const { return mType == Array ? mArray.size() : mType == Object ? mObject.size() : 0; }

    // operator[] on a missing key yields null (nlohmann's const behaviour is UB; the reference only tests is_null()).
    const Json& operator[](const std::string& key) const {
        static const Json kNull;
        if (mType != Object) return kNull;
        auto it = mObject.find(key);
        return it == mObject.end() ? kNull : it->second;
    }
    const Json& operator[](size_t i) const {
        if (mType != Array || i >= mArray.size()) throw std::runtime_error("json: array index out of range");
        return mArray[i];
    }
    double number() const {
        if (mType == Number) return mNumber;
        if (mType == Bool) return mBool ? 1.0 : 0.0;
        throw std::runtime_error("json: value is not a number");
    }
    float as_float() const { return (float)number(); }
    int as_int() const { return (int)number(); }
    bool as_bool() const {
        if (mType == Bool) return mBool;
        if (mType == Number) return mNumber != 0.0;
        throw std::runtime_error("json: value is not a bool");
    }
    const std::string& as_string() const {
        if (mType != String) throw std::runtime_error("json: value is not a string");
        return mString;
    }
    // required key (nlohmann throws on a type mismatch; a missing required key is an error here too)
    const Json& at(const std::string& key) const {
        if (!contains(key)) throw std::runtime_error("json: missing key \"" + key + "\"");
        return mObject.at(key);
    }
    const std::map<std::string, Json>& items() const { return mObject; }

    static Json parse(const std::string& text) {
        size_t pos = 0;
        Json j = parse_value(text, pos);
        skip_ws(text, pos);
        if (pos != text.size()) throw std::runtime_error("json: trailing characters");
        return j;
    }
    static Json parse_file(const std::string& path) {
        std::ifstream ifs(path);
        if (!ifs.is_open()) throw std::runtime_error("json: cannot open " + path);
        std::stringstream ss;
        ss << ifs.rdbuf();
        return parse(ss.str());
    }

    // builders (stat file output)
    static Json make_object() { Json j; j.mType = Object; return j; }
    void set(const std::string& key, double v) { Json n; n.mType = Number; n.mNumber = v; mObject[key] = n; }
    void set(const std::string& key, const std::string& v) { Json n; n.mType = String; n.mString = v; mObject[key] = n; }
    std::string dump(int indent = 4) const {
        std::ostringstream os;
        dump_to(os, indent, 0);
        return os.str();
    }

private:
    Type mType;
    bool mBool = false;
    double mNumber = 0.0;
    std::string mString;
    std::vector<Json> mArray;
    std::map<std::string, Json> mObject;

    static void skip_ws(const std::string& s, size_t& p) {
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++;
    }
    static std::string parse_string(const std::string& s, size_t& p) {
        if (s[p] != '"') throw std::runtime_error("json: expected string");
        p++;
        std::string out;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\') {
                p++;
                if (p >= s.size()) break;
                switch (s[p]) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        unsigned code = (unsigned)strtoul(s.substr(p + 1, 4).c_str(), nullptr, 16);
                        p += 4;
                        if (code < 0x80) out += (char)code;
                        else if (code < 0x800) { out += (char)(0xC0 | (code >> 6)); out += (char)(0x80 | (code & 0x3F)); }
                        else { out += (char)(0xE0 | (code >> 12)); out += (char)(0x80 | ((code >> 6) & 0x3F)); out += (char)(0x80 | (code & 0x3F)); }
                        break;
                    }
                    default: out += s[p];
                }
                p++;
            } else {
                out += s[p++];
            }
        }
        if (p >= s.size()) throw std::runtime_error("json: unterminated string");
        p++;
        return out;
    }
    static Json parse_value(const std::string& s, size_t& p) {
        skip_ws(s, p);
        if (p >= s.size()) throw std::runtime_error("json: unexpected end");
        Json j;
        char c = s[p];
        if (c == '{') {
            j.mType = Object;
            p++;
            skip_ws(s, p);
            if (p < s.size() && s[p] == '}') { p++; return j; }
            while (true) {
                skip_ws(s, p);
                std::string key = parse_string(s, p);
                skip_ws(s, p);
                if (p >= s.size() || s[p] != ':') throw std::runtime_error("json: expected ':'");
                p++;
                j.mObject[key] = parse_value(s, p);
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        } else if (c == '[') {
            j.mType = Array;
            p++;
            skip_ws(s, p);
            if (p < s.size() && s[p] == ']') { p++; return j; }
            while (true) {
                j.mArray.push_back(parse_value(s, p));
                skip_ws(s, p);
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        } else if (c == '"') {
            j.mType = String;
            j.mString = parse_string(s, p);
        } else if (s.compare(p, 4, "true") == 0) {
            j.mType = Bool; j.mBool = true; p += 4;
        } else if (s.compare(p, 5, "false") == 0) {
            j.mType = Bool; j.mBool = false; p += 5;
        } else if (s.compare(p, 4, "null") == 0) {
            p += 4;
        } else {
            const char* start = s.c_str() + p;
            char* end = nullptr;
            double v = strtod(start, &end);
            if (end == start) throw std::runtime_error("json: bad value");
            p += (size_t)(end - start);
            j.mType = Number;
            j.mNumber = v;
        }
        return j;
    }
    void dump_to(std::ostringstream& os, int indent, int depth) const {
        std::string pad((size_t)(indent * (depth + 1)), ' '), padEnd((size_t)(indent * depth), ' ');
        switch (mType) {
            case Null: os << "null"; break;
            case Bool: os << (mBool ? "true" : "false"); break;
            case Number: { std::ostringstream t; t.precision(17); t << mNumber; os << t.str(); break; }
            case String: os << '"' << mString << '"'; break;
            case Array: {
                os << "[\n";
                for (size_t i = 0; i < mArray.size(); i++) { os << pad; mArray[i].dump_to(os, indent, depth + 1); os << (i + 1 < mArray.size() ? ",\n" : "\n"); }
                os << padEnd << "]";
                break;
            }
            case Object: {
                os << "{\n";
                size_t i = 0;
                for (auto& kv : mObject) { os << pad << '"' << kv.first << "\": "; kv.second.dump_to(os, indent, depth + 1); os << (++i < mObject.size() ? ",\n" : "\n"); }
                os << padEnd << "}";
                break;
            }
        }
    }
};

}  // namespace evplp_host
