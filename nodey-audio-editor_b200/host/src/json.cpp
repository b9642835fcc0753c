// json.cpp -- implementation of the Json:: subset declared in shim/json/json.h
#include <json/json.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace Json {

const Value& Value::nullSingleton()
{
    static const Value v;
    return v;
}

std::string Value::asString() const
{
    switch (type_) {
    case nullValue: return "";
    case stringValue: return str_;
    case booleanValue: return bool_ ? "true" : "false";
    case intValue: case uintValue: return std::to_string(int_);
    case realValue: { char b[64]; snprintf(b, sizeof(b), "%.17g", real_); return b; }
    default: throw LogicError("Type is not convertible to string");
    }
}

std::int64_t Value::asInt64() const
{
    switch (type_) {
    case nullValue: return 0;
    case intValue: case uintValue: return int_;
    case realValue:
        if (!(real_ >= -9.3e18 && real_ <= 9.3e18)) throw LogicError("double out of Int64 range");
        return (std::int64_t)real_;
    case booleanValue: return bool_ ? 1 : 0;
    default: throw LogicError("Value is not convertible to Int64.");
    }
}

int Value::asInt() const
{
    if (type_ == stringValue || type_ == arrayValue || type_ == objectValue) throw LogicError("Value is not convertible to Int.");
    const std::int64_t v = asInt64();
    if (v < std::numeric_limits<int>::min() || v > std::numeric_limits<int>::max()) throw LogicError("LargestInt out of Int range");
    return (int)v;
}

unsigned Value::asUInt() const
{
    const std::int64_t v = asInt64();
    if (v < 0 || v > (std::int64_t)std::numeric_limits<unsigned>::max()) throw LogicError("LargestInt out of UInt range");
    return (unsigned)v;
}

double Value::asDouble() const
{
    switch (type_) {
    case nullValue: return 0.0;
    case intValue: case uintValue: return (double)int_;
    case realValue: return real_;
    case booleanValue: return bool_ ? 1.0 : 0.0;
    default: throw LogicError("Value is not convertible to double.");
    }
}

bool Value::asBool() const
{
    switch (type_) {
    case nullValue: return false;
    case booleanValue: return bool_;
    case intValue: case uintValue: return int_ != 0;
    case realValue: return real_ != 0.0 && !std::isnan(real_);
    default: throw LogicError("Value is not convertible to bool.");
    }
}

Value& Value::operator[](const std::string& key)
{
    if (type_ == nullValue) type_ = objectValue;
    if (type_ != objectValue) throw LogicError("in Json::Value::operator[](key): requires objectValue");
    return obj_[key];
}

const Value& Value::operator[](const std::string& key) const
{
    if (type_ == nullValue) return nullSingleton();
    if (type_ != objectValue) throw LogicError("in Json::Value::operator[](key) const: requires objectValue");
    const auto it = obj_.find(key);
    return it == obj_.end() ? nullSingleton() : it->second;
}

Value::Members Value::getMemberNames() const
{
    if (type_ == nullValue) return {};
    if (type_ != objectValue) throw LogicError("in Json::Value::getMemberNames(), value must be objectValue");
    Members m;
    for (const auto& kv : obj_) m.push_back(kv.first);
    return m;
}

Value& Value::operator[](int index)
{
    if (type_ == nullValue) type_ = arrayValue;
    if (type_ != arrayValue || index < 0) throw LogicError("in Json::Value::operator[](index): requires arrayValue");
    if ((size_t)index >= arr_.size()) arr_.resize((size_t)index + 1);
    return arr_[(size_t)index];
}

const Value& Value::operator[](int index) const
{
    if (type_ == nullValue) return nullSingleton();
    if (type_ != arrayValue || index < 0) throw LogicError("in Json::Value::operator[](index) const: requires arrayValue");
    return (size_t)index < arr_.size() ? arr_[(size_t)index] : nullSingleton();
}

Value& Value::append(const Value& v)
{
    if (type_ == nullValue) type_ = arrayValue;
    if (type_ != arrayValue) throw LogicError("in Json::Value::append: requires arrayValue");
    arr_.push_back(v);
    return arr_.back();
}

Value& Value::append(Value&& v)
{
    if (type_ == nullValue) type_ = arrayValue;
    if (type_ != arrayValue) throw LogicError("in Json::Value::append: requires arrayValue");
    arr_.push_back(std::move(v));
    return arr_.back();
}

void Value::flatten() const
{
    flat_.clear();
    if (type_ == objectValue) for (const auto& kv : obj_) flat_.push_back(kv.second);
}

bool Value::operator==(const Value& o) const
{
    if (isDouble() && o.isDouble() && (type_ == realValue || o.type_ == realValue)) return asDouble() == o.asDouble();
    if (type_ != o.type_) {
        const bool ints = (type_ == intValue || type_ == uintValue) && (o.type_ == intValue || o.type_ == uintValue);
        if (!ints) return false;
    }
    switch (type_) {
    case nullValue: return true;
    case intValue: case uintValue: return int_ == o.int_;
    case realValue: return real_ == o.real_;
    case booleanValue: return bool_ == o.bool_;
    case stringValue: return str_ == o.str_;
    case arrayValue: return arr_ == o.arr_;
    case objectValue: return obj_ == o.obj_;
    }
    return false;
}

// ---- writer ----------------------------------------------------------------------------------------
static void write_string(std::string& out, const std::string& s)
{
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\b': out += "\\b"; break;
        case '\f': out += "\\f"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
            if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; }
            else out += (char)c;
        }
    }
    out += '"';
}

static std::string real_to_string(double v)
{
    if (std::isnan(v)) return "null";
    if (std::isinf(v)) return v > 0 ? "1e+9999" : "-1e+9999";
    char b[64];
    snprintf(b, sizeof(b), "%.17g", v);
    // shortest representation that round-trips
    for (int prec = 1; prec < 17; prec++) {
        char t[64];
        snprintf(t, sizeof(t), "%.*g", prec, v);
        if (strtod(t, nullptr) == v) { memcpy(b, t, sizeof(t)); break; }
    }
    std::string s = b;
    if (s.find_first_of(".eE") == std::string::npos && s.find("inf") == std::string::npos) s += ".0";
    return s;
}

struct WriterImpl {
    std::string indent;      // empty => compact
    std::string out;
    void value(const Value& v, int depth, ValueType t, std::int64_t i, double r, bool b, const std::string& s,
               const std::vector<Value>& arr, const std::map<std::string, Value>& obj);
};

static void write_value(std::string& out, const Value& v, const std::string& indent, int depth);

static void newline(std::string& out, const std::string& indent, int depth)
{
    if (indent.empty()) return;
    out += '\n';
    for (int i = 0; i < depth; i++) out += indent;
}

static void write_value(std::string& out, const Value& v, const std::string& indent, int depth)
{
    switch (v.type()) {
    case nullValue: out += "null"; break;
    case intValue: case uintValue: out += std::to_string(v.asInt64()); break;
    case realValue: out += real_to_string(v.asDouble()); break;
    case booleanValue: out += v.asBool() ? "true" : "false"; break;
    case stringValue: write_string(out, v.asString()); break;
    case arrayValue: {
        if (v.size() == 0) { out += "[]"; break; }
        out += '[';
        bool first = true;
        for (const Value& e : v) {
            if (!first) out += ',';
            first = false;
            newline(out, indent, depth + 1);
            write_value(out, e, indent, depth + 1);
        }
        newline(out, indent, depth);
        out += ']';
        break;
    }
    case objectValue: {
        const auto names = v.getMemberNames();
        if (names.empty()) { out += "{}"; break; }
        out += '{';
        bool first = true;
        for (const auto& k : names) {
            if (!first) out += ',';
            first = false;
            newline(out, indent, depth + 1);
            write_string(out, k);
            out += indent.empty() ? ":" : " : ";
            write_value(out, v[k], indent, depth + 1);
        }
        newline(out, indent, depth);
        out += '}';
        break;
    }
    }
}

std::string writeString(const Value& root, const std::string& indentation)
{
    std::string out;
    write_value(out, root, indentation, 0);
    if (!indentation.empty()) out += '\n';
    return out;
}

std::string Value::toStyledString() const { return writeString(*this, "   "); }
std::string StyledWriter::write(const Value& root) const { return writeString(root, "   "); }
std::string FastWriter::write(const Value& root) const { return writeString(root, "") + "\n"; }

// ---- reader ----------------------------------------------------------------------------------------
namespace {
struct Parser {
    const char* p; const char* end; std::string err;
    int depth = 0;

    void skip()
    {
        for (;;) {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
            if (p + 1 < end && p[0] == '/' && p[1] == '/') { while (p < end && *p != '\n') p++; continue; }
            if (p + 1 < end && p[0] == '/' && p[1] == '*') {
                p += 2;
                while (p + 1 < end && !(p[0] == '*' && p[1] == '/')) p++;
                p = p + 2 <= end ? p + 2 : end;
                continue;
            }
            break;
        }
    }
    bool fail(const std::string& m) { if (err.empty()) err = m + " at offset " + std::to_string((long)(end - p)); return false; }

    static void put_utf8(std::string& s, unsigned cp)
    {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
        else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    }
    bool hex4(unsigned& v)
    {
        if (end - p < 4) return false;
        v = 0;
        for (int i = 0; i < 4; i++) {
            const char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else return false;
        }
        return true;
    }
    bool string(std::string& s)
    {
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                const char c = *p++;
                switch (c) {
                case '"': s += '"'; break; case '\\': s += '\\'; break; case '/': s += '/'; break;
                case 'b': s += '\b'; break; case 'f': s += '\f'; break; case 'n': s += '\n'; break;
                case 'r': s += '\r'; break; case 't': s += '\t'; break;
                case 'u': {
                    unsigned cp;
                    if (!hex4(cp)) return fail("bad unicode escape");
                    if (cp >= 0xD800 && cp <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        unsigned lo;
                        if (!hex4(lo)) return fail("bad unicode escape");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    put_utf8(s, cp);
                    break;
                }
                default: return fail("bad escape");
                }
            } else s += *p++;
        }
        if (p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    bool value(Value& out)
    {
        if (++depth > 512) return fail("nesting too deep");
        skip();
        if (p >= end) return fail("unexpected end of document");
        bool ok = true;
        const char c = *p;
        if (c == '{') {
            p++;
            out = Value(objectValue);
            skip();
            if (p < end && *p == '}') p++;
            else for (;;) {
                skip();
                std::string key;
                if (!string(key)) { ok = false; break; }
                skip();
                if (p >= end || *p != ':') { ok = fail("expected ':'"); break; }
                p++;
                Value v;
                if (!value(v)) { ok = false; break; }
                out[key] = std::move(v);
                skip();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = fail("expected ',' or '}'");
                break;
            }
        } else if (c == '[') {
            p++;
            out = Value(arrayValue);
            skip();
            if (p < end && *p == ']') p++;
            else for (;;) {
                Value v;
                if (!value(v)) { ok = false; break; }
                out.append(std::move(v));
                skip();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = fail("expected ',' or ']'");
                break;
            }
        } else if (c == '"') {
            std::string s;
            ok = string(s);
            if (ok) out = Value(s);
        } else if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; out = Value(true); }
        else if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; out = Value(false); }
        else if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; out = Value(); }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s = p;
            bool real = false;
            if (*p == '-') p++;
            while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
                if (*p == '.' || *p == 'e' || *p == 'E') real = true;
                p++;
            }
            const std::string tok(s, p);
            if (!real) {
                errno = 0;
                char* e = nullptr;
                const long long v = strtoll(tok.c_str(), &e, 10);
                if (errno == 0 && e && *e == 0) out = Value((std::int64_t)v); else real = true;
            }
            if (real) {
                char* e = nullptr;
                const double v = strtod(tok.c_str(), &e);
                if (!e || *e != 0) ok = fail("bad number");
                else out = Value(v);
            }
        } else ok = fail("unexpected character");
        depth--;
        return ok;
    }
};
}  // namespace

bool Reader::parse(const std::string& document, Value& root, bool)
{
    Parser ps{document.data(), document.data() + document.size(), "", 0};
    Value v;
    if (!ps.value(v)) { error_ = ps.err; return false; }
    ps.skip();
    if (ps.p != ps.end) { error_ = "extra characters after the document"; return false; }
    root = std::move(v);
    error_.clear();
    return true;
}

}  // namespace Json
