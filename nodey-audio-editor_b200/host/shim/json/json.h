// json/json.h -- a small, self-contained subset of the jsoncpp 1.9 API (Json::Value, Json::Reader,
// Json::StyledWriter, Json::Exception): exactly what the processor plugin API and the project file
// format need.  Written for this engine; not jsoncpp code.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Json {

enum ValueType { nullValue = 0, intValue, uintValue, realValue, stringValue, booleanValue, arrayValue, objectValue };

class Exception : public std::runtime_error {
  public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class LogicError : public Exception {
  public:
    explicit LogicError(const std::string& m) : Exception(m) {}
};

class Value {
  public:
    using Members = std::vector<std::string>;
    using Int = int;
    using UInt = unsigned;
    using ArrayIndex = unsigned;

    Value(ValueType t = nullValue) : type_(t) {}
    Value(int v) : type_(intValue), int_(v) {}
    Value(unsigned v) : type_(uintValue), int_((std::int64_t)v) {}
    Value(std::int64_t v) : type_(intValue), int_(v) {}
    Value(std::uint64_t v) : type_(uintValue), int_((std::int64_t)v) {}
    Value(double v) : type_(realValue), real_(v) {}
    Value(float v) : type_(realValue), real_((double)v) {}
    Value(bool v) : type_(booleanValue), bool_(v) {}
    Value(const char* s) : type_(stringValue), str_(s) {}
    Value(const std::string& s) : type_(stringValue), str_(s) {}

    ValueType type() const { return type_; }
    bool isNull() const { return type_ == nullValue; }
    bool isBool() const { return type_ == booleanValue; }
    bool isInt() const { return type_ == intValue || type_ == uintValue || (type_ == realValue && real_ == (double)(std::int64_t)real_); }
    bool isIntegral() const { return isInt(); }
    bool isDouble() const { return type_ == realValue || type_ == intValue || type_ == uintValue; }
    bool isNumeric() const { return isDouble(); }
    bool isString() const { return type_ == stringValue; }
    bool isArray() const { return type_ == arrayValue; }
    bool isObject() const { return type_ == objectValue; }

    std::string asString() const;
    int asInt() const;
    unsigned asUInt() const;
    std::int64_t asInt64() const;
    double asDouble() const;
    float asFloat() const { return (float)asDouble(); }
    bool asBool() const;

    // object access (a null value silently becomes an object / array, as in jsoncpp)
    Value& operator[](const std::string& key);
    Value& operator[](const char* key) { return (*this)[std::string(key)]; }
    const Value& operator[](const std::string& key) const;
    const Value& operator[](const char* key) const { return (*this)[std::string(key)]; }
    bool isMember(const std::string& key) const { return type_ == objectValue && obj_.count(key) != 0; }
    Members getMemberNames() const;
    void removeMember(const std::string& key) { obj_.erase(key); }

    // array access
    Value& operator[](int index);
    Value& operator[](unsigned index) { return (*this)[(int)index]; }
    const Value& operator[](int index) const;
    const Value& operator[](unsigned index) const { return (*this)[(int)index]; }
    Value& append(const Value& v);
    Value& append(Value&& v);
    unsigned size() const { return type_ == arrayValue ? (unsigned)arr_.size() : type_ == objectValue ? (unsigned)obj_.size() : 0u; }
    bool empty() const { return size() == 0; }

    // iteration over array elements (objects iterate their values in key order)
    std::vector<Value>::const_iterator begin() const { flatten(); return type_ == arrayValue ? arr_.begin() : flat_.begin(); }
    std::vector<Value>::const_iterator end() const { return type_ == arrayValue ? arr_.end() : flat_.end(); }

    bool operator==(const Value& o) const;
    bool operator!=(const Value& o) const { return !(*this == o); }

    std::string toStyledString() const;
    static const Value& nullSingleton();

  private:
    void flatten() const;
    friend class StyledWriter;
    friend class FastWriter;
    ValueType type_ = nullValue;
    std::int64_t int_ = 0;
    double real_ = 0;
    bool bool_ = false;
    std::string str_;
    std::vector<Value> arr_;
    std::map<std::string, Value> obj_;
    mutable std::vector<Value> flat_;
};

class Reader {
  public:
    bool parse(const std::string& document, Value& root, bool collectComments = true);
    std::string getFormattedErrorMessages() const { return error_; }

  private:
    std::string error_;
};

class StyledWriter {
  public:
    std::string write(const Value& root) const;   // 3-space indent like jsoncpp's StyledWriter
};

class FastWriter {
  public:
    std::string write(const Value& root) const;   // one line
};

// indentation-configurable writer (the reference saves projects with a 2-space indent, app.cpp:837-839)
std::string writeString(const Value& root, const std::string& indentation);

}  // namespace Json
