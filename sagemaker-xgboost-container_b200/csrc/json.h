// json.h -- small JSON / UBJSON document model for configs and model files (product code).
// Model schema written/read: SURVEY.md section 8(c) (xgboost 3.x UBJSON/JSON model document), used by
// Booster.save_model / load_model / save_config which the container calls at
// algorithm_mode/train.py:480-485, serve_utils.py:180-193, serve.py:85-88, checkpointing.py:375,428.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <cmath>
#include <cstdio>

namespace b200 {

struct JValue;
using JPtr = std::shared_ptr<JValue>;

struct JValue {
  enum Type { kNull, kBool, kInt, kFloat, kString, kArray, kObject, kF32Array, kI32Array, kI64Array, kU8Array } type = kNull;
  bool b = false; int64_t i = 0; double d = 0; std::string s;
  std::vector<JPtr> arr;
  std::vector<std::pair<std::string, JPtr>> obj;      // insertion-ordered
  std::vector<float> f32; std::vector<int32_t> i32; std::vector<int64_t> i64; std::vector<uint8_t> u8;

  static JPtr Null() { return std::make_shared<JValue>(); }
  static JPtr Bool(bool v) { auto p = std::make_shared<JValue>(); p->type = kBool; p->b = v; return p; }
  static JPtr Int(int64_t v) { auto p = std::make_shared<JValue>(); p->type = kInt; p->i = v; return p; }
  static JPtr Float(double v) { auto p = std::make_shared<JValue>(); p->type = kFloat; p->d = v; return p; }
  static JPtr Str(const std::string& v) { auto p = std::make_shared<JValue>(); p->type = kString; p->s = v; return p; }
  static JPtr Array() { auto p = std::make_shared<JValue>(); p->type = kArray; return p; }
  static JPtr Object() { auto p = std::make_shared<JValue>(); p->type = kObject; return p; }
  static JPtr F32(std::vector<float> v) { auto p = std::make_shared<JValue>(); p->type = kF32Array; p->f32 = std::move(v); return p; }
  static JPtr I32(std::vector<int32_t> v) { auto p = std::make_shared<JValue>(); p->type = kI32Array; p->i32 = std::move(v); return p; }
  static JPtr I64(std::vector<int64_t> v) { auto p = std::make_shared<JValue>(); p->type = kI64Array; p->i64 = std::move(v); return p; }
  static JPtr U8(std::vector<uint8_t> v) { auto p = std::make_shared<JValue>(); p->type = kU8Array; p->u8 = std::move(v); return p; }

  void set(const std::string& k, JPtr v) { for (auto& kv : obj) if (kv.first == k) { kv.second = v; return; } obj.emplace_back(k, v); }
  JPtr get(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return kv.second; return nullptr; }
  const JValue& at(const std::string& k) const { auto p = get(k); if (!p) throw std::runtime_error("model/config: missing key '" + k + "'"); return *p; }
  bool has(const std::string& k) const { return get(k) != nullptr; }

  // numeric views tolerant of how the document was encoded (typed array, generic array, strings)
  double as_double() const {
    switch (type) { case kInt: return (double)i; case kFloat: return d; case kBool: return b ? 1 : 0;
      case kString: { std::string t = s; if (!t.empty() && t.front() == '[') t = t.substr(1, t.size() - 2); return std::strtod(t.c_str(), nullptr); }
      default: throw std::runtime_error("json: not a number"); }
  }
  int64_t as_int() const { return type == kInt ? i : (int64_t)as_double(); }
  size_t length() const {
    switch (type) { case kArray: return arr.size(); case kF32Array: return f32.size(); case kI32Array: return i32.size();
      case kI64Array: return i64.size(); case kU8Array: return u8.size(); default: return 0; }
  }
  double num_at(size_t k) const {
    switch (type) { case kArray: return arr[k]->as_double(); case kF32Array: return f32[k]; case kI32Array: return i32[k];
      case kI64Array: return (double)i64[k]; case kU8Array: return u8[k]; default: throw std::runtime_error("json: not an array"); }
  }
};

// ----------------------------------------------------------------------------------------- JSON text
class JsonReader {
 public:
  explicit JsonReader(const char* p, size_t n) : p_(p), e_(p + n) {}
  JPtr parse() { ws(); JPtr v = value(); return v; }
 private:
  const char* p_; const char* e_; int depth_ = 0;
  static constexpr int kMaxDepth = 64;              // model / config documents nest 7 deep; a hostile file must not exhaust the stack
  struct Nest { int& d; explicit Nest(int& x) : d(x) { if (++d > kMaxDepth) throw std::runtime_error("json parse error: nesting too deep"); } ~Nest() { --d; } };
  void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json parse error: ") + m); }
  bool lit(const char* w, size_t n) { if ((size_t)(e_ - p_) < n || memcmp(p_, w, n) != 0) return false; p_ += n; return true; }     // length first: the buffer need not be terminated
  JPtr value() {
    ws(); if (p_ >= e_) fail("eof");
    char c = *p_;
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') return JValue::Str(string());
    if (lit("true", 4)) return JValue::Bool(true);
    if (lit("false", 5)) return JValue::Bool(false);
    if (lit("null", 4)) return JValue::Null();
    if (lit("NaN", 3)) return JValue::Float(NAN);
    if (lit("Infinity", 8)) return JValue::Float(INFINITY);
    if (lit("-Infinity", 9)) return JValue::Float(-INFINITY);
    return number();
  }
  JPtr number() {
    const char* s = p_; bool isf = false;
    if (p_ < e_ && (*p_ == '-' || *p_ == '+')) ++p_;
    while (p_ < e_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
      if (*p_ == '.' || *p_ == 'e' || *p_ == 'E') isf = true; ++p_; }
    if (s == p_) fail("bad number");
    std::string t(s, p_);
    if (isf) return JValue::Float(std::strtod(t.c_str(), nullptr));
    return JValue::Int(std::strtoll(t.c_str(), nullptr, 10));
  }
  std::string string() {
    ++p_; std::string out;
    while (p_ < e_ && *p_ != '"') {
      if (*p_ == '\\') { ++p_; if (p_ >= e_) fail("eof in string");
        switch (*p_) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case 'b': out += '\b'; break;
          case 'f': out += '\f'; break; case 'u': { unsigned cp = 0; for (int k = 0; k < 4 && p_ + 1 < e_; ++k) { ++p_; const unsigned char h = (unsigned char)*p_; if (!std::isxdigit(h)) fail("bad \\u escape"); cp = cp * 16 + (unsigned)(std::isdigit(h) ? h - '0' : (std::tolower(h) - 'a' + 10)); }
            if (cp < 0x80) out += (char)cp; else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); } break; }
          default: out += *p_; }
        ++p_; }
      else out += *p_++;
    }
    if (p_ >= e_) fail("unterminated string");
    ++p_; return out;
  }
  JPtr array() {
    Nest nest(depth_);
    ++p_; JPtr a = JValue::Array(); ws();
    if (p_ < e_ && *p_ == ']') { ++p_; return a; }
    while (true) { a->arr.push_back(value()); ws(); if (p_ >= e_) fail("eof in array"); if (*p_ == ',') { ++p_; continue; } if (*p_ == ']') { ++p_; break; } fail("expected , or ]"); }
    return a;
  }
  JPtr object() {
    Nest nest(depth_);
    ++p_; JPtr o = JValue::Object(); ws();
    if (p_ < e_ && *p_ == '}') { ++p_; return o; }
    while (true) { ws(); if (p_ >= e_ || *p_ != '"') fail("expected key"); std::string k = string(); ws(); if (p_ >= e_ || *p_ != ':') fail("expected :"); ++p_;
      o->obj.emplace_back(k, value()); ws(); if (p_ >= e_) fail("eof in object"); if (*p_ == ',') { ++p_; continue; } if (*p_ == '}') { ++p_; break; } fail("expected , or }"); }
    return o;
  }
};

inline void json_escape(const std::string& s, std::string* out) {
  out->push_back('"');
  for (unsigned char c : s) {
    switch (c) { case '"': *out += "\\\""; break; case '\\': *out += "\\\\"; break; case '\n': *out += "\\n"; break; case '\t': *out += "\\t"; break;
      case '\r': *out += "\\r"; break; default: if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); *out += b; } else out->push_back((char)c); }
  }
  out->push_back('"');
}
inline void json_float(double v, std::string* out, bool f32) {
  if (std::isnan(v)) { *out += "NaN"; return; }
  if (std::isinf(v)) { *out += v > 0 ? "Infinity" : "-Infinity"; return; }
  char b[40]; snprintf(b, sizeof b, f32 ? "%.9G" : "%.17G", v);
  *out += b;
  if (!strpbrk(b, ".EN")) *out += "E0";     // keep it a float on re-parse
}
inline void json_write(const JValue& v, std::string* out) {
  switch (v.type) {
    case JValue::kNull: *out += "null"; break;
    case JValue::kBool: *out += v.b ? "true" : "false"; break;
    case JValue::kInt: *out += std::to_string(v.i); break;
    case JValue::kFloat: json_float(v.d, out, true); break;
    case JValue::kString: json_escape(v.s, out); break;
    case JValue::kArray: { out->push_back('['); bool first = true; for (auto& x : v.arr) { if (!first) out->push_back(','); first = false; json_write(*x, out); } out->push_back(']'); break; }
    case JValue::kObject: { out->push_back('{'); bool first = true; for (auto& kv : v.obj) { if (!first) out->push_back(','); first = false; json_escape(kv.first, out); out->push_back(':'); json_write(*kv.second, out); } out->push_back('}'); break; }
    case JValue::kF32Array: { out->push_back('['); for (size_t k = 0; k < v.f32.size(); ++k) { if (k) out->push_back(','); json_float(v.f32[k], out, true); } out->push_back(']'); break; }
    case JValue::kI32Array: { out->push_back('['); for (size_t k = 0; k < v.i32.size(); ++k) { if (k) out->push_back(','); *out += std::to_string(v.i32[k]); } out->push_back(']'); break; }
    case JValue::kI64Array: { out->push_back('['); for (size_t k = 0; k < v.i64.size(); ++k) { if (k) out->push_back(','); *out += std::to_string(v.i64[k]); } out->push_back(']'); break; }
    case JValue::kU8Array: { out->push_back('['); for (size_t k = 0; k < v.u8.size(); ++k) { if (k) out->push_back(','); *out += std::to_string((int)v.u8[k]); } out->push_back(']'); break; }
  }
}

// ----------------------------------------------------------------------------------------- UBJSON
template <typename T> inline void be_put(T v, std::string* out) { unsigned char b[sizeof(T)]; memcpy(b, &v, sizeof(T)); for (size_t k = 0; k < sizeof(T); ++k) out->push_back((char)b[sizeof(T) - 1 - k]); }
template <typename T> inline T be_get(const unsigned char* p) { unsigned char b[sizeof(T)]; for (size_t k = 0; k < sizeof(T); ++k) b[k] = p[sizeof(T) - 1 - k]; T v; memcpy(&v, b, sizeof(T)); return v; }

inline void ubj_len(int64_t n, std::string* out) { out->push_back('L'); be_put<int64_t>(n, out); }
inline void ubj_str(const std::string& s, std::string* out) { ubj_len((int64_t)s.size(), out); *out += s; }
inline void ubj_write(const JValue& v, std::string* out) {
  switch (v.type) {
    case JValue::kNull: out->push_back('Z'); break;
    case JValue::kBool: out->push_back(v.b ? 'T' : 'F'); break;
    case JValue::kInt: out->push_back('L'); be_put<int64_t>(v.i, out); break;
    case JValue::kFloat: out->push_back('d'); be_put<float>((float)v.d, out); break;
    case JValue::kString: out->push_back('S'); ubj_str(v.s, out); break;
    case JValue::kArray: out->push_back('['); for (auto& x : v.arr) ubj_write(*x, out); out->push_back(']'); break;
    case JValue::kObject: out->push_back('{'); for (auto& kv : v.obj) { ubj_str(kv.first, out); ubj_write(*kv.second, out); } out->push_back('}'); break;
    case JValue::kF32Array: *out += "[$d#"; ubj_len((int64_t)v.f32.size(), out); for (float x : v.f32) be_put<float>(x, out); break;
    case JValue::kI32Array: *out += "[$l#"; ubj_len((int64_t)v.i32.size(), out); for (int32_t x : v.i32) be_put<int32_t>(x, out); break;
    case JValue::kI64Array: *out += "[$L#"; ubj_len((int64_t)v.i64.size(), out); for (int64_t x : v.i64) be_put<int64_t>(x, out); break;
    case JValue::kU8Array: *out += "[$U#"; ubj_len((int64_t)v.u8.size(), out); for (uint8_t x : v.u8) out->push_back((char)x); break;
  }
}

class UbjReader {
 public:
  UbjReader(const unsigned char* p, size_t n) : p_(p), e_(p + n) {}
  JPtr parse() { return value(take()); }
 private:
  const unsigned char* p_; const unsigned char* e_; int depth_ = 0;
  static constexpr int kMaxDepth = 64;
  struct Nest { int& d; explicit Nest(int& x) : d(x) { if (++d > kMaxDepth) throw std::runtime_error("ubjson parse error: nesting too deep"); } ~Nest() { --d; } };
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("ubjson parse error: ") + m); }
  // a count field is checked against the bytes that are left BEFORE anything is allocated for it
  size_t counted(int64_t cnt, size_t itemsize) { if (cnt < 0 || (uint64_t)cnt > (uint64_t)(e_ - p_) / itemsize) fail("count runs past the end of the buffer"); return (size_t)cnt; }
  unsigned char take() { if (p_ >= e_) fail("eof"); return *p_++; }
  unsigned char peek() { if (p_ >= e_) fail("eof"); return *p_; }
  const unsigned char* bytes(size_t n) { if ((size_t)(e_ - p_) < n) fail("eof"); const unsigned char* r = p_; p_ += n; return r; }
  int64_t integer(unsigned char m) {
    switch (m) { case 'i': return (int8_t)*bytes(1); case 'U': return *bytes(1); case 'I': return be_get<int16_t>(bytes(2));
      case 'l': return be_get<int32_t>(bytes(4)); case 'L': return be_get<int64_t>(bytes(8)); default: fail("bad integer marker"); }
  }
  std::string str() { const size_t n = counted(integer(take()), 1); const unsigned char* b = bytes(n); return std::string((const char*)b, n); }
  JPtr value(unsigned char m) {
    switch (m) {
      case 'Z': return JValue::Null(); case 'T': return JValue::Bool(true); case 'F': return JValue::Bool(false);
      case 'i': case 'U': case 'I': case 'l': case 'L': return JValue::Int(integer(m));
      case 'd': return JValue::Float(be_get<float>(bytes(4))); case 'D': return JValue::Float(be_get<double>(bytes(8)));
      case 'S': return JValue::Str(str()); case 'C': return JValue::Str(std::string(1, (char)take()));
      case '[': return array(); case '{': return object();
      default: fail("bad marker");
    }
  }
  JPtr array() {
    Nest nest(depth_);
    int typ = 0; int64_t cnt = -1;
    if (peek() == '$') { take(); typ = take(); }
    if (peek() == '#') { take(); cnt = integer(take()); if (cnt < 0) fail("negative count"); }
    if (typ) {
      if (cnt < 0) fail("typed array without count");
      const size_t item = typ == 'd' || typ == 'l' ? 4 : typ == 'D' || typ == 'L' ? 8 : typ == 'I' ? 2 : typ == 'U' || typ == 'i' || typ == 'C' ? 1 : 0;
      if (item == 0 && (typ == 'Z' || typ == 'T' || typ == 'F' || typ == 'N') && cnt > (1 << 20)) fail("implausible count of payload-free values");
      size_t n = item ? counted(cnt, item) : (size_t)cnt;
      if (typ == 'd') { std::vector<float> v(n); const unsigned char* b = bytes(4 * n); for (size_t k = 0; k < n; ++k) v[k] = be_get<float>(b + 4 * k); return JValue::F32(std::move(v)); }
      if (typ == 'D') { std::vector<float> v(n); const unsigned char* b = bytes(8 * n); for (size_t k = 0; k < n; ++k) v[k] = (float)be_get<double>(b + 8 * k); return JValue::F32(std::move(v)); }
      if (typ == 'l') { std::vector<int32_t> v(n); const unsigned char* b = bytes(4 * n); for (size_t k = 0; k < n; ++k) v[k] = be_get<int32_t>(b + 4 * k); return JValue::I32(std::move(v)); }
      if (typ == 'L') { std::vector<int64_t> v(n); const unsigned char* b = bytes(8 * n); for (size_t k = 0; k < n; ++k) v[k] = be_get<int64_t>(b + 8 * k); return JValue::I64(std::move(v)); }
      if (typ == 'U') { std::vector<uint8_t> v(n); const unsigned char* b = bytes(n); memcpy(v.data(), b, n); return JValue::U8(std::move(v)); }
      if (typ == 'i') { std::vector<int32_t> v(n); const unsigned char* b = bytes(n); for (size_t k = 0; k < n; ++k) v[k] = (int8_t)b[k]; return JValue::I32(std::move(v)); }
      if (typ == 'I') { std::vector<int32_t> v(n); const unsigned char* b = bytes(2 * n); for (size_t k = 0; k < n; ++k) v[k] = be_get<int16_t>(b + 2 * k); return JValue::I32(std::move(v)); }
      JPtr a = JValue::Array(); for (size_t k = 0; k < n; ++k) a->arr.push_back(value((unsigned char)typ)); return a;
    }
    JPtr a = JValue::Array();
    if (cnt >= 0) { for (int64_t k = 0; k < cnt; ++k) a->arr.push_back(value(take())); return a; }
    while (peek() != ']') a->arr.push_back(value(take()));
    take(); return a;
  }
  JPtr object() {
    Nest nest(depth_);
    int64_t cnt = -1;
    if (peek() == '#') { take(); cnt = integer(take()); if (cnt < 0) fail("negative count"); }
    JPtr o = JValue::Object();
    if (cnt >= 0) { for (int64_t k = 0; k < cnt; ++k) { std::string key = str(); o->obj.emplace_back(key, value(take())); } return o; }
    while (peek() != '}') { std::string key = str(); o->obj.emplace_back(key, value(take())); }
    take(); return o;
  }
};

inline JPtr parse_json(const std::string& s) { return JsonReader(s.data(), s.size()).parse(); }

}  // namespace b200
