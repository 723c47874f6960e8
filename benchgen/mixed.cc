// mixed.cc - configs[4] of BASELINE.json: mixed replay over many methods with Zipf-distributed sizes.
//
// Benchmark infrastructure: neither product nor oracle.  The Python side (benchgen.mixed) reads the
// FileDescriptorSet and hands this generator a flat "plan" of the messages; for every call the generator
// draws a method (Zipf, s = 1.1) and a target size (Zipf, s = 1.2, over 64 B ... 64 KiB), fills the method's
// input message and output message with random values until the target is met, and writes
//   request : canonical `arguments` JSON as encoding/json.Marshal prints a map[string]interface{}
//             (compact, keys sorted bytewise, HTML-safe escapes, numbers in float64 shortest form)
//   reply   : protobuf wire bytes as a generated-code backend would send them (fields by number, packed
//             repeated scalars, implicit-presence zeros omitted, map entries in insertion order).
// Deterministic (splitmix64 seeded per item).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
  uint32_t range(uint32_t lo, uint32_t hi) { return lo + below(hi - lo + 1); }
  double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

// FieldDescriptorProto.Type
enum { T_DOUBLE = 1, T_FLOAT, T_INT64, T_UINT64, T_INT32, T_FIXED64, T_FIXED32, T_BOOL, T_STRING, T_GROUP, T_MESSAGE, T_BYTES,
       T_UINT32, T_ENUM, T_SFIXED32, T_SFIXED64, T_SINT32, T_SINT64 };
enum { FL_REPEATED = 1, FL_PACKED = 2, FL_MAP = 4, FL_ONEOF = 8, FL_PRESENCE = 16, FL_TIMESTAMP = 32 };

struct Field {
  uint32_t number, type, flags;
  int32_t oneof, child;
  std::string name, json_name;
};
struct Msg { std::vector<Field> fields; };
struct EnumV { int32_t number; std::string name; };
struct Plan {
  std::vector<Msg> msgs;
  std::vector<std::vector<EnumV>> enums;
  std::vector<std::pair<uint32_t, uint32_t>> methods;
};

struct Cur {
  const uint8_t* p; const uint8_t* e;
  uint32_t u32() { uint32_t v = 0; if (p + 4 <= e) { memcpy(&v, p, 4); p += 4; } return v; }
  std::string str() { uint32_t n = u32(); std::string s; if (p + n <= e) { s.assign((const char*)p, n); p += n; } return s; }
};
Plan parse_plan(const uint8_t* b, uint64_t n) {
  Plan P;
  Cur c{b, b + n};
  uint32_t nm = c.u32();
  P.msgs.resize(nm);
  for (auto& m : P.msgs) {
    uint32_t nf = c.u32();
    m.fields.resize(nf);
    for (auto& f : m.fields) {
      f.number = c.u32(); f.type = c.u32(); f.flags = c.u32(); f.oneof = (int32_t)c.u32(); f.child = (int32_t)c.u32();
      f.name = c.str(); f.json_name = c.str();
    }
  }
  uint32_t ne = c.u32();
  P.enums.resize(ne);
  for (auto& e : P.enums) {
    uint32_t nv = c.u32();
    e.resize(nv);
    for (auto& v : e) { v.number = (int32_t)c.u32(); v.name = c.str(); }
  }
  uint32_t nmeth = c.u32();
  for (uint32_t i = 0; i < nmeth; i++) { uint32_t a = c.u32(), b2 = c.u32(); P.methods.push_back({a, b2}); }
  return P;
}

// ---- values -------------------------------------------------------------------------------------
struct Val {
  uint64_t u = 0;       // integers (two's complement), bool, enum number, float bits
  double d = 0;         // float / double value
  std::string s;        // string / bytes
  std::vector<std::pair<const Field*, std::vector<Val>>> fields;  // message: set fields, each with 1..n values
  std::vector<Val> kv;  // map entry: {key, value}
};

std::string rand_text(Rng& r, uint32_t len) {
  static const char A[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 _-.,:;!?@#$%^*()[]{}+=~|'/";
  static const char* M[] = {"\xC3\xA9", "\xC3\xB6", "\xC3\xB1", "\xE5\xBC\xA0", "\xE4\xB8\x89", "\xE6\x97\xA5", "\xE2\x82\xAC", "\xC2\xA0"};
  double kind = r.unit();
  std::string s;
  while (s.size() < len) {
    double u = r.unit();
    if (kind >= 0.90 && kind < 0.95 && u < 0.12) {
      static const char E[] = {'"', '\\', '\n', '<', '\t'};
      s.push_back(E[r.below(5)]);
    } else if (kind >= 0.95 && u < 0.25) {
      s += M[r.below(8)];
    } else {
      s.push_back(A[r.below(sizeof(A) - 1)]);
    }
  }
  return s;
}

struct Gen {
  const Plan& P;
  Rng& r;
  Gen(const Plan& p, Rng& rr) : P(p), r(rr) {}

  uint64_t rand_int(int bits, bool sign) {
    double u = r.unit();
    uint64_t v;
    if (u < 0.45) v = r.below(128);
    else if (u < 0.75) v = r.next() & 0x7FFFFFFFull;
    else if (u < 0.9 || bits == 32) v = r.next() & (bits == 32 ? 0xFFFFFFFFull : 0xFFFFFFFFFFFFull);
    else v = r.next();
    if (bits == 32) v &= 0xFFFFFFFFull;
    if (sign) {
      if (bits == 32) v = (uint64_t)(int64_t)(int32_t)(uint32_t)v;
      if (r.unit() < 0.3) v = (uint64_t)(0 - (int64_t)(v & 0x7FFFFFFFFFFFFFFFull));
      if (bits == 32) v = (uint64_t)(int64_t)(int32_t)(uint32_t)v;
    }
    return v;
  }
  Val scalar(const Field& f, uint32_t budget) {
    Val v;
    switch (f.type) {
      case T_INT32: case T_SINT32: case T_SFIXED32: v.u = rand_int(32, true); break;
      case T_UINT32: case T_FIXED32: v.u = rand_int(32, false); break;
      case T_INT64: case T_SINT64: case T_SFIXED64: v.u = rand_int(64, true); break;
      case T_UINT64: case T_FIXED64: v.u = rand_int(64, false); break;
      case T_BOOL: v.u = r.below(2); break;
      case T_ENUM: { const auto& e = P.enums[f.child]; v.u = (uint64_t)(int64_t)e[r.below((uint32_t)e.size())].number; break; }
      case T_FLOAT: case T_DOUBLE: {
        // multiples of 1/8 below 2^20: exact in float32 and float64, shortest decimal form has at most 3 decimals
        int64_t k = (int64_t)r.below(1u << 23) - (1 << 22);
        if (r.unit() < 0.2) k = 0;
        v.d = (double)k / 8.0;
        break;
      }
      case T_STRING: v.s = rand_text(r, budget ? r.range(budget / 2, budget) : r.range(0, 24)); break;
      case T_BYTES: {
        uint32_t n = budget ? r.range(budget / 2, budget) : r.range(0, 24);
        v.s.resize(n);
        for (uint32_t i = 0; i < n; i++) v.s[i] = (char)r.next();
        break;
      }
    }
    return v;
  }
  // message of type mi filling roughly `budget` bytes of text
  Val message(uint32_t mi, uint32_t budget, int depth) {
    Val m;
    const Msg& M = P.msgs[mi];
    // which member of each oneof
    std::map<int32_t, const Field*> pick;
    for (auto& f : M.fields)
      if (f.flags & FL_ONEOF) {
        if (!pick.count(f.oneof) || r.below(3) == 0) pick[f.oneof] = &f;
      }
    // bulk fields share the budget
    std::vector<const Field*> bulk;
    for (auto& f : M.fields) {
      if ((f.flags & FL_ONEOF) && pick[f.oneof] != &f) continue;
      const bool b = f.type == T_STRING || f.type == T_BYTES || (f.flags & (FL_REPEATED | FL_MAP)) || (f.type == T_MESSAGE && !(f.flags & FL_TIMESTAMP));
      if (b) bulk.push_back(&f);
    }
    const uint32_t share = bulk.empty() ? 0 : budget / (uint32_t)bulk.size();
    for (auto& f : M.fields) {
      if ((f.flags & FL_ONEOF) && pick[f.oneof] != &f) continue;
      const bool is_bulk = std::find(bulk.begin(), bulk.end(), &f) != bulk.end();
      if (!is_bulk && !(f.flags & FL_ONEOF) && r.unit() < 0.25) continue;  // unset
      if (is_bulk && share < 8 && r.unit() < 0.5) continue;
      std::vector<Val> vals;
      if (f.flags & FL_TIMESTAMP) {
        Val t;
        t.u = 1500000000ull + r.below(300000000u);
        static const uint32_t nanos[] = {0, 0, 120000000, 123456000, 123456789};
        t.d = (double)nanos[r.below(5)];
        vals.push_back(t);
      } else if (f.flags & FL_MAP) {
        const Msg& E = P.msgs[f.child];
        const Field& kf = E.fields[0];
        const Field& vf = E.fields[1];
        uint32_t per = vf.type == T_STRING ? 48 : (vf.type == T_MESSAGE ? 64 : 16);
        uint32_t cnt = std::min<uint32_t>(std::max<uint32_t>(share / per, 1), 400);
        std::map<std::string, bool> seen;
        for (uint32_t i = 0; i < cnt; i++) {
          Val e;
          Val k = kf.type == T_STRING ? scalar(kf, 0) : scalar(kf, 0);
          if (kf.type == T_STRING) { k.s = "k" + std::to_string(i) + "_" + rand_text(r, r.range(2, 8)); }
          std::string id = kf.type == T_STRING ? k.s : std::to_string(k.u);
          if (seen.count(id)) continue;
          seen[id] = true;
          Val v = vf.type == T_MESSAGE ? (depth < 3 ? message((uint32_t)vf.child, 40, depth + 1) : Val()) : scalar(vf, vf.type == T_STRING ? 32 : 0);
          e.kv.push_back(k);
          e.kv.push_back(v);
          vals.push_back(e);
        }
      } else if (f.flags & FL_REPEATED) {
        uint32_t per = (f.type == T_STRING || f.type == T_BYTES) ? 40 : (f.type == T_MESSAGE ? 80 : 6);
        uint32_t cnt = std::min<uint32_t>(std::max<uint32_t>(share / per, 1), 2000);
        if (f.type == T_MESSAGE && depth >= 3) cnt = 0;
        for (uint32_t i = 0; i < cnt; i++)
          vals.push_back(f.type == T_MESSAGE ? message((uint32_t)f.child, 64, depth + 1) : scalar(f, (f.type == T_STRING || f.type == T_BYTES) ? 32 : 0));
      } else if (f.type == T_MESSAGE) {
        if (depth >= 3) continue;
        vals.push_back(message((uint32_t)f.child, share, depth + 1));
      } else {
        vals.push_back(scalar(f, (f.type == T_STRING || f.type == T_BYTES) ? share : 0));
      }
      if (!vals.empty() || (f.flags & (FL_REPEATED | FL_MAP))) m.fields.push_back({&f, vals});
    }
    return m;
  }
};

// ---- JSON (request side) ----------------------------------------------------------------------
void json_string(std::string& o, const std::string& s) {
  static const char hex[] = "0123456789abcdef";
  o.push_back('"');
  for (size_t i = 0; i < s.size(); i++) {
    unsigned char c = (unsigned char)s[i];
    if (c >= 0x80) o.push_back((char)c);
    else if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back((char)c); }
    else if (c == '\n') o += "\\n";
    else if (c == '\r') o += "\\r";
    else if (c == '\t') o += "\\t";
    else if (c == '\b') o += "\\b";
    else if (c == '\f') o += "\\f";
    else if (c < 0x20 || c == '<' || c == '>' || c == '&') { o += "\\u00"; o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
    else o.push_back((char)c);
  }
  o.push_back('"');
}
void base64(std::string& o, const std::string& s) {
  static const char A[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  size_t i = 0;
  for (; i + 3 <= s.size(); i += 3) {
    uint32_t v = ((uint8_t)s[i] << 16) | ((uint8_t)s[i + 1] << 8) | (uint8_t)s[i + 2];
    o.push_back(A[v >> 18]); o.push_back(A[(v >> 12) & 63]); o.push_back(A[(v >> 6) & 63]); o.push_back(A[v & 63]);
  }
  if (i + 1 == s.size()) {
    uint32_t v = (uint8_t)s[i] << 16;
    o.push_back(A[v >> 18]); o.push_back(A[(v >> 12) & 63]); o += "==";
  } else if (i + 2 == s.size()) {
    uint32_t v = ((uint8_t)s[i] << 16) | ((uint8_t)s[i + 1] << 8);
    o.push_back(A[v >> 18]); o.push_back(A[(v >> 12) & 63]); o.push_back(A[(v >> 6) & 63]); o.push_back('=');
  }
}
void json_eighth(std::string& o, double d) {  // k / 8: at most three decimals, no exponent in this range
  char buf[48];
  snprintf(buf, sizeof buf, "%.3f", d);
  std::string t = buf;
  while (!t.empty() && t.back() == '0') t.pop_back();
  if (!t.empty() && t.back() == '.') t.pop_back();
  if (t == "-0") t = "0";
  o += t;
}
void json_scalar(std::string& o, const Plan& P, const Field& f, const Val& v, Rng& r, bool as_key) {
  char buf[48];
  switch (f.type) {
    case T_INT32: case T_SINT32: case T_SFIXED32: {
      snprintf(buf, sizeof buf, "%d", (int32_t)(uint32_t)v.u);
      if (as_key) { o.push_back('"'); o += buf; o.push_back('"'); } else o += buf;
      break;
    }
    case T_UINT32: case T_FIXED32: {
      snprintf(buf, sizeof buf, "%u", (uint32_t)v.u);
      if (as_key) { o.push_back('"'); o += buf; o.push_back('"'); } else o += buf;
      break;
    }
    case T_INT64: case T_SINT64: case T_SFIXED64: case T_UINT64: case T_FIXED64: {
      const bool sg = f.type == T_INT64 || f.type == T_SINT64 || f.type == T_SFIXED64;
      if (sg) snprintf(buf, sizeof buf, "%lld", (long long)(int64_t)v.u);
      else snprintf(buf, sizeof buf, "%llu", (unsigned long long)v.u);
      // a number survives json.Marshal's float64 round trip only up to 2^53: clients send the rest as strings
      const uint64_t mag = sg && (int64_t)v.u < 0 ? (uint64_t)(0 - (int64_t)v.u) : v.u;
      const bool quote = as_key || mag >= (1ull << 53) || r.below(4) == 0;
      if (quote) { o.push_back('"'); o += buf; o.push_back('"'); } else o += buf;
      break;
    }
    case T_BOOL:
      if (as_key) o += v.u ? "\"true\"" : "\"false\"";
      else o += v.u ? "true" : "false";
      break;
    case T_ENUM: {
      const auto& e = P.enums[f.child];
      const EnumV* hit = nullptr;
      for (auto& x : e) if ((int64_t)x.number == (int64_t)(int32_t)(uint32_t)v.u) { hit = &x; break; }
      if (hit && r.below(5) != 0) { o.push_back('"'); o += hit->name; o.push_back('"'); }
      else { snprintf(buf, sizeof buf, "%d", (int32_t)(uint32_t)v.u); o += buf; }
      break;
    }
    case T_FLOAT: case T_DOUBLE: json_eighth(o, v.d); break;
    case T_STRING: json_string(o, v.s); break;
    case T_BYTES: o.push_back('"'); base64(o, v.s); o.push_back('"'); break;
  }
}
void civil_from_days(int64_t z, int& y, unsigned& m, unsigned& d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const unsigned doe = (unsigned)(z - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  y = (int)(yoe + era * 400);
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const unsigned mp = (5 * doy + 2) / 153;
  d = doy - (153 * mp + 2) / 5 + 1;
  m = mp < 10 ? mp + 3 : mp - 9;
  y += m <= 2;
}
void json_timestamp(std::string& o, const Val& t) {
  int64_t s = (int64_t)t.u;
  uint32_t ns = (uint32_t)t.d;
  int y; unsigned mo, d;
  civil_from_days(s / 86400, y, mo, d);
  int64_t rem = s % 86400;
  char buf[64];
  snprintf(buf, sizeof buf, "\"%04d-%02u-%02uT%02d:%02d:%02d", y, mo, d, (int)(rem / 3600), (int)(rem / 60 % 60), (int)(rem % 60));
  o += buf;
  if (ns) {
    snprintf(buf, sizeof buf, ".%09u", ns);
    std::string f = buf;
    while (f.size() > 4 && f.substr(f.size() - 3) == "000") f.resize(f.size() - 3);
    o += f;
  }
  o += "Z\"";
}
void json_message(std::string& o, const Plan& P, const Val& m, Rng& r, bool json_names);
void json_value(std::string& o, const Plan& P, const Field& f, const Val& v, Rng& r, bool json_names) {
  if (f.flags & FL_TIMESTAMP) json_timestamp(o, v);
  else if (f.type == T_MESSAGE) json_message(o, P, v, r, json_names);
  else json_scalar(o, P, f, v, r, false);
}
void json_message(std::string& o, const Plan& P, const Val& m, Rng& r, bool json_names) {
  // encoding/json sorts the keys of a map bytewise
  std::vector<std::pair<std::string, size_t>> keys;
  for (size_t i = 0; i < m.fields.size(); i++) keys.push_back({json_names ? m.fields[i].first->json_name : m.fields[i].first->name, i});
  std::sort(keys.begin(), keys.end());
  o.push_back('{');
  bool first = true;
  for (auto& kx : keys) {
    const Field& f = *m.fields[kx.second].first;
    const std::vector<Val>& vals = m.fields[kx.second].second;
    if (!first) o.push_back(',');
    first = false;
    json_string(o, kx.first);
    o.push_back(':');
    if (f.flags & FL_MAP) {
      const Msg& E = P.msgs[f.child];
      std::vector<std::pair<std::string, size_t>> ek;
      for (size_t i = 0; i < vals.size(); i++) {
        std::string k;
        if (E.fields[0].type == T_STRING) k = vals[i].kv[0].s;
        else {
          std::string t;
          json_scalar(t, P, E.fields[0], vals[i].kv[0], r, true);
          k = t.substr(1, t.size() - 2);
        }
        ek.push_back({k, i});
      }
      std::sort(ek.begin(), ek.end());
      o.push_back('{');
      for (size_t i = 0; i < ek.size(); i++) {
        if (i) o.push_back(',');
        json_string(o, ek[i].first);
        o.push_back(':');
        json_value(o, P, E.fields[1], vals[ek[i].second].kv[1], r, json_names);
      }
      o.push_back('}');
    } else if (f.flags & FL_REPEATED) {
      o.push_back('[');
      for (size_t i = 0; i < vals.size(); i++) {
        if (i) o.push_back(',');
        json_value(o, P, f, vals[i], r, json_names);
      }
      o.push_back(']');
    } else {
      json_value(o, P, f, vals[0], r, json_names);
    }
  }
  o.push_back('}');
}

// ---- wire (reply side) ---------------------------------------------------------------------------
void put_varint(std::string& b, uint64_t v) {
  while (v >= 0x80) { b.push_back((char)(v | 0x80)); v >>= 7; }
  b.push_back((char)v);
}
uint32_t wire_type(uint32_t t) {
  switch (t) {
    case T_DOUBLE: case T_FIXED64: case T_SFIXED64: return 1;
    case T_FLOAT: case T_FIXED32: case T_SFIXED32: return 5;
    case T_STRING: case T_BYTES: case T_MESSAGE: return 2;
    default: return 0;
  }
}
void wire_message(std::string& b, const Plan& P, const Val& m);
void wire_scalar_payload(std::string& b, const Field& f, const Val& v) {
  switch (f.type) {
    case T_INT32: case T_INT64: case T_UINT32: case T_UINT64: case T_BOOL: case T_ENUM: put_varint(b, v.u); break;
    case T_SINT32: { uint32_t x = (uint32_t)v.u; put_varint(b, (uint32_t)((x << 1) ^ (uint32_t)((int32_t)x >> 31))); break; }
    case T_SINT64: put_varint(b, (v.u << 1) ^ (uint64_t)((int64_t)v.u >> 63)); break;
    case T_FIXED32: case T_SFIXED32: { uint32_t x = (uint32_t)v.u; b.append((const char*)&x, 4); break; }
    case T_FIXED64: case T_SFIXED64: { uint64_t x = v.u; b.append((const char*)&x, 8); break; }
    case T_FLOAT: { float x = (float)v.d; b.append((const char*)&x, 4); break; }
    case T_DOUBLE: { double x = v.d; b.append((const char*)&x, 8); break; }
    case T_STRING: case T_BYTES: put_varint(b, v.s.size()); b += v.s; break;
  }
}
bool is_zero(const Field& f, const Val& v) {
  switch (f.type) {
    case T_STRING: case T_BYTES: return v.s.empty();
    case T_FLOAT: case T_DOUBLE: return v.d == 0 && !std::signbit(v.d);
    default: return v.u == 0;
  }
}
void wire_value(std::string& b, const Plan& P, const Field& f, const Val& v) {
  if (f.flags & FL_TIMESTAMP) {
    std::string t;
    if (v.u) { put_varint(t, 1 << 3); put_varint(t, v.u); }
    if ((uint32_t)v.d) { put_varint(t, 2 << 3); put_varint(t, (uint32_t)v.d); }
    put_varint(b, t.size());
    b += t;
  } else if (f.type == T_MESSAGE) {
    std::string sub;
    wire_message(sub, P, v);
    put_varint(b, sub.size());
    b += sub;
  } else {
    wire_scalar_payload(b, f, v);
  }
}
void wire_message(std::string& b, const Plan& P, const Val& m) {
  std::vector<size_t> order(m.fields.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t c) { return m.fields[a].first->number < m.fields[c].first->number; });
  for (size_t oi : order) {
    const Field& f = *m.fields[oi].first;
    const std::vector<Val>& vals = m.fields[oi].second;
    if (f.flags & FL_MAP) {
      const Msg& E = P.msgs[f.child];
      for (auto& e : vals) {
        std::string ent;
        // generated code writes key and value of an entry unconditionally
        put_varint(ent, (1 << 3) | wire_type(E.fields[0].type));
        wire_value(ent, P, E.fields[0], e.kv[0]);
        put_varint(ent, (2 << 3) | wire_type(E.fields[1].type));
        wire_value(ent, P, E.fields[1], e.kv[1]);
        put_varint(b, ((uint64_t)f.number << 3) | 2);
        put_varint(b, ent.size());
        b += ent;
      }
    } else if ((f.flags & FL_REPEATED) && (f.flags & FL_PACKED)) {
      if (vals.empty()) continue;
      std::string run;
      for (auto& v : vals) wire_scalar_payload(run, f, v);
      put_varint(b, ((uint64_t)f.number << 3) | 2);
      put_varint(b, run.size());
      b += run;
    } else if (f.flags & FL_REPEATED) {
      for (auto& v : vals) {
        put_varint(b, ((uint64_t)f.number << 3) | wire_type(f.type));
        wire_value(b, P, f, v);
      }
    } else {
      const Val& v = vals[0];
      const bool presence = (f.flags & (FL_PRESENCE | FL_ONEOF)) || f.type == T_MESSAGE;
      if (!presence && is_zero(f, v)) continue;
      put_varint(b, ((uint64_t)f.number << 3) | wire_type(f.type));
      wire_value(b, P, f, v);
    }
  }
}

uint32_t zipf_pick(Rng& r, const std::vector<double>& cdf) {
  const double u = r.unit() * cdf.back();
  return (uint32_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
}

}  // namespace

extern "C" {

// configs[4]: method[i] = index into the plan's method list.  Returns 0, or -1 when a buffer is too small.
int ggr_gen_mixed(const uint8_t* plan, uint64_t plan_bytes, uint64_t seed, int64_t n, uint8_t* json, uint64_t json_cap, uint64_t* json_off,
                  uint8_t* wire, uint64_t wire_cap, uint64_t* wire_off, int32_t* method) {
  const Plan P = parse_plan(plan, plan_bytes);
  const uint32_t nm = (uint32_t)P.methods.size();
  if (!nm) return -2;
  std::vector<double> mcdf(nm), scdf(10);
  double acc = 0;
  for (uint32_t k = 0; k < nm; k++) { acc += 1.0 / std::pow((double)(k + 1), 1.1); mcdf[k] = acc; }
  acc = 0;
  for (uint32_t k = 0; k < 10; k++) { acc += 1.0 / std::pow((double)(k + 1), 1.2); scdf[k] = acc; }  // 64 B ... 64 KiB
  // popularity rank -> method: a fixed shuffle so that the popular methods are not the first ones declared
  std::vector<uint32_t> rank(nm);
  for (uint32_t k = 0; k < nm; k++) rank[k] = k;
  {
    Rng rs(0xB2000005ull);
    for (uint32_t k = nm - 1; k > 0; k--) std::swap(rank[k], rank[rs.below(k + 1)]);
  }
  // items are independent (seeded per index): generated by a few threads into local buffers, then laid out
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 32) nt = 32;
  if (nt < 1) nt = 1;
  if ((int64_t)nt > n) nt = n > 0 ? (unsigned)n : 1;
  struct Part { std::string js, wb; std::vector<uint32_t> jl, wl; };
  std::vector<Part> parts(nt);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) {
    th.emplace_back([&, t]() {
      Part& pt = parts[t];
      const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
      for (int64_t i = lo; i < hi; i++) {
        Rng r(seed + 0x100000001B3ull * (uint64_t)i);
        const uint32_t mi = rank[zipf_pick(r, mcdf)];
        method[i] = (int32_t)mi;
        const uint32_t bucket = zipf_pick(r, scdf);               // 64 * 2^bucket ... 64 * 2^(bucket + 1)
        const uint32_t target = (64u << bucket) + r.below(64u << bucket);
        Gen g(P, r);
        const Val req = g.message(P.methods[mi].first, target, 0);
        const Val rep = g.message(P.methods[mi].second, target * 3 / 5, 0);
        const size_t j0 = pt.js.size(), w0 = pt.wb.size();
        json_message(pt.js, P, req, r, r.below(2) == 0);
        wire_message(pt.wb, P, rep);
        pt.jl.push_back((uint32_t)(pt.js.size() - j0));
        pt.wl.push_back((uint32_t)(pt.wb.size() - w0));
      }
    });
  }
  for (auto& x : th) x.join();
  uint64_t jn = 0, wn = 0;
  bool ok = true;
  int64_t i = 0;
  for (unsigned t = 0; t < nt; t++) {
    const Part& pt = parts[t];
    if (jn + pt.js.size() > json_cap || wn + pt.wb.size() > wire_cap) ok = false;
    else {
      memcpy(json + jn, pt.js.data(), pt.js.size());
      memcpy(wire + wn, pt.wb.data(), pt.wb.size());
    }
    uint64_t ja = jn, wa = wn;
    for (size_t k = 0; k < pt.jl.size(); k++, i++) {
      json_off[i] = ja;
      wire_off[i] = wa;
      ja += pt.jl[k];
      wa += pt.wl[k];
    }
    jn += pt.js.size();
    wn += pt.wb.size();
  }
  json_off[n] = jn;
  wire_off[n] = wn;
  return ok ? 0 : -1;
}

}  // extern "C"
