// orc_gojson.h - restatement of Go 1.23 encoding/json as the reference uses it
// (TEST INFRASTRUCTURE, see ggr_oracle.h).
//
//   decode into interface{} : /root/reference/pkg/server/handler.go:83-84 (json.NewDecoder.Decode),
//                             /root/reference/pkg/mcp/types.go:19-33 (RequestID.UnmarshalJSON)
//   marshal interface{}     : /root/reference/pkg/server/handler.go:224-231 (json.Marshal(args))
//   string escaping         : /root/reference/pkg/server/handler.go:290-297 (Encoder.Encode, HTML-safe)
// [upstream encoding/json/{scanner,decode,encode,tables}.go]
#pragma once
#include <map>
#include <string>
#include <vector>

#include "orc_util.h"

namespace orc {

struct JVal {
  enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
  bool b = false;
  double n = 0;
  Bytes s;
  std::vector<JVal> a;
  std::vector<std::pair<Bytes, JVal>> mem;  // object members in document order, duplicates kept
  // Go map view: duplicate keys -> last wins; json.Marshal sorts keys bytewise
  std::map<Bytes, const JVal*> as_map() const {
    std::map<Bytes, const JVal*> m;
    for (auto& kv : mem) m[kv.first] = &kv.second;
    return m;
  }
  const JVal* get(const Bytes& k) const {
    const JVal* r = nullptr;
    for (auto& kv : mem)
      if (kv.first == k) r = &kv.second;
    return r;
  }
};

struct GoJsonParser {
  const uint8_t* p;
  const uint8_t* e;
  int depth = 0;
  GoJsonParser(const uint8_t* b, size_t n) : p(b), e(b + n) {}
  void ws() {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
  }
  static int hexv(uint8_t c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
  }
  int getu4(const uint8_t* q) {  // q points at "\uXXXX"
    if (e - q < 6 || q[0] != '\\' || q[1] != 'u') return -1;
    int r = 0;
    for (int i = 2; i < 6; i++) {
      int h = hexv(q[i]);
      if (h < 0) return -1;
      r = r * 16 + h;
    }
    return r;
  }
  // scanner + unquote: control chars rejected, invalid UTF-8 -> U+FFFD, lone surrogates -> U+FFFD
  bool string(Bytes& out) {
    if (p >= e || *p != '"') return false;
    p++;
    while (true) {
      if (p >= e) return false;
      uint8_t c = *p;
      if (c == '"') {
        p++;
        return true;
      }
      if (c < 0x20) return false;
      if (c == '\\') {
        if (p + 1 >= e) return false;
        uint8_t d = p[1];
        switch (d) {
          case '"': out.push_back('"'); p += 2; break;
          case '\\': out.push_back('\\'); p += 2; break;
          case '/': out.push_back('/'); p += 2; break;
          case 'b': out.push_back('\b'); p += 2; break;
          case 'f': out.push_back('\f'); p += 2; break;
          case 'n': out.push_back('\n'); p += 2; break;
          case 'r': out.push_back('\r'); p += 2; break;
          case 't': out.push_back('\t'); p += 2; break;
          case 'u': {
            int rr = getu4(p);
            if (rr < 0) return false;
            p += 6;
            if (rr >= 0xD800 && rr < 0xE000) {
              int rr1 = getu4(p);
              if (rr >= 0xD800 && rr < 0xDC00 && rr1 >= 0xDC00 && rr1 < 0xE000) {
                uint32_t dec = 0x10000 + (((uint32_t)rr - 0xD800) << 10) + ((uint32_t)rr1 - 0xDC00);
                p += 6;
                utf8_append(out, dec);
                break;
              }
              rr = 0xFFFD;
            }
            utf8_append(out, (uint32_t)rr);
            break;
          }
          default: return false;
        }
        continue;
      }
      if (c < 0x80) {
        out.push_back((char)c);
        p++;
        continue;
      }
      int n;
      uint32_t r = utf8_decode(p, (size_t)(e - p), n);
      if (r == 0xFFFD && n == 1) {
        utf8_append(out, 0xFFFD);
        p += 1;
      } else {
        out.append((const char*)p, (size_t)n);
        p += n;
      }
    }
  }
  bool number(double& v) {
    const uint8_t* s = p;
    if (p < e && *p == '-') p++;
    if (p >= e) return false;
    if (*p == '0') p++;
    else if (*p >= '1' && *p <= '9') {
      while (p < e && *p >= '0' && *p <= '9') p++;
    } else return false;
    if (p < e && *p == '.') {
      p++;
      if (p >= e || *p < '0' || *p > '9') return false;
      while (p < e && *p >= '0' && *p <= '9') p++;
    }
    if (p < e && (*p == 'e' || *p == 'E')) {
      p++;
      if (p < e && (*p == '+' || *p == '-')) p++;
      if (p >= e || *p < '0' || *p > '9') return false;
      while (p < e && *p >= '0' && *p <= '9') p++;
    }
    std::string t((const char*)s, (size_t)(p - s));
    v = strtod(t.c_str(), nullptr);
    if (std::isinf(v)) return false;  // strconv.ParseFloat ErrRange -> UnmarshalTypeError
    return true;
  }
  bool lit(const char* w) {
    size_t n = strlen(w);
    if ((size_t)(e - p) < n || memcmp(p, w, n) != 0) return false;
    p += n;
    return true;
  }
  bool value(JVal& v) {
    ws();
    if (p >= e) return false;
    if (++depth > 10000) return false;  // scanner maxNestingDepth
    bool ok = value1(v);
    depth--;
    return ok;
  }
  bool value1(JVal& v) {
    switch (*p) {
      case '{': {
        p++;
        v.t = JVal::Obj;
        ws();
        if (p < e && *p == '}') {
          p++;
          return true;
        }
        while (true) {
          ws();
          Bytes k;
          if (!string(k)) return false;
          ws();
          if (p >= e || *p != ':') return false;
          p++;
          JVal c;
          if (!value(c)) return false;
          v.mem.emplace_back(std::move(k), std::move(c));
          ws();
          if (p < e && *p == ',') {
            p++;
            continue;
          }
          if (p < e && *p == '}') {
            p++;
            return true;
          }
          return false;
        }
      }
      case '[': {
        p++;
        v.t = JVal::Arr;
        ws();
        if (p < e && *p == ']') {
          p++;
          return true;
        }
        while (true) {
          JVal c;
          if (!value(c)) return false;
          v.a.push_back(std::move(c));
          ws();
          if (p < e && *p == ',') {
            p++;
            continue;
          }
          if (p < e && *p == ']') {
            p++;
            return true;
          }
          return false;
        }
      }
      case '"': v.t = JVal::Str; return string(v.s);
      case 't': v.t = JVal::Bool; v.b = true; return lit("true");
      case 'f': v.t = JVal::Bool; v.b = false; return lit("false");
      case 'n': v.t = JVal::Null; return lit("null");
      default: v.t = JVal::Num; return number(v.n);
    }
  }
};

// encoding/json appendString with escapeHTML = true (json.Marshal and Encoder default)
inline void go_json_string(Bytes& out, const uint8_t* s, size_t n) {
  static const char hex[] = "0123456789abcdef";
  out.push_back('"');
  size_t i = 0;
  while (i < n) {
    uint8_t b = s[i];
    if (b < 0x80) {
      if (b >= 0x20 && b != '"' && b != '\\' && b != '<' && b != '>' && b != '&') {
        out.push_back((char)b);
        i++;
        continue;
      }
      out.push_back('\\');
      switch (b) {
        case '\\': case '"': out.push_back((char)b); break;
        case '\b': out.push_back('b'); break;
        case '\f': out.push_back('f'); break;
        case '\n': out.push_back('n'); break;
        case '\r': out.push_back('r'); break;
        case '\t': out.push_back('t'); break;
        default:
          out += "u00";
          out.push_back(hex[b >> 4]);
          out.push_back(hex[b & 0xF]);
      }
      i++;
      continue;
    }
    int w;
    uint32_t c = utf8_decode(s + i, n - i, w);
    if (c == 0xFFFD && w == 1) {
      out += "\\ufffd";
      i += 1;
      continue;
    }
    if (c == 0x2028 || c == 0x2029) {
      out += "\\u202";
      out.push_back(hex[c & 0xF]);
      i += w;
      continue;
    }
    out.append((const char*)s + i, (size_t)w);
    i += w;
  }
  out.push_back('"');
}

// json.Marshal of the interface{} tree: compact, map keys sorted, floats in ES6 style.
// Returns false for NaN/Inf (UnsupportedValueError) - cannot occur for values that came from
// decoding JSON text.
inline bool go_json_marshal(Bytes& out, const JVal& v) {
  switch (v.t) {
    case JVal::Null: out += "null"; return true;
    case JVal::Bool: out += v.b ? "true" : "false"; return true;
    case JVal::Num:
      if (std::isnan(v.n) || std::isinf(v.n)) return false;
      out += format_float_go(v.n, 64);
      return true;
    case JVal::Str: go_json_string(out, (const uint8_t*)v.s.data(), v.s.size()); return true;
    case JVal::Arr:
      out.push_back('[');
      for (size_t i = 0; i < v.a.size(); i++) {
        if (i) out.push_back(',');
        if (!go_json_marshal(out, v.a[i])) return false;
      }
      out.push_back(']');
      return true;
    case JVal::Obj: {
      out.push_back('{');
      bool first = true;
      for (auto& kv : v.as_map()) {
        if (!first) out.push_back(',');
        first = false;
        go_json_string(out, (const uint8_t*)kv.first.data(), kv.first.size());
        out.push_back(':');
        if (!go_json_marshal(out, *kv.second)) return false;
      }
      out.push_back('}');
      return true;
    }
  }
  return false;
}

}  // namespace orc
