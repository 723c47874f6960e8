// orc_util.h - byte-level helpers of the CPU oracle (TEST INFRASTRUCTURE, see ggr_oracle.h).
//
// Restates small pieces of Go's standard library that the reference path depends on
// [upstream, Go 1.23; not vendored under /root/reference]:
//   unicode/utf8 DecodeRune         -> utf8_decode
//   encoding/base64 (Std/URL, +-pad) -> b64_encode / b64_decode_go
//   strconv.AppendFloat(-1 precision, 'e'/'f') as used by encoding/json floatEncoder and
//   protojson's internal/encoding/json appendFloat -> format_float_go
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace orc {

typedef std::string Bytes;  // arbitrary bytes

// ---------- varint ----------
inline void put_varint(Bytes& b, uint64_t v) {
  while (v >= 0x80) {
    b.push_back((char)(v | 0x80));
    v >>= 7;
  }
  b.push_back((char)v);
}
inline int varint_size(uint64_t v) {
  int n = 1;
  while (v >= 0x80) {
    v >>= 7;
    n++;
  }
  return n;
}
// protowire.ConsumeVarint: up to 10 bytes, the 10th must be <= 1.
inline bool get_varint(const uint8_t*& p, const uint8_t* e, uint64_t& v) {
  v = 0;
  for (int i = 0; i < 10; i++) {
    if (p >= e) return false;
    uint8_t c = *p++;
    if (i == 9 && c > 1) return false;  // overflow
    v |= (uint64_t)(c & 0x7f) << (7 * i);
    if (c < 0x80) return true;
  }
  return false;
}
inline uint64_t zigzag64(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
inline int64_t unzigzag64(uint64_t v) { return (int64_t)(v >> 1) ^ -(int64_t)(v & 1); }

// ---------- UTF-8 (Go unicode/utf8 semantics) ----------
// returns rune and sets n; invalid -> 0xFFFD with n == 1 (n == 0 when empty)
inline uint32_t utf8_decode(const uint8_t* p, size_t len, int& n) {
  if (len == 0) {
    n = 0;
    return 0xFFFD;
  }
  uint8_t c = p[0];
  if (c < 0x80) {
    n = 1;
    return c;
  }
  n = 1;
  if (c < 0xC2 || c > 0xF4) return 0xFFFD;
  if (c < 0xE0) {
    if (len < 2 || (p[1] & 0xC0) != 0x80) return 0xFFFD;
    n = 2;
    return ((c & 0x1F) << 6) | (p[1] & 0x3F);
  }
  if (c < 0xF0) {
    uint8_t lo = 0x80, hi = 0xBF;
    if (c == 0xE0) lo = 0xA0;
    if (c == 0xED) hi = 0x9F;
    if (len < 2 || p[1] < lo || p[1] > hi) return 0xFFFD;
    if (len < 3 || (p[2] & 0xC0) != 0x80) return 0xFFFD;
    n = 3;
    return ((c & 0x0F) << 12) | ((p[1] & 0x3F) << 6) | (p[2] & 0x3F);
  }
  uint8_t lo = 0x80, hi = 0xBF;
  if (c == 0xF0) lo = 0x90;
  if (c == 0xF4) hi = 0x8F;
  if (len < 2 || p[1] < lo || p[1] > hi) return 0xFFFD;
  if (len < 3 || (p[2] & 0xC0) != 0x80) return 0xFFFD;
  if (len < 4 || (p[3] & 0xC0) != 0x80) return 0xFFFD;
  n = 4;
  return ((c & 0x07) << 18) | ((p[1] & 0x3F) << 12) | ((p[2] & 0x3F) << 6) | (p[3] & 0x3F);
}
inline bool utf8_valid(const uint8_t* p, size_t len) {
  size_t i = 0;
  while (i < len) {
    if (p[i] < 0x80) {
      i++;
      continue;
    }
    int n;
    uint32_t r = utf8_decode(p + i, len - i, n);
    if (r == 0xFFFD && n == 1) return false;
    i += n;
  }
  return true;
}
inline void utf8_append(Bytes& out, uint32_t r) {
  if (r < 0x80) {
    out.push_back((char)r);
  } else if (r < 0x800) {
    out.push_back((char)(0xC0 | (r >> 6)));
    out.push_back((char)(0x80 | (r & 0x3F)));
  } else if (r < 0x10000) {
    out.push_back((char)(0xE0 | (r >> 12)));
    out.push_back((char)(0x80 | ((r >> 6) & 0x3F)));
    out.push_back((char)(0x80 | (r & 0x3F)));
  } else {
    out.push_back((char)(0xF0 | (r >> 18)));
    out.push_back((char)(0x80 | ((r >> 12) & 0x3F)));
    out.push_back((char)(0x80 | ((r >> 6) & 0x3F)));
    out.push_back((char)(0x80 | (r & 0x3F)));
  }
}

// ---------- base64 ----------
inline void b64_encode(Bytes& out, const uint8_t* p, size_t n) {
  static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  size_t i = 0;
  for (; i + 3 <= n; i += 3) {
    uint32_t v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2];
    out.push_back(T[v >> 18]);
    out.push_back(T[(v >> 12) & 63]);
    out.push_back(T[(v >> 6) & 63]);
    out.push_back(T[v & 63]);
  }
  if (n - i == 1) {
    uint32_t v = p[i] << 16;
    out.push_back(T[v >> 18]);
    out.push_back(T[(v >> 12) & 63]);
    out.push_back('=');
    out.push_back('=');
  } else if (n - i == 2) {
    uint32_t v = (p[i] << 16) | (p[i + 1] << 8);
    out.push_back(T[v >> 18]);
    out.push_back(T[(v >> 12) & 63]);
    out.push_back(T[(v >> 6) & 63]);
    out.push_back('=');
  }
}
// Go encoding/base64 (*Encoding).DecodeString, non-strict, as protojson unmarshalBytes calls it
// [upstream encoding/protojson/decode.go unmarshalBytes]: alphabet URL iff the text contains '-'
// or '_'; padding NoPadding iff len%4 != 0.  '\r' and '\n' are skipped by the decoder.
inline bool b64_decode_go(const uint8_t* s, size_t n, Bytes& out) {
  bool url = false;
  for (size_t i = 0; i < n; i++)
    if (s[i] == '-' || s[i] == '_') url = true;
  bool padded = (n % 4) == 0;
  auto val = [&](uint8_t c) -> int {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (!url && c == '+') return 62;
    if (!url && c == '/') return 63;
    if (url && c == '-') return 62;
    if (url && c == '_') return 63;
    return -1;
  };
  size_t i = 0;
  bool end = false;
  while (!end) {
    // decode one quantum (mirrors decodeQuantum)
    int dbuf[4];
    int dlen = 4;
    int j = 0;
    for (; j < 4; j++) {
      if (i == n) {
        if (j == 0) {
          return true;  // clean end
        }
        if (j == 1 || padded) return false;
        dlen = j;
        end = true;
        break;
      }
      uint8_t c = s[i++];
      int v = val(c);
      if (v >= 0) {
        dbuf[j] = v;
        continue;
      }
      if (c == '\n' || c == '\r') {
        j--;
        continue;
      }
      if (c != '=' || !padded) return false;  // NoPadding: '=' is not in the alphabet
      // padding
      if (j == 0 || j == 1) return false;
      if (j == 2) {
        // expect a second '=' (skipping newlines)
        while (i < n && (s[i] == '\n' || s[i] == '\r')) i++;
        if (i == n) return false;
        if (s[i] != '=') return false;
        i++;
      }
      // skip over newlines; anything else after padding is an error
      while (i < n && (s[i] == '\n' || s[i] == '\r')) i++;
      if (i < n) return false;
      dlen = j;
      end = true;
      break;
    }
    uint32_t v = 0;
    for (int k = 0; k < dlen; k++) v |= (uint32_t)dbuf[k] << (18 - 6 * k);
    if (dlen == 4) {
      out.push_back((char)(v >> 16));
      out.push_back((char)(v >> 8));
      out.push_back((char)v);
    } else if (dlen == 3) {
      out.push_back((char)(v >> 16));
      out.push_back((char)(v >> 8));
    } else if (dlen == 2) {
      out.push_back((char)(v >> 16));
    }
  }
  return true;
}

// ---------- shortest float formatting (strconv.FormatFloat(f, fmt, -1, bits)) ----------
// Produces the shortest digit string that round-trips, closest to the true value among the
// shortest.  Implemented by search over precisions using glibc's correctly rounded
// printf/strtod; neighbours are tried so asymmetric (power-of-two) intervals are honoured.
inline bool roundtrips(const char* s, double v, int bits) {
  if (bits == 32) return strtof(s, nullptr) == (float)v;
  return strtod(s, nullptr) == v;
}
// digits: decimal digits without trailing zeros; dexp: value = 0.d1d2.. * 10^dexp
inline void shortest_digits(double v, int bits, std::string& digits, int& dexp) {
  // v finite, > 0
  char buf[64];
  int maxp = bits == 32 ? 9 : 17;
  for (int p = 1; p <= maxp; p++) {
    snprintf(buf, sizeof buf, "%.*e", p - 1, v);
    bool ok = roundtrips(buf, v, bits);
    std::string cand(buf);
    if (!ok && p < maxp) {
      // try the p-digit neighbours of the correctly rounded candidate
      // parse mantissa digits and exponent
      std::string m;
      int e10 = 0;
      {
        const char* q = buf;
        for (; *q && *q != 'e'; q++)
          if (*q >= '0' && *q <= '9') m.push_back(*q);
        e10 = atoi(q + 1);
      }
      double best = -1;
      for (int delta = -1; delta <= 1; delta += 2) {
        std::string mm = m;
        int ee = e10;
        // add delta in the last place
        int k = (int)mm.size() - 1;
        if (delta > 0) {
          while (k >= 0 && mm[k] == '9') mm[k--] = '0';
          if (k < 0) {
            mm.insert(mm.begin(), '1');
            mm.pop_back();
            ee++;
          } else
            mm[k]++;
        } else {
          while (k >= 0 && mm[k] == '0') mm[k--] = '9';
          if (k < 0) continue;
          mm[k]--;
          if (mm[0] == '0') continue;  // would lose a digit; not a p-digit number
        }
        std::string s2;
        s2.push_back(mm[0]);
        if (mm.size() > 1) {
          s2.push_back('.');
          s2.append(mm, 1, std::string::npos);
        }
        char eb[16];
        snprintf(eb, sizeof eb, "e%+03d", ee);
        s2 += eb;
        if (roundtrips(s2.c_str(), v, bits)) {
          long double d = fabsl(strtold(s2.c_str(), nullptr) - (long double)v);
          if (best < 0 || d < best) {
            best = (double)d;
            cand = s2;
            ok = true;
          }
        }
      }
    }
    if (ok) {
      const char* q = cand.c_str();
      digits.clear();
      for (; *q && *q != 'e'; q++)
        if (*q >= '0' && *q <= '9') digits.push_back(*q);
      int e10 = atoi(q + 1);
      while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
      dexp = e10 + 1;
      return;
    }
  }
  // unreachable: maxp digits always round-trip
  digits = "0";
  dexp = 1;
}
// Formatting rule shared by encoding/json floatEncoder.encode and protojson appendFloat
// [upstream encoding/json/encode.go; protobuf-go internal/encoding/json/encode.go]:
// 'e' iff abs < 1e-6 || abs >= 1e21 (compared in float32 for 32-bit values); "e-07" -> "e-7".
// NaN / Inf are the caller's business.
inline std::string format_float_go(double v, int bits) {
  std::string out;
  if (v == 0) {
    return std::signbit(v) ? "-0" : "0";
  }
  double a = fabs(v);
  bool efmt;
  if (bits == 32) {
    float fa = (float)a;
    efmt = fa < 1e-6f || fa >= 1e21f;
  } else {
    efmt = a < 1e-6 || a >= 1e21;
  }
  std::string d;
  int x;
  shortest_digits(a, bits, d, x);
  if (v < 0) out.push_back('-');
  int n = (int)d.size();
  if (efmt) {
    out.push_back(d[0]);
    if (n > 1) {
      out.push_back('.');
      out.append(d, 1, std::string::npos);
    }
    int e = x - 1;
    out.push_back('e');
    if (e < 0) {
      out.push_back('-');
      e = -e;
      // Go prints at least two exponent digits; the cleanup turns e-0X into e-X
      out += std::to_string(e);
    } else {
      out.push_back('+');
      if (e < 10) out.push_back('0');
      out += std::to_string(e);
    }
  } else {
    if (x <= 0) {
      out += "0.";
      out.append((size_t)(-x), '0');
      out += d;
    } else if (n <= x) {
      out += d;
      out.append((size_t)(x - n), '0');
    } else {
      out.append(d, 0, (size_t)x);
      out.push_back('.');
      out.append(d, (size_t)x, std::string::npos);
    }
  }
  return out;
}

inline void set_err(char* err, size_t cap, const std::string& s) {
  if (!err || cap == 0) return;
  size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
  memcpy(err, s.data(), n);
  err[n] = 0;
}

}  // namespace orc
