// ggr_wrap.cuh - MCP result wrapping on the device.
//
// handler.go:265-270 builds ToolCallResult{Content: [TextContent(protojson text)]} and
// handler.go:290-297 writes JSONRPCResponse through json.NewEncoder: the body is
//   {"jsonrpc":"2.0","result":{"content":[{"type":"text","text":"<T escaped>"}]},"id":<id>}\n
// with T escaped by encoding/json's appendString (escapeHTML = true): " and \ get a backslash,
// < > & become \u003c \u003e \u0026, control characters \b \f \n \r \t or \u00XX, U+2028 / U+2029
// \u2028 / \u2029; T is valid UTF-8 here (protojson produced it).  One warp per item: lane k
// owns 8 consecutive bytes of T per round; sizes first, the batch-wide scan, then the write.
#pragma once
#include "ggr_warp.cuh"

#define GGR_WRAP_P1 "{\"jsonrpc\":\"2.0\",\"result\":{\"content\":[{\"type\":\"text\",\"text\":\""
#define GGR_WRAP_P2 "\"}]},\"id\":"
#define GGR_WRAP_P3 "}\n"
#define GGR_WRAP_P1_LEN 61u
#define GGR_WRAP_P2_LEN 10u
#define GGR_WRAP_P3_LEN 2u

// bytes byte i of t[0, n) contributes to the escaped text
GGR_DEV u32 wrap_len_at(const u8* t, u32 i, u32 n) {
  const u32 c = t[i];
  if (c < 0x80u) {
    if (c >= 0x20u) return (c == '"' || c == '\\') ? 2u : (c == '<' || c == '>' || c == '&') ? 6u : 1u;
    return (c == 8u || c == 9u || c == 10u || c == 12u || c == 13u) ? 2u : 6u;
  }
  // U+2028 / U+2029 = E2 80 A8 / E2 80 A9: six characters at the lead byte, none at the other two
  if (c == 0xE2u) return (i + 2u < n && t[i + 1] == 0x80u && (t[i + 2] == 0xA8u || t[i + 2] == 0xA9u)) ? 6u : 1u;
  if (c == 0x80u) return (i >= 1u && i + 1u < n && t[i - 1] == 0xE2u && (t[i + 1] == 0xA8u || t[i + 1] == 0xA9u)) ? 0u : 1u;
  if (c == 0xA8u || c == 0xA9u) return (i >= 2u && t[i - 1] == 0x80u && t[i - 2] == 0xE2u) ? 0u : 1u;
  return 1u;
}
// writes the escaped form of byte i at d, returns its length (wrap_len_at)
GGR_DEV u32 wrap_put_at(const u8* t, u32 i, u32 n, u8* d) {
  const u32 l = wrap_len_at(t, i, n);
  const u32 c = t[i];
  if (l == 1u) {
    d[0] = (u8)c;
  } else if (l == 2u) {
    d[0] = '\\';
    d[1] = (u8)(c == 8u ? 'b' : c == 12u ? 'f' : c == 10u ? 'n' : c == 13u ? 'r' : c == 9u ? 't' : c);
  } else if (l == 6u) {
    u32 hi, lo;
    d[0] = '\\';
    d[1] = 'u';
    if (c == 0xE2u) {  // \u2028 or \u2029
      d[2] = '2'; d[3] = '0'; d[4] = '2';
      d[5] = (u8)(t[i + 2] == 0xA8u ? '8' : '9');
      return l;
    }
    hi = c >> 4;
    lo = c & 15u;
    d[2] = '0'; d[3] = '0';
    d[4] = (u8)('0' + hi);
    d[5] = (u8)(lo < 10u ? '0' + lo : 'a' + lo - 10u);
  }
  return l;
}

// all lanes: size of the whole body for text t[0, n) and an id token of idn bytes
GGR_DEV u32 wrap_size_item(const u8* t, u32 n, u32 idn) {
  const u32 lane = wp_lane();
  u32 sum = 0;
  for (u32 i = lane; i < n; i += 32) sum += wrap_len_at(t, i, n);
  for (u32 d = 16; d >= 1; d >>= 1) sum += WP_SHFL(sum, lane ^ d);
  return GGR_WRAP_P1_LEN + sum + GGR_WRAP_P2_LEN + idn + GGR_WRAP_P3_LEN;
}

// all lanes: the body at out[0 ..)
GGR_DEV void wrap_write_item(const u8* t, u32 n, const u8* id, u32 idn, u8* out) {
  const u32 lane = wp_lane();
  const char* p1 = GGR_WRAP_P1;
  for (u32 i = lane; i < GGR_WRAP_P1_LEN; i += 32) out[i] = (u8)p1[i];
  u32 base = GGR_WRAP_P1_LEN;
  for (u32 r = 0; r < n; r += 256u) {  // lane k: bytes [r + 8k, r + 8k + 8)
    const u32 b0 = r + lane * 8u;
    u32 mine = 0;
    for (u32 j = 0; j < 8u; j++)
      if (b0 + j < n) mine += wrap_len_at(t, b0 + j, n);
    u32 tot;
    u32 o = base + WP_EXCL_SCAN(mine, &tot);
    for (u32 j = 0; j < 8u; j++)
      if (b0 + j < n) o += wrap_put_at(t, b0 + j, n, out + o);
    base += tot;
  }
  const char* p2 = GGR_WRAP_P2;
  if (lane < GGR_WRAP_P2_LEN) out[base + lane] = (u8)p2[lane];
  base += GGR_WRAP_P2_LEN;
  for (u32 i = lane; i < idn; i += 32) out[base + i] = id[i];
  base += idn;
  if (lane == 0) {
    out[base] = '}';
    out[base + 1] = '\n';
  }
}
