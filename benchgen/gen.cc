// gen.cc - seeded synthetic workload generator for the BASELINE.json configs.
//
// Produces, for each item, the request-side input (canonical `arguments` JSON exactly as
// encoding/json.Marshal would print it: compact, keys sorted, HTML-safe escapes) and the
// reply-side input (protobuf wire bytes as a generated-code backend would send them).
// Benchmark infrastructure: neither product nor oracle.  Deterministic (splitmix64).
//
//   config 2  bench.Flat  {int32 a1..a8; string s1..s4}                     ~256 B JSON
//   config 3  ProcessNodeRequest trees (4 levels, ~40 nodes) / 20% CreateDocumentRequest with a
//             24-entry map; replies: Node echo / GetUserProfileResponse      ~4 KB JSON
//   config 4  bench.Blob replies with 64 KiB of random bytes
#include <cstdint>
#include <cstring>
#include <string>
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

struct Out {
  uint8_t* p;
  uint64_t cap, n;
  bool ok;
  Out(uint8_t* b, uint64_t c) : p(b), cap(c), n(0), ok(true) {}
  void put(const void* s, size_t k) {
    if (n + k > cap) { ok = false; n += k; return; }
    memcpy(p + n, s, k);
    n += k;
  }
  void ch(char c) { put(&c, 1); }
  void str(const char* s) { put(s, strlen(s)); }
  void str(const std::string& s) { put(s.data(), s.size()); }
};

void put_varint(std::string& b, uint64_t v) {
  while (v >= 0x80) { b.push_back((char)(v | 0x80)); v >>= 7; }
  b.push_back((char)v);
}
void put_len_field(std::string& b, uint32_t num, const std::string& payload) {
  put_varint(b, (num << 3) | 2);
  put_varint(b, payload.size());
  b += payload;
}

// random text: 90% printable ASCII, 5% with characters that need escaping, 5% with 2-3 byte UTF-8
std::string rand_text(Rng& r, uint32_t lo, uint32_t hi) {
  static const char A[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 _-.,:;!?@#$%^*()[]{}+=~|'/";
  static const char* M[] = {"\xC3\xA9", "\xC3\xB6", "\xC3\xB1", "\xE5\xBC\xA0", "\xE4\xB8\x89", "\xE6\x97\xA5", "\xE2\x82\xAC", "\xC2\xA0"};
  uint32_t len = r.range(lo, hi);
  double kind = r.unit();
  std::string s;
  while (s.size() < len) {
    double u = r.unit();
    if (kind >= 0.90 && kind < 0.95 && u < 0.12) {
      static const char E[] = {'"', '\\', '\n'};
      s.push_back(E[r.below(3)]);
    } else if (kind >= 0.95 && u < 0.25) {
      s += M[r.below(8)];
    } else {
      s.push_back(A[r.below(sizeof(A) - 1)]);
    }
  }
  return s;
}
// encoding/json appendString, escapeHTML=true
void json_string(Out& o, const std::string& s) {
  static const char hex[] = "0123456789abcdef";
  o.ch('"');
  for (size_t i = 0; i < s.size(); i++) {
    unsigned char c = (unsigned char)s[i];
    if (c >= 0x80) {
      // generator only emits valid UTF-8; U+2028/9 never produced
      o.ch((char)c);
    } else if (c == '"' || c == '\\') { o.ch('\\'); o.ch((char)c); }
    else if (c == '\n') o.str("\\n");
    else if (c == '\r') o.str("\\r");
    else if (c == '\t') o.str("\\t");
    else if (c == '\b') o.str("\\b");
    else if (c == '\f') o.str("\\f");
    else if (c < 0x20 || c == '<' || c == '>' || c == '&') { o.str("\\u00"); o.ch(hex[c >> 4]); o.ch(hex[c & 15]); }
    else o.ch((char)c);
  }
  o.ch('"');
}

int32_t rand_i32(Rng& r) {
  double u = r.unit();
  if (u < 0.5) return (int32_t)r.below(128);
  if (u < 0.75) return (int32_t)(r.next() & 0x7FFFFFFF);
  return -(int32_t)(r.next() & 0x7FFFFFFF) - 1;
}

struct Node {
  std::string id, value;
  std::vector<Node> kids;
};
void build_tree(Rng& r, Node& n, int level) {
  n.id = rand_text(r, 6, 12);
  n.value = rand_text(r, 30, 60);
  if (level < 4) {
    uint32_t k = 2 + (r.below(8) >= 1) + (r.below(8) >= 6);  // mean 3.125 -> ~44 nodes, ~4 KB
    n.kids.resize(k);
    for (auto& c : n.kids) build_tree(r, c, level + 1);
  }
}
void node_json(Out& o, const Node& n) {  // keys sorted: children, id, value
  o.ch('{');
  if (!n.kids.empty()) {
    o.str("\"children\":[");
    for (size_t i = 0; i < n.kids.size(); i++) {
      if (i) o.ch(',');
      node_json(o, n.kids[i]);
    }
    o.str("],");
  }
  o.str("\"id\":");
  json_string(o, n.id);
  o.str(",\"value\":");
  json_string(o, n.value);
  o.ch('}');
}
void node_wire(std::string& b, const Node& n) {
  if (!n.id.empty()) put_len_field(b, 1, n.id);
  if (!n.value.empty()) put_len_field(b, 2, n.value);
  for (auto& c : n.kids) {
    std::string sub;
    node_wire(sub, c);
    put_len_field(b, 3, sub);
  }
}

}  // namespace

extern "C" {

// config 2: returns 0 on success, -1 when a buffer is too small
int ggr_gen_flat(uint64_t seed, int64_t n, uint8_t* json, uint64_t json_cap, uint64_t* json_off, uint8_t* wire,
                 uint64_t wire_cap, uint64_t* wire_off) {
  Out j(json, json_cap), w(wire, wire_cap);
  for (int64_t i = 0; i < n; i++) {
    Rng r(seed + 0x100000001B3ull * (uint64_t)i);
    json_off[i] = j.n;
    wire_off[i] = w.n;
    std::string wb;
    j.ch('{');
    bool first = true;
    for (int k = 1; k <= 8; k++) {
      int32_t v = rand_i32(r);
      if (v != 0) {
        put_varint(wb, (uint32_t)k << 3);
        put_varint(wb, (uint64_t)(int64_t)v);
      }
      // json.Marshal prints every key that is in the map, zero or not
      if (!first) j.ch(',');
      first = false;
      char buf[32];
      snprintf(buf, sizeof buf, "\"a%d\":%d", k, v);
      j.str(buf);
    }
    for (int k = 1; k <= 4; k++) {
      std::string s = rand_text(r, 8, 40);
      put_len_field(wb, 8 + k, s);
      char buf[16];
      snprintf(buf, sizeof buf, ",\"s%d\":", k);
      j.str(buf);
      json_string(j, s);
    }
    j.ch('}');
    w.put(wb.data(), wb.size());
  }
  json_off[n] = j.n;
  wire_off[n] = w.n;
  return (j.ok && w.ok) ? 0 : -1;
}

// config 3: kind[i] = 0 ProcessNodeRequest (reply: Node echo), 1 CreateDocumentRequest (reply:
// GetUserProfileResponse)
int ggr_gen_nested(uint64_t seed, int64_t n, uint8_t* json, uint64_t json_cap, uint64_t* json_off, int32_t* kind,
                   uint8_t* wire, uint64_t wire_cap, uint64_t* wire_off) {
  Out j(json, json_cap), w(wire, wire_cap);
  for (int64_t i = 0; i < n; i++) {
    Rng r(seed + 0x100000001B3ull * (uint64_t)i);
    json_off[i] = j.n;
    wire_off[i] = w.n;
    if (r.unit() < 0.8) {
      kind[i] = 0;
      Node root;
      build_tree(r, root, 1);
      j.str("{\"root_node\":");
      node_json(j, root);
      j.ch('}');
      std::string wb;
      node_wire(wb, root);  // Echo(Node) -> Node
      w.put(wb.data(), wb.size());
    } else {
      kind[i] = 1;
      // {"document":{"content":..,"document_id":..,("simple_summary":..|"structured_metadata_wrapper":{"data":{..}}),"title":..}}
      bool structured = r.unit() < 0.5;
      j.str("{\"document\":{\"content\":");
      json_string(j, rand_text(r, 600, 1000));
      j.str(",\"document_id\":");
      json_string(j, rand_text(r, 8, 16));
      if (structured) {
        j.str(",\"structured_metadata_wrapper\":{\"data\":{");
        // 24 distinct keys, emitted in bytewise order as json.Marshal would
        std::vector<std::string> keys;
        for (int k = 0; k < 24; k++) {
          char buf[32];
          snprintf(buf, sizeof buf, "k%02d_", k);
          keys.push_back(std::string(buf) + rand_text(r, 4, 10));
        }
        for (int k = 0; k < 24; k++) {
          if (k) j.ch(',');
          json_string(j, keys[k]);
          j.ch(':');
          json_string(j, rand_text(r, 80, 120));
        }
        j.str("}}");
      } else {
        j.str(",\"simple_summary\":");
        json_string(j, rand_text(r, 2400, 3000));
      }
      j.str(",\"title\":");
      json_string(j, rand_text(r, 20, 40));
      j.str("}}");
      // reply: GetUserProfileResponse{profile{user_id,display_name,email,user_type,last_login}}
      std::string prof;
      put_len_field(prof, 1, rand_text(r, 6, 12));
      put_len_field(prof, 2, rand_text(r, 12, 24));
      put_len_field(prof, 3, rand_text(r, 12, 24));
      put_varint(prof, 4 << 3);
      put_varint(prof, r.range(1, 3));
      std::string ts;
      put_varint(ts, 1 << 3);
      put_varint(ts, 1600000000ull + r.below(200000000));
      static const uint32_t nanos[] = {0, 123000000, 123456000, 123456789};
      uint32_t ns = nanos[r.below(4)];
      if (ns) {
        put_varint(ts, 2 << 3);
        put_varint(ts, ns);
      }
      put_len_field(prof, 5, ts);
      std::string wb;
      put_len_field(wb, 1, prof);
      w.put(wb.data(), wb.size());
    }
  }
  json_off[n] = j.n;
  wire_off[n] = w.n;
  return (j.ok && w.ok) ? 0 : -1;
}

// config 4: bench.Blob replies {bytes data = 1 (payload_bytes random bytes); string name = 2}
int ggr_gen_blob(uint64_t seed, int64_t n, uint32_t payload_bytes, uint8_t* wire, uint64_t wire_cap, uint64_t* wire_off) {
  Out w(wire, wire_cap);
  std::string data(payload_bytes, '\0');
  for (int64_t i = 0; i < n; i++) {
    Rng r(seed + 0x100000001B3ull * (uint64_t)i);
    wire_off[i] = w.n;
    for (uint32_t k = 0; k + 8 <= payload_bytes; k += 8) {
      uint64_t v = r.next();
      memcpy(&data[k], &v, 8);
    }
    for (uint32_t k = payload_bytes & ~7u; k < payload_bytes; k++) data[k] = (char)r.next();
    std::string wb;
    put_len_field(wb, 1, data);
    std::string name = "blob-";
    for (int k = 0; k < 11; k++) name.push_back((char)('a' + r.below(26)));
    put_len_field(wb, 2, name);
    w.put(wb.data(), wb.size());
  }
  wire_off[n] = w.n;
  return w.ok ? 0 : -1;
}

}  // extern "C"
