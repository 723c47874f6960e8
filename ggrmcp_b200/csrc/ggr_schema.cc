// ggr_schema.cc - see ggr_schema.h.
#include "ggr_schema.h"

#include <algorithm>
#include <cstring>

namespace ggr {

// must match khash_mix / khash_finish in ggr_json_in.cuh
uint32_t key_hash(const uint8_t* p, size_t n) {
  uint32_t h = 0x811C9DC5u;
  for (size_t i = 0; i < n; i += 4) {
    uint32_t w = 0;
    for (size_t j = 0; j < 4 && i + j < n; j++) w |= (uint32_t)p[i + j] << (8 * j);
    h = (h ^ w) * 0x9E3779B1u;
    h ^= h >> 15;
  }
  h ^= (uint32_t)n;
  h *= 0x85EBCA6Bu;
  return h ^ (h >> 13);
}

namespace {

// ---- protowire cursor over descriptor.proto messages ----
struct Cur {
  const uint8_t* p;
  const uint8_t* e;
  bool ok = true;
  Cur(const uint8_t* b, size_t n) : p(b), e(b + n) {}
  bool more() const { return ok && p < e; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int s = 0; s < 70; s += 7) {
      if (p >= e) { ok = false; return 0; }
      uint8_t c = *p++;
      v |= (uint64_t)(c & 0x7f) << s;
      if (!(c & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  // reads one field header + value; LEN payload returned as (ptr,len), scalars in val
  bool field(uint32_t* num, int* wt, uint64_t* val, const uint8_t** lp, size_t* ln) {
    uint64_t tag = varint();
    if (!ok) return false;
    *num = (uint32_t)(tag >> 3);
    *wt = (int)(tag & 7);
    *val = 0; *lp = nullptr; *ln = 0;
    switch (*wt) {
      case 0: *val = varint(); return ok;
      case 1: if (e - p < 8) return ok = false; memcpy(val, p, 8); p += 8; return true;
      case 5: if (e - p < 4) return ok = false; { uint32_t x; memcpy(&x, p, 4); *val = x; } p += 4; return true;
      case 2: {
        uint64_t l = varint();
        if (!ok || (uint64_t)(e - p) < l) return ok = false;
        *lp = p; *ln = (size_t)l; p += l;
        return true;
      }
      default: return ok = false;
    }
  }
};
#define FOR_FIELDS(cur) \
  uint32_t num; int wt; uint64_t val; const uint8_t* lp; size_t ln; \
  while ((cur).more() && (cur).field(&num, &wt, &val, &lp, &ln))

struct FieldIn {
  std::string name, json_name, type_name;
  int32_t number = 0;
  int label = 1, type = 0;
  int oneof_index = -1;
  bool proto3_optional = false, have_json = false;
  int packed_opt = 0;  // 0 unset, 1 true, 2 false
  int decl = 0;
  int child = -1;
};
struct MsgIn {
  std::string full_name;
  std::vector<FieldIn> fields;
  int n_oneofs = 0;
  bool map_entry = false, proto3 = true;
};
struct EnumIn {
  std::string full_name;
  std::vector<std::pair<std::string, int32_t>> values;
};
struct MethodIn {
  std::string service, name, in, out;
  bool cs = false, ss = false;
};

struct Parser {
  std::vector<MsgIn> msgs;
  std::vector<EnumIn> enums;
  std::vector<MethodIn> methods;
  bool bad = false;

  static std::string str(const uint8_t* p, size_t n) { return std::string((const char*)p, n); }
  static std::string join(const std::string& scope, const std::string& name) { return scope.empty() ? name : scope + "." + name; }

  void parse_enum(const uint8_t* b, size_t n, const std::string& scope) {
    EnumIn e;
    Cur c(b, n);
    FOR_FIELDS(c) {
      if (num == 1 && wt == 2) e.full_name = join(scope, str(lp, ln));
      else if (num == 2 && wt == 2) {
        Cur v(lp, ln);
        std::string nm; int32_t number = 0;
        { FOR_FIELDS(v) { if (num == 1 && wt == 2) nm = str(lp, ln); else if (num == 2 && wt == 0) number = (int32_t)val; } }
        if (!v.ok) bad = true;
        e.values.push_back({nm, number});
      }
    }
    if (!c.ok) bad = true;
    enums.push_back(e);
  }

  void parse_field(const uint8_t* b, size_t n, FieldIn& f) {
    Cur c(b, n);
    FOR_FIELDS(c) {
      switch (num) {
        case 1: f.name = str(lp, ln); break;
        case 3: f.number = (int32_t)val; break;
        case 4: f.label = (int)val; break;
        case 5: f.type = (int)val; break;
        case 6: f.type_name = str(lp, ln); break;
        case 8: {
          Cur o(lp, ln);
          { FOR_FIELDS(o) { if (num == 2 && wt == 0) f.packed_opt = val ? 1 : 2; } }
          break;
        }
        case 9: f.oneof_index = (int)val; break;
        case 10: f.json_name = str(lp, ln); f.have_json = true; break;
        case 17: f.proto3_optional = val != 0; break;
        default: break;
      }
    }
    if (!c.ok) bad = true;
  }

  void parse_message(const uint8_t* b, size_t n, const std::string& scope, bool proto3) {
    MsgIn m;
    m.proto3 = proto3;
    std::vector<std::pair<const uint8_t*, size_t>> nested, nenums;
    Cur c(b, n);
    std::string simple;
    FOR_FIELDS(c) {
      switch (num) {
        case 1: simple = str(lp, ln); break;
        case 2: {
          FieldIn f;
          parse_field(lp, ln, f);
          f.decl = (int)m.fields.size();
          m.fields.push_back(f);
          break;
        }
        case 3: nested.push_back({lp, ln}); break;
        case 4: nenums.push_back({lp, ln}); break;
        case 7: {
          Cur o(lp, ln);
          { FOR_FIELDS(o) { if (num == 7 && wt == 0) m.map_entry = val != 0; } }
          break;
        }
        case 8: m.n_oneofs++; break;
        default: break;
      }
    }
    if (!c.ok) bad = true;
    m.full_name = join(scope, simple);
    std::string fq = m.full_name;
    msgs.push_back(m);
    for (auto& e : nenums) parse_enum(e.first, e.second, fq);
    for (auto& s : nested) parse_message(s.first, s.second, fq, proto3);
  }

  void parse_file(const uint8_t* b, size_t n) {
    std::string pkg, syntax;
    {
      Cur c(b, n);
      FOR_FIELDS(c) {
        if (num == 2 && wt == 2) pkg = str(lp, ln);
        else if (num == 12 && wt == 2) syntax = str(lp, ln);
      }
      if (!c.ok) bad = true;
    }
    bool proto3 = syntax == "proto3";
    Cur c(b, n);
    FOR_FIELDS(c) {
      if (num == 4 && wt == 2) parse_message(lp, ln, pkg, proto3);
      else if (num == 5 && wt == 2) parse_enum(lp, ln, pkg);
      else if (num == 6 && wt == 2) {
        std::string sname;
        std::vector<std::pair<const uint8_t*, size_t>> ms;
        Cur s(lp, ln);
        { FOR_FIELDS(s) { if (num == 1 && wt == 2) sname = str(lp, ln); else if (num == 2 && wt == 2) ms.push_back({lp, ln}); } }
        for (auto& mm : ms) {
          MethodIn mi;
          mi.service = join(pkg, sname);
          Cur q(mm.first, mm.second);
          FOR_FIELDS(q) {
            if (num == 1 && wt == 2) mi.name = str(lp, ln);
            else if (num == 2 && wt == 2) mi.in = str(lp, ln);
            else if (num == 3 && wt == 2) mi.out = str(lp, ln);
            else if (num == 5 && wt == 0) mi.cs = val != 0;
            else if (num == 6 && wt == 0) mi.ss = val != 0;
          }
          methods.push_back(mi);
        }
      }
    }
  }
};

std::string default_json_name(const std::string& s) {  // protoc / strs.JSONCamelCase
  std::string o;
  bool up = false;
  for (char ch : s) {
    if (ch == '_') { up = true; continue; }
    if (up && ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
    up = false;
    o.push_back(ch);
  }
  return o;
}

// protojson's JSON string escaping for the key text written on the response side
std::string json_quote(const std::string& s) {
  static const char hex[] = "0123456789abcdef";
  std::string o = "\"";
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back((char)c); }
    else if (c == '\b') o += "\\b";
    else if (c == '\f') o += "\\f";
    else if (c == '\n') o += "\\n";
    else if (c == '\r') o += "\\r";
    else if (c == '\t') o += "\\t";
    else if (c < 0x20) { o += "\\u00"; o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
    else o.push_back((char)c);
  }
  o.push_back('"');
  return o;
}

int wire_type_for(int kind) {
  switch (kind) {
    case GK_DOUBLE: case GK_FIXED64: case GK_SFIXED64: return 1;
    case GK_FLOAT: case GK_FIXED32: case GK_SFIXED32: return 5;
    case GK_STRING: case GK_BYTES: case GK_MESSAGE: return 2;
    case GK_GROUP: return 3;
    default: return 0;
  }
}
int varint_len(uint64_t v) { int n = 1; while (v >= 0x80) { v >>= 7; n++; } return n; }

struct Blob {
  std::vector<GgrMsg> msgs;
  std::vector<GgrField> fields;
  std::vector<GgrEnum> enums;
  std::vector<GgrEnumValue> evals;
  std::vector<GgrHashEnt> hash;
  std::vector<uint16_t> u16;
  std::string pool;
  uint32_t intern(const std::string& s) {
    uint32_t off = (uint32_t)pool.size();
    pool += s;
    return off;
  }
  // open-addressing table over (name -> value); returns first index, sets mask
  uint32_t add_table(const std::vector<std::pair<std::string, int32_t>>& items, uint32_t* mask) {
    uint32_t size = 4;
    while (size < items.size() * 2 + 1) size <<= 1;
    uint32_t first = (uint32_t)hash.size();
    GgrHashEnt empty = {0, 0, 0xFFFFFFFFu, -1, {0, 0, 0, 0}};
    hash.resize(first + size, empty);
    for (auto& it : items) {
      uint32_t h = key_hash((const uint8_t*)it.first.data(), it.first.size());
      uint32_t slot = h & (size - 1);
      while (hash[first + slot].name_len != 0xFFFFFFFFu) slot = (slot + 1) & (size - 1);
      GgrHashEnt& e = hash[first + slot];
      e.hash = h;
      e.name_off = intern(it.first);
      e.name_len = (uint32_t)it.first.size();
      e.value = it.second;
      for (size_t j = 0; j < 16 && j < it.first.size(); j++) e.w[j / 4] |= (uint32_t)(uint8_t)it.first[j] << (8 * (j % 4));
    }
    *mask = size - 1;
    return first;
  }
};

template <class T>
uint32_t append_section(std::vector<uint8_t>& out, const T* p, size_t n_bytes) {
  while (out.size() % 16) out.push_back(0);
  uint32_t off = (uint32_t)out.size();
  out.insert(out.end(), (const uint8_t*)p, (const uint8_t*)p + n_bytes);
  return off;
}

}  // namespace

bool compile_schema(const uint8_t* fds, size_t n, WireOrder order, CompiledSchema* out, std::string* err, bool short_service_names) {
  Parser P;
  {
    Cur c(fds, n);
    FOR_FIELDS(c) {
      if (num == 1 && wt == 2) P.parse_file(lp, ln);
    }
    if (!c.ok || P.bad) { *err = "malformed FileDescriptorSet"; return false; }
  }
  if (P.msgs.size() > 65535) { *err = "too many message types"; return false; }
  std::map<std::string, int> msg_by, enum_by;
  for (size_t i = 0; i < P.msgs.size(); i++) msg_by[P.msgs[i].full_name] = (int)i;
  for (size_t i = 0; i < P.enums.size(); i++) enum_by[P.enums[i].full_name] = (int)i;
  auto strip = [](const std::string& s) { return (!s.empty() && s[0] == '.') ? s.substr(1) : s; };

  Blob B;
  // ---- enums ----
  for (auto& e : P.enums) {
    GgrEnum ge;
    std::vector<std::pair<int32_t, std::string>> by_num;  // first declared name per number
    for (auto& v : e.values) {
      bool seen = false;
      for (auto& x : by_num) if (x.first == v.second) seen = true;
      if (!seen) by_num.push_back({v.second, v.first});
    }
    std::stable_sort(by_num.begin(), by_num.end(), [](const std::pair<int32_t, std::string>& a, const std::pair<int32_t, std::string>& b) { return a.first < b.first; });
    ge.val_first = (uint32_t)B.evals.size();
    ge.n_vals = (uint32_t)by_num.size();
    for (auto& x : by_num) {
      GgrEnumValue ev = {x.first, B.intern(x.second), (uint32_t)x.second.size(), 0};
      B.evals.push_back(ev);
    }
    std::vector<std::pair<std::string, int32_t>> items;  // ByName: names are unique within an enum
    for (auto& v : e.values) {
      bool dup = false;
      for (auto& it : items) if (it.first == v.first) dup = true;
      if (!dup) items.push_back({v.first, v.second});
    }
    ge.hash_first = B.add_table(items, &ge.hash_mask);
    B.enums.push_back(ge);
  }
  // ---- messages ----
  out->max_msg_fields = 0;
  for (size_t mi = 0; mi < P.msgs.size(); mi++) {
    MsgIn& m = P.msgs[mi];
    if (m.fields.size() > 4096) { *err = "message " + m.full_name + " has too many fields"; return false; }
    for (auto& f : m.fields) {
      if (f.type == GK_MESSAGE || f.type == GK_GROUP) {
        auto it = msg_by.find(strip(f.type_name));
        if (it == msg_by.end()) { *err = "unresolved type " + f.type_name; return false; }
        f.child = it->second;
      } else if (f.type == GK_ENUM) {
        auto it = enum_by.find(strip(f.type_name));
        if (it == enum_by.end()) { *err = "unresolved enum " + f.type_name; return false; }
        f.child = it->second;
      }
      if (!f.have_json) f.json_name = default_json_name(f.name);
    }
  }
  for (size_t mi = 0; mi < P.msgs.size(); mi++) {
    MsgIn& m = P.msgs[mi];
    GgrMsg gm;
    memset(&gm, 0, sizeof gm);
    // emit order
    std::vector<int> emit(m.fields.size());
    for (size_t i = 0; i < emit.size(); i++) emit[i] = (int)i;
    auto real_oneof = [&](const FieldIn& f) { return f.oneof_index >= 0 && !f.proto3_optional; };
    if (order == ORDER_GO_LEGACY) {
      std::stable_sort(emit.begin(), emit.end(), [&](int a, int b) {
        const FieldIn& x = m.fields[a]; const FieldIn& y = m.fields[b];
        bool ox = real_oneof(x), oy = real_oneof(y);
        if (ox != oy) return !ox && oy;
        if (ox && oy && x.oneof_index != y.oneof_index) return x.oneof_index < y.oneof_index;
        return x.number < y.number;
      });
    } else {
      std::stable_sort(emit.begin(), emit.end(), [&](int a, int b) { return m.fields[a].number < m.fields[b].number; });
    }
    gm.field_first = (uint32_t)B.fields.size();
    gm.n_fields = (uint16_t)m.fields.size();
    gm.n_oneofs = (uint32_t)m.n_oneofs;
    if (m.map_entry) gm.flags |= GM_MAP_ENTRY;
    if (m.full_name == "google.protobuf.Timestamp") gm.wkt = GGR_WKT_TIMESTAMP;
    else if (m.full_name == "google.protobuf.Duration") gm.wkt = GGR_WKT_DURATION;
    else if (m.full_name == "google.protobuf.Empty") gm.wkt = GGR_WKT_EMPTY;
    else if (m.full_name == "google.protobuf.FieldMask" && m.fields.size() == 1 && m.fields[0].number == 1) gm.wkt = GGR_WKT_FIELDMASK;
    else if (m.full_name.rfind("google.protobuf.", 0) == 0) {
      // the nine wrappers: the JSON form is the bare value of field 1 (protojson marshalWrapperType)
      static const char* wr[] = {"DoubleValue", "FloatValue", "Int64Value", "UInt64Value", "Int32Value", "UInt32Value", "BoolValue",
                                 "StringValue", "BytesValue"};
      for (const char* w : wr)
        if (m.full_name == std::string("google.protobuf.") + w && m.fields.size() == 1 && m.fields[0].number == 1) gm.wkt = GGR_WKT_WRAPPER;
      // recognised and refused: dynamic JSON (Struct / Value / ListValue), type URLs (Any)
      static const char* wk[] = {"Any", "Struct", "Value", "ListValue"};
      for (const char* w : wk) if (m.full_name == std::string("google.protobuf.") + w) gm.wkt = GGR_WKT_UNSUPPORTED;
    }
    bool decl_is_emit = true;
    std::vector<uint16_t> decl_perm(m.fields.size());
    uint32_t max_num = 0;
    for (size_t ei = 0; ei < emit.size(); ei++) {
      const FieldIn& f = m.fields[emit[ei]];
      if ((size_t)f.decl != ei) decl_is_emit = false;
      decl_perm[f.decl] = (uint16_t)ei;
      GgrField gf;
      memset(&gf, 0, sizeof gf);
      gf.number = (uint32_t)f.number;
      gf.kind = (uint8_t)f.type;
      gf.wt = (uint8_t)wire_type_for(f.type);
      bool repeated = f.label == 3;
      bool is_map = repeated && f.type == GK_MESSAGE && P.msgs[f.child].map_entry;
      bool packable = repeated && f.type != GK_STRING && f.type != GK_BYTES && f.type != GK_MESSAGE && f.type != GK_GROUP;
      bool packed = packable && (m.proto3 ? f.packed_opt != 2 : f.packed_opt == 1);
      if (repeated) gf.flags |= GF_REPEATED;
      if (is_map) gf.flags |= GF_MAP;
      if (packable) gf.flags |= GF_PACKABLE;
      if (packed) gf.flags |= GF_PACKED;
      if (!repeated) {
        if (f.type == GK_MESSAGE || f.type == GK_GROUP || f.oneof_index >= 0 || !m.proto3) gf.flags |= GF_PRESENCE;
      }
      gf.tag = ((uint32_t)f.number << 3) | (uint32_t)(packed ? 2 : gf.wt);
      gf.tag_len = (uint8_t)varint_len(gf.tag);
      gf.oneof = real_oneof(f) ? (int16_t)f.oneof_index : (int16_t)-1;
      gf.decl_index = (uint16_t)f.decl;
      gf.child = f.child;
      std::string key = json_quote(f.json_name) + ":";
      gf.name_off = B.intern(key);
      gf.name_len = (uint16_t)key.size();
      B.fields.push_back(gf);
      if ((uint32_t)f.number > max_num) max_num = (uint32_t)f.number;
    }
    if (decl_is_emit) gm.flags |= GM_DECL_IS_EMIT;
    gm.decl_first = (uint32_t)B.u16.size();
    B.u16.insert(B.u16.end(), decl_perm.begin(), decl_perm.end());
    // number LUT (dense part only)
    uint32_t lut_n = max_num + 1 > 1024 ? 1024 : max_num + 1;
    gm.lut_first = (uint32_t)B.u16.size();
    gm.lut_n = lut_n;
    B.u16.resize(B.u16.size() + lut_n, 0);
    for (size_t ei = 0; ei < emit.size(); ei++) {
      const FieldIn& f = m.fields[emit[ei]];
      if ((uint32_t)f.number < lut_n && B.u16[gm.lut_first + f.number] == 0) B.u16[gm.lut_first + f.number] = (uint16_t)(ei + 1);
    }
    // JSON key table: ByJSONName first, then ByTextName; first declared wins within each
    // [upstream filedesc Fields.lazyInit; protojson unmarshalMessage]
    std::map<std::string, int> by_json, by_text;
    for (auto& f : m.fields) {
      if (!by_json.count(f.json_name)) by_json[f.json_name] = f.decl;
      if (!by_text.count(f.name)) by_text[f.name] = f.decl;
    }
    std::vector<std::pair<std::string, int32_t>> items;
    for (auto& kv : by_json) items.push_back({kv.first, (int32_t)decl_perm[kv.second]});
    for (auto& kv : by_text) if (!by_json.count(kv.first)) items.push_back({kv.first, (int32_t)decl_perm[kv.second]});
    gm.key_hash_first = B.add_table(items, &gm.key_hash_mask);
    B.msgs.push_back(gm);
    if (m.fields.size() > out->max_msg_fields) out->max_msg_fields = (uint32_t)m.fields.size();
    out->msg_index[m.full_name] = (int32_t)mi;
    out->msg_names.push_back(m.full_name);
  }
  // ---- methods ----
  for (auto& mi : P.methods) {
    MethodInfo M;
    M.name = mi.name;
    M.service_name = mi.service;
    if (short_service_names) {  // extractServiceNameForCompatibility: the last two segments of the full name
      size_t d2 = mi.service.rfind('.');
      if (d2 != std::string::npos && d2 > 0) {
        size_t d1 = mi.service.rfind('.', d2 - 1);
        if (d1 != std::string::npos) M.service_name = mi.service.substr(d1 + 1);
      }
    }
    M.full_name = mi.service + "." + mi.name;
    M.input_type = mi.in;
    M.output_type = mi.out;
    auto a = msg_by.find(strip(mi.in));
    auto b = msg_by.find(strip(mi.out));
    if (a == msg_by.end() || b == msg_by.end()) { *err = "unresolved method type in " + M.full_name; return false; }
    M.input_msg = a->second;
    M.output_msg = b->second;
    M.client_streaming = mi.cs;
    M.server_streaming = mi.ss;
    std::string t = M.service_name;
    for (auto& ch : t) { if (ch == '.') ch = '_'; else if (ch >= 'A' && ch <= 'Z') ch = (char)(ch + 32); }
    std::string mn = mi.name;
    for (auto& ch : mn) if (ch >= 'A' && ch <= 'Z') ch = (char)(ch + 32);
    M.tool_name = t + "_" + mn;
    M.grpc_path = "/" + mi.service + "/" + mi.name;
    out->tool_index[M.tool_name] = (int32_t)out->methods.size();
    out->methods.push_back(M);
  }
  // ---- tool table for the request envelope: tool name -> method index (same key tables as the
  // message keys), u16[tool_methods_first + m] = input message of method m (0xFFFF: streaming) ----
  uint32_t tool_hash_first = 0, tool_hash_mask = 0, tool_methods_first = 0;
  {
    std::vector<std::pair<std::string, int32_t>> items;
    for (size_t m = 0; m < out->methods.size(); m++) items.push_back({out->methods[m].tool_name, (int32_t)m});
    tool_hash_first = B.add_table(items, &tool_hash_mask);
    tool_methods_first = (uint32_t)B.u16.size();
    for (auto& M : out->methods)
      B.u16.push_back((M.client_streaming || M.server_streaming || M.input_msg >= 0xFFFF) ? (uint16_t)0xFFFF : (uint16_t)M.input_msg);
  }
  // ---- serialize ----
  GgrSchemaHdr h;
  memset(&h, 0, sizeof h);
  std::vector<uint8_t>& o = out->blob;
  o.clear();
  o.resize(sizeof h, 0);
  h.magic = GGR_SCHEMA_MAGIC;
  h.n_msgs = (uint32_t)B.msgs.size();
  h.msgs_off = append_section(o, B.msgs.data(), B.msgs.size() * sizeof(GgrMsg));
  h.n_fields = (uint32_t)B.fields.size();
  h.fields_off = append_section(o, B.fields.data(), B.fields.size() * sizeof(GgrField));
  h.n_enums = (uint32_t)B.enums.size();
  h.enums_off = append_section(o, B.enums.data(), B.enums.size() * sizeof(GgrEnum));
  h.n_evals = (uint32_t)B.evals.size();
  h.evals_off = append_section(o, B.evals.data(), B.evals.size() * sizeof(GgrEnumValue));
  h.n_hash = (uint32_t)B.hash.size();
  h.hash_off = append_section(o, B.hash.data(), B.hash.size() * sizeof(GgrHashEnt));
  h.n_u16 = (uint32_t)B.u16.size();
  h.u16_off = append_section(o, B.u16.data(), B.u16.size() * sizeof(uint16_t));
  B.pool.append(32, '\0');  // readers fetch 16-byte chunks
  h.pool_bytes = (uint32_t)B.pool.size();
  h.pool_off = append_section(o, B.pool.data(), B.pool.size());
  while (o.size() % 16) o.push_back(0);
  {  // trailer (the last 16 bytes of the blob): GgrToolsTrailer
    GgrToolsTrailer tt = {tool_hash_first, tool_hash_mask, tool_methods_first, (uint32_t)out->methods.size()};
    append_section(o, &tt, sizeof tt);
  }
  h.total_bytes = (uint32_t)o.size();
  memcpy(o.data(), &h, sizeof h);
  return true;
}

}  // namespace ggr
