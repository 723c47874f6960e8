// orc_schema.h - descriptor model of the CPU oracle (TEST INFRASTRUCTURE, see ggr_oracle.h).
//
// Parses a serialized google.protobuf.FileDescriptorSet (what
// /root/reference/pkg/descriptors/loader.go:33-64 reads from disk, and what the reflection route
// of /root/reference/pkg/grpc/reflection.go:196-254 receives file by file) into the few facts the
// transcode path needs from protoreflect descriptors:
//   field number / kind / cardinality / packedness / presence / oneof / json name / decl index,
//   enum name<->number, map-entry flag, well-known-type flag, service methods.
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "orc_util.h"

namespace orc {

// FieldDescriptorProto.Type values
enum {
  T_DOUBLE = 1, T_FLOAT = 2, T_INT64 = 3, T_UINT64 = 4, T_INT32 = 5, T_FIXED64 = 6,
  T_FIXED32 = 7, T_BOOL = 8, T_STRING = 9, T_GROUP = 10, T_MESSAGE = 11, T_BYTES = 12,
  T_UINT32 = 13, T_ENUM = 14, T_SFIXED32 = 15, T_SFIXED64 = 16, T_SINT32 = 17, T_SINT64 = 18
};
enum { WKT_NONE = 0, WKT_TIMESTAMP = 1, WKT_DURATION = 2, WKT_WRAPPER = 3, WKT_EMPTY = 4, WKT_FIELDMASK = 5, WKT_OTHER = 99 };

struct EnumDesc {
  std::string full_name;
  std::vector<std::pair<std::string, int32_t>> values;  // declaration order
  const std::string* name_of(int32_t num) const {       // ByNumber: first declared wins
    for (auto& v : values)
      if (v.second == num) return &v.first;
    return nullptr;
  }
  bool number_of(const std::string& name, int32_t& num) const {
    for (auto& v : values)
      if (v.first == name) {
        num = v.second;
        return true;
      }
    return false;
  }
};

struct FieldDesc {
  std::string name, json_name, type_name;
  int32_t number = 0;
  int type = 0;
  bool repeated = false;
  bool packed = false;
  bool has_presence = false;
  bool is_map = false;
  bool proto3_optional = false;
  int oneof_index = -1;  // real oneofs only (synthetic -> -1, has_presence = true)
  int raw_oneof_index = -1;
  int msg = -1;   // index into Schema::msgs for T_MESSAGE/T_GROUP
  int enm = -1;   // index into Schema::enums for T_ENUM
  int index = 0;  // declaration index
  int has_packed_opt = 0;  // 0 unset, 1 true, 2 false
};

struct MsgDesc {
  std::string full_name;
  std::vector<FieldDesc> fields;  // declaration order
  std::vector<std::string> oneofs;
  bool map_entry = false;
  bool proto3 = true;
  int wkt = WKT_NONE;
  std::map<int32_t, int> by_number;
  std::map<std::string, int> by_json, by_text;  // first declared wins [upstream filedesc lazyInit]
  const FieldDesc* find_number(int32_t n) const {
    auto it = by_number.find(n);
    return it == by_number.end() ? nullptr : &fields[it->second];
  }
  // protojson field lookup: JSON name first, then text (proto) name
  // [upstream encoding/protojson/decode.go unmarshalMessage]
  const FieldDesc* find_json_key(const std::string& k) const {
    auto it = by_json.find(k);
    if (it != by_json.end()) return &fields[it->second];
    it = by_text.find(k);
    if (it != by_text.end()) return &fields[it->second];
    return nullptr;
  }
};

struct MethodDesc {
  std::string service_full, name, tool_name, path;
  int input = -1, output = -1;
  bool client_streaming = false, server_streaming = false;
};

struct Schema {
  std::vector<MsgDesc> msgs;
  std::vector<EnumDesc> enums;
  std::map<std::string, int> msg_by_name, enum_by_name;
  std::vector<MethodDesc> methods;
  std::map<std::string, int> method_by_tool;
};

// ---- minimal protowire reader for descriptor.proto messages ----
struct PField {
  uint32_t num;
  int wt;
  uint64_t v;          // varint / fixed
  const uint8_t* p;    // LEN payload
  size_t n;
};
inline bool next_field(const uint8_t*& p, const uint8_t* e, PField& f) {
  uint64_t tag;
  if (!get_varint(p, e, tag)) return false;
  f.num = (uint32_t)(tag >> 3);
  f.wt = (int)(tag & 7);
  f.v = 0;
  f.p = nullptr;
  f.n = 0;
  switch (f.wt) {
    case 0:
      return get_varint(p, e, f.v);
    case 1:
      if (e - p < 8) return false;
      memcpy(&f.v, p, 8);
      p += 8;
      return true;
    case 5:
      if (e - p < 4) return false;
      {
        uint32_t x;
        memcpy(&x, p, 4);
        f.v = x;
      }
      p += 4;
      return true;
    case 2: {
      uint64_t len;
      if (!get_varint(p, e, len)) return false;
      if ((uint64_t)(e - p) < len) return false;
      f.p = p;
      f.n = (size_t)len;
      p += len;
      return true;
    }
    default:
      return false;
  }
}
inline std::string pstr(const PField& f) { return std::string((const char*)f.p, f.n); }

// protoc's default json_name [upstream internal/strs JSONCamelCase]
inline std::string json_camel(const std::string& s) {
  std::string out;
  bool up = false;
  for (char c : s) {
    if (c == '_') {
      up = true;
      continue;
    }
    if (up && c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
    up = false;
    out.push_back(c);
  }
  return out;
}

struct SchemaBuilder {
  Schema& S;
  std::string err;
  struct PendingSvc {
    std::string service_full, name, in, out;
    bool cs, ss;
  };
  std::vector<PendingSvc> pending;
  bool short_service_names = false;  // tool names of the FileDescriptorSet route
  explicit SchemaBuilder(Schema& s) : S(s) {}

  bool parse_enum(const uint8_t* p, size_t n, const std::string& scope) {
    EnumDesc e;
    const uint8_t* end = p + n;
    PField f;
    while (p < end) {
      if (!next_field(p, end, f)) return false;
      if (f.num == 1 && f.wt == 2)
        e.full_name = scope.empty() ? pstr(f) : scope + "." + pstr(f);
      else if (f.num == 2 && f.wt == 2) {
        const uint8_t* q = f.p;
        const uint8_t* qe = f.p + f.n;
        PField g;
        std::string nm;
        int32_t num = 0;
        while (q < qe) {
          if (!next_field(q, qe, g)) return false;
          if (g.num == 1 && g.wt == 2) nm = pstr(g);
          if (g.num == 2 && g.wt == 0) num = (int32_t)g.v;
        }
        e.values.push_back({nm, num});
      }
    }
    S.enum_by_name[e.full_name] = (int)S.enums.size();
    S.enums.push_back(e);
    return true;
  }

  bool parse_field(const uint8_t* p, size_t n, FieldDesc& fd) {
    const uint8_t* end = p + n;
    PField f;
    int label = 1;
    bool have_json = false;
    while (p < end) {
      if (!next_field(p, end, f)) return false;
      switch (f.num) {
        case 1: fd.name = pstr(f); break;
        case 3: fd.number = (int32_t)f.v; break;
        case 4: label = (int)f.v; break;
        case 5: fd.type = (int)f.v; break;
        case 6: fd.type_name = pstr(f); break;
        case 8: {  // FieldOptions
          const uint8_t* q = f.p;
          const uint8_t* qe = f.p + f.n;
          PField g;
          while (q < qe) {
            if (!next_field(q, qe, g)) return false;
            if (g.num == 2 && g.wt == 0) fd.has_packed_opt = g.v ? 1 : 2;
          }
          break;
        }
        case 9: fd.raw_oneof_index = (int)f.v; break;
        case 10: fd.json_name = pstr(f); have_json = true; break;
        case 17: fd.proto3_optional = f.v != 0; break;
        default: break;
      }
    }
    fd.repeated = label == 3;
    if (!have_json) fd.json_name = json_camel(fd.name);
    return true;
  }

  bool parse_message(const uint8_t* p, size_t n, const std::string& scope, bool proto3) {
    MsgDesc m;
    m.proto3 = proto3;
    const uint8_t* end = p + n;
    PField f;
    std::vector<std::pair<const uint8_t*, size_t>> nested, enums;
    // first pass: name
    {
      const uint8_t* q = p;
      while (q < end) {
        if (!next_field(q, end, f)) return false;
        if (f.num == 1 && f.wt == 2) m.full_name = scope.empty() ? pstr(f) : scope + "." + pstr(f);
      }
    }
    while (p < end) {
      if (!next_field(p, end, f)) return false;
      switch (f.num) {
        case 2: {
          FieldDesc fd;
          if (!parse_field(f.p, f.n, fd)) return false;
          fd.index = (int)m.fields.size();
          m.fields.push_back(fd);
          break;
        }
        case 3: nested.push_back({f.p, f.n}); break;
        case 4: enums.push_back({f.p, f.n}); break;
        case 7: {
          const uint8_t* q = f.p;
          const uint8_t* qe = f.p + f.n;
          PField g;
          while (q < qe) {
            if (!next_field(q, qe, g)) return false;
            if (g.num == 7 && g.wt == 0) m.map_entry = g.v != 0;
          }
          break;
        }
        case 8: {
          const uint8_t* q = f.p;
          const uint8_t* qe = f.p + f.n;
          PField g;
          std::string nm;
          while (q < qe) {
            if (!next_field(q, qe, g)) return false;
            if (g.num == 1 && g.wt == 2) nm = pstr(g);
          }
          m.oneofs.push_back(nm);
          break;
        }
        default: break;
      }
    }
    if (m.full_name == "google.protobuf.Timestamp") m.wkt = WKT_TIMESTAMP;
    else if (m.full_name == "google.protobuf.Duration") m.wkt = WKT_DURATION;
    else if (m.full_name == "google.protobuf.Empty") m.wkt = WKT_EMPTY;
    else if (m.full_name == "google.protobuf.FieldMask") m.wkt = WKT_FIELDMASK;
    else if (m.full_name == "google.protobuf.Any" || m.full_name == "google.protobuf.Struct" ||
             m.full_name == "google.protobuf.Value" || m.full_name == "google.protobuf.ListValue")
      m.wkt = WKT_OTHER;
    else {
      // [upstream genid wrappers: BoolValue, Int32Value, Int64Value, UInt32Value, UInt64Value, FloatValue, DoubleValue,
      //  StringValue, BytesValue]
      static const char* const wr[] = {"BoolValue", "Int32Value", "Int64Value", "UInt32Value", "UInt64Value", "FloatValue",
                                       "DoubleValue", "StringValue", "BytesValue"};
      for (const char* w : wr)
        if (m.full_name == std::string("google.protobuf.") + w) m.wkt = WKT_WRAPPER;
    }
    std::string fq = m.full_name;
    S.msg_by_name[fq] = (int)S.msgs.size();
    S.msgs.push_back(m);
    for (auto& e : enums)
      if (!parse_enum(e.first, e.second, fq)) return false;
    for (auto& nm : nested)
      if (!parse_message(nm.first, nm.second, fq, proto3)) return false;
    return true;
  }

  bool parse_file(const uint8_t* p, size_t n) {
    const uint8_t* end = p + n;
    PField f;
    std::string pkg, syntax;
    {
      const uint8_t* q = p;
      while (q < end) {
        if (!next_field(q, end, f)) return false;
        if (f.num == 2 && f.wt == 2) pkg = pstr(f);
        if (f.num == 12 && f.wt == 2) syntax = pstr(f);
      }
    }
    bool proto3 = syntax == "proto3";
    while (p < end) {
      if (!next_field(p, end, f)) return false;
      if (f.num == 4 && f.wt == 2) {
        if (!parse_message(f.p, f.n, pkg, proto3)) return false;
      } else if (f.num == 5 && f.wt == 2) {
        if (!parse_enum(f.p, f.n, pkg)) return false;
      } else if (f.num == 6 && f.wt == 2) {
        const uint8_t* q = f.p;
        const uint8_t* qe = f.p + f.n;
        PField g;
        std::string sname;
        std::vector<std::pair<const uint8_t*, size_t>> ms;
        while (q < qe) {
          if (!next_field(q, qe, g)) return false;
          if (g.num == 1 && g.wt == 2) sname = pstr(g);
          if (g.num == 2 && g.wt == 2) ms.push_back({g.p, g.n});
        }
        std::string sfull = pkg.empty() ? sname : pkg + "." + sname;
        for (auto& mm : ms) {
          const uint8_t* r = mm.first;
          const uint8_t* re = mm.first + mm.second;
          PField h;
          PendingSvc ps{sfull, "", "", "", false, false};
          while (r < re) {
            if (!next_field(r, re, h)) return false;
            if (h.num == 1 && h.wt == 2) ps.name = pstr(h);
            if (h.num == 2 && h.wt == 2) ps.in = pstr(h);
            if (h.num == 3 && h.wt == 2) ps.out = pstr(h);
            if (h.num == 5 && h.wt == 0) ps.cs = h.v != 0;
            if (h.num == 6 && h.wt == 0) ps.ss = h.v != 0;
          }
          pending.push_back(ps);
        }
      }
    }
    return true;
  }

  static std::string strip_dot(const std::string& s) {
    return (!s.empty() && s[0] == '.') ? s.substr(1) : s;
  }

  bool resolve() {
    for (auto& m : S.msgs) {
      // which oneofs are synthetic (hold exactly one proto3_optional field)
      for (auto& f : m.fields) {
        if (f.type == T_MESSAGE || f.type == T_GROUP) {
          auto it = S.msg_by_name.find(strip_dot(f.type_name));
          if (it == S.msg_by_name.end()) {
            err = "unresolved message type " + f.type_name;
            return false;
          }
          f.msg = it->second;
        } else if (f.type == T_ENUM) {
          auto it = S.enum_by_name.find(strip_dot(f.type_name));
          if (it == S.enum_by_name.end()) {
            err = "unresolved enum type " + f.type_name;
            return false;
          }
          f.enm = it->second;
        }
      }
    }
    for (auto& m : S.msgs) {
      for (auto& f : m.fields) {
        f.is_map = f.repeated && f.type == T_MESSAGE && S.msgs[f.msg].map_entry;
        bool packable = f.repeated && f.type != T_STRING && f.type != T_BYTES &&
                        f.type != T_MESSAGE && f.type != T_GROUP;
        if (packable) {
          if (m.proto3) f.packed = f.has_packed_opt != 2;
          else f.packed = f.has_packed_opt == 1;
        }
        if (f.raw_oneof_index >= 0 && !f.proto3_optional) f.oneof_index = f.raw_oneof_index;
        if (!f.repeated) {
          if (f.type == T_MESSAGE || f.type == T_GROUP) f.has_presence = true;
          else if (f.raw_oneof_index >= 0) f.has_presence = true;  // real or synthetic oneof
          else if (!m.proto3) f.has_presence = true;               // proto2 optional/required
        }
        if (!m.by_number.count(f.number)) m.by_number[f.number] = f.index;
        if (!m.by_json.count(f.json_name)) m.by_json[f.json_name] = f.index;
        if (!m.by_text.count(f.name)) m.by_text[f.name] = f.index;
      }
    }
    for (auto& ps : pending) {
      MethodDesc md;
      md.service_full = ps.service_full;
      md.name = ps.name;
      auto a = S.msg_by_name.find(strip_dot(ps.in));
      auto b = S.msg_by_name.find(strip_dot(ps.out));
      if (a == S.msg_by_name.end() || b == S.msg_by_name.end()) {
        err = "unresolved method type";
        return false;
      }
      md.input = a->second;
      md.output = b->second;
      md.client_streaming = ps.cs;
      md.server_streaming = ps.ss;
      // types.MethodInfo.GenerateToolName, /root/reference/pkg/types/service.go:53-61
      std::string t = ps.service_full;
      if (short_service_names) {  // extractServiceNameForCompatibility, /root/reference/pkg/descriptors/loader.go:221-235
        size_t d2 = t.rfind('.');
        if (d2 != std::string::npos && d2 > 0) {
          size_t d1 = t.rfind('.', d2 - 1);
          if (d1 != std::string::npos) t = t.substr(d1 + 1);
        }
      }
      for (auto& c : t) {
        if (c == '.') c = '_';
        else if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
      }
      std::string mn = ps.name;
      for (auto& c : mn)
        if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
      md.tool_name = t + "_" + mn;
      // "/%s/%s" of FullName[:LastIndex(".")] and Name, /root/reference/pkg/grpc/reflection.go:367
      md.path = "/" + ps.service_full + "/" + ps.name;
      S.method_by_tool[md.tool_name] = (int)S.methods.size();
      S.methods.push_back(md);
    }
    return true;
  }

  bool build(const uint8_t* p, size_t n) {
    const uint8_t* end = p + n;
    PField f;
    while (p < end) {
      if (!next_field(p, end, f)) {
        err = "malformed FileDescriptorSet";
        return false;
      }
      if (f.num == 1 && f.wt == 2) {
        if (!parse_file(f.p, f.n)) {
          if (err.empty()) err = "malformed FileDescriptorProto";
          return false;
        }
      }
    }
    return resolve();
  }
};

}  // namespace orc
