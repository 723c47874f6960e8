// orc_dyn.h - dynamic message + proto wire marshal/unmarshal of the CPU oracle
// (TEST INFRASTRUCTURE, see ggr_oracle.h).
//
// Restates what grpc-go's proto codec does with the dynamicpb messages the reference hands to
// conn.Invoke (/root/reference/pkg/grpc/reflection.go:351,363,373):
//   proto.Marshal   [upstream proto/encode.go marshalMessageSlow, types/dynamicpb/dynamic.go Range/isSet]
//   proto.Unmarshal [upstream proto/decode.go unmarshalMessageSlow]
// Canonical field order: ascending field number, map entries sorted by
// key; ORC_F_GO_LEGACY_ORDER selects Go's order.LegacyFieldOrder instead.
#pragma once
#include <map>
#include <memory>

#include "ggr_oracle.h"
#include "orc_schema.h"

namespace orc {

struct DynMsg;
struct Val {
  uint64_t u = 0;  // bool / ints (sign-extended to 64) / enum / float32 bits / float64 bits
  Bytes s;         // string / bytes
  std::shared_ptr<DynMsg> m;
};
struct FieldVal {
  std::vector<Val> list;                    // singular: exactly one element
  std::vector<std::pair<Val, Val>> map;     // insertion order; sorted at marshal time
};
struct DynMsg {
  const MsgDesc* d = nullptr;
  std::map<int32_t, FieldVal> known;  // keyed by field number
};

struct Err {
  int code = ORC_OK;
  std::string msg;
  bool fail(int c, const std::string& m) {
    if (code == ORC_OK) {
      code = c;
      msg = m;
    }
    return false;
  }
};

inline double bits_to_double(uint64_t u) {
  double d;
  memcpy(&d, &u, 8);
  return d;
}
inline float bits_to_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint64_t double_bits(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}
inline uint32_t float_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

// dynamicpb isSet for singular fields [upstream types/dynamicpb/dynamic.go isSet]
inline bool scalar_is_set(const FieldDesc& f, const Val& v) {
  if (f.has_presence) return true;
  switch (f.type) {
    case T_STRING: case T_BYTES: return !v.s.empty();
    case T_FLOAT: return (uint32_t)v.u != 0;   // value != 0 || signbit  <=> any bit set
    case T_DOUBLE: return v.u != 0;
    default: return v.u != 0;
  }
}

inline int wire_type_of(int type) {
  switch (type) {
    case T_DOUBLE: case T_FIXED64: case T_SFIXED64: return 1;
    case T_FLOAT: case T_FIXED32: case T_SFIXED32: return 5;
    case T_STRING: case T_BYTES: case T_MESSAGE: return 2;
    case T_GROUP: return 3;
    default: return 0;
  }
}

struct WireMarshal {
  const Schema& S;
  uint32_t flags;
  WireMarshal(const Schema& s, uint32_t f) : S(s), flags(f) {}

  void put_tag(Bytes& b, int32_t num, int wt) { put_varint(b, ((uint64_t)(uint32_t)num << 3) | (uint64_t)wt); }

  void scalar_payload(Bytes& b, int type, const Val& v) {
    switch (type) {
      case T_BOOL: put_varint(b, v.u ? 1 : 0); break;
      case T_INT32: case T_INT64: case T_ENUM: put_varint(b, v.u); break;  // sign-extended
      case T_UINT32: put_varint(b, (uint32_t)v.u); break;
      case T_UINT64: put_varint(b, v.u); break;
      case T_SINT32: put_varint(b, (uint32_t)zigzag64((int64_t)(int32_t)v.u) ); break;
      case T_SINT64: put_varint(b, zigzag64((int64_t)v.u)); break;
      case T_FIXED32: case T_SFIXED32: case T_FLOAT: {
        uint32_t x = (uint32_t)v.u;
        b.append((const char*)&x, 4);
        break;
      }
      case T_FIXED64: case T_SFIXED64: case T_DOUBLE: {
        uint64_t x = v.u;
        b.append((const char*)&x, 8);
        break;
      }
      case T_STRING: case T_BYTES:
        put_varint(b, v.s.size());
        b += v.s;
        break;
      case T_MESSAGE: {
        Bytes sub;
        message(sub, *v.m);
        put_varint(b, sub.size());
        b += sub;
        break;
      }
      default: break;
    }
  }
  void field_one(Bytes& b, const FieldDesc& f, const Val& v) {
    put_tag(b, f.number, wire_type_of(f.type));
    scalar_payload(b, f.type, v);
  }

  static bool key_less(int ktype, const Val& a, const Val& b) {
    switch (ktype) {
      case T_STRING: return a.s < b.s;
      case T_BOOL: return a.u < b.u;
      case T_UINT32: case T_UINT64: case T_FIXED32: case T_FIXED64: return a.u < b.u;
      default: return (int64_t)a.u < (int64_t)b.u;
    }
  }

  void field(Bytes& b, const FieldDesc& f, const FieldVal& fv) {
    if (f.is_map) {
      const MsgDesc& e = S.msgs[f.msg];
      const FieldDesc& kf = *e.find_number(1);
      const FieldDesc& vf = *e.find_number(2);
      std::vector<const std::pair<Val, Val>*> ents;
      for (auto& kv : fv.map) ents.push_back(&kv);
      std::stable_sort(ents.begin(), ents.end(), [&](const std::pair<Val, Val>* x, const std::pair<Val, Val>* y) {
        return key_less(kf.type, x->first, y->first);
      });
      for (auto* kv : ents) {
        Bytes sub;
        field_one(sub, kf, kv->first);   // marshalMap always writes both key and value
        field_one(sub, vf, kv->second);
        put_tag(b, f.number, 2);
        put_varint(b, sub.size());
        b += sub;
      }
      return;
    }
    if (f.repeated) {
      if (fv.list.empty()) return;
      if (f.packed) {
        Bytes sub;
        for (auto& v : fv.list) scalar_payload(sub, f.type, v);
        put_tag(b, f.number, 2);
        put_varint(b, sub.size());
        b += sub;
      } else {
        for (auto& v : fv.list) field_one(b, f, v);
      }
      return;
    }
    if (!scalar_is_set(f, fv.list[0])) return;
    field_one(b, f, fv.list[0]);
  }

  void message(Bytes& b, const DynMsg& m) {
    std::vector<const FieldDesc*> order;
    for (auto& kv : m.known) {
      const FieldDesc* f = m.d->find_number(kv.first);
      if (f) order.push_back(f);
    }
    if (flags & ORC_F_GO_LEGACY_ORDER) {
      // [upstream internal/order LegacyFieldOrder]: non-oneof fields by number, then real-oneof
      // members grouped by oneof declaration index, by number within
      std::stable_sort(order.begin(), order.end(), [](const FieldDesc* x, const FieldDesc* y) {
        bool ox = x->oneof_index >= 0, oy = y->oneof_index >= 0;
        if (ox != oy) return !ox && oy;
        if (ox && oy && x->oneof_index != y->oneof_index) return x->oneof_index < y->oneof_index;
        return x->number < y->number;
      });
    }  // else std::map iteration order == ascending field number
    for (auto* f : order) field(b, *f, m.known.at(f->number));
  }
};

// ---------------- proto.Unmarshal ----------------
struct WireUnmarshal {
  const Schema& S;
  Err& err;
  int depth_limit = 10000;  // protowire.DefaultRecursionLimit
  WireUnmarshal(const Schema& s, Err& e) : S(s), err(e) {}

  bool bad() { return err.fail(ORC_BAD_WIRE, "proto: cannot parse invalid wire-format data"); }

  // skip one field value of wire type wt (after the tag); handles groups
  bool skip_value(const uint8_t*& p, const uint8_t* e, uint32_t num, int wt, int depth) {
    uint64_t v;
    switch (wt) {
      case 0: return get_varint(p, e, v) || bad();
      case 1: if (e - p < 8) return bad(); p += 8; return true;
      case 5: if (e - p < 4) return bad(); p += 4; return true;
      case 2:
        if (!get_varint(p, e, v)) return bad();
        if ((uint64_t)(e - p) < v) return bad();
        p += v;
        return true;
      case 3: {
        if (depth <= 0) return bad();
        while (true) {
          uint64_t tag;
          if (!get_varint(p, e, tag)) return bad();
          uint64_t n2 = tag >> 3;
          int w2 = (int)(tag & 7);
          if (n2 == 0 || n2 > 0x1FFFFFFF) return bad();
          if (w2 == 4) {
            if (n2 != num) return bad();
            return true;
          }
          if (!skip_value(p, e, (uint32_t)n2, w2, depth - 1)) return false;
        }
      }
      default: return bad();
    }
  }

  // decode one scalar of `type` with wire type wt; returns false+no error when the wire type
  // does not fit (caller treats the field as unknown) via *mismatch
  bool scalar(const uint8_t*& p, const uint8_t* e, const FieldDesc& f, int type, int wt, Val& out,
              bool& mismatch, int depth) {
    mismatch = false;
    if (wt != wire_type_of(type)) {
      mismatch = true;
      return true;
    }
    uint64_t v;
    switch (type) {
      case T_BOOL: if (!get_varint(p, e, v)) return bad(); out.u = v != 0; return true;
      case T_INT32: case T_ENUM: if (!get_varint(p, e, v)) return bad(); out.u = (uint64_t)(int64_t)(int32_t)(uint32_t)v; return true;
      case T_SINT32: if (!get_varint(p, e, v)) return bad(); out.u = (uint64_t)(int64_t)(int32_t)unzigzag64((uint64_t)(uint32_t)v); return true;
      case T_UINT32: if (!get_varint(p, e, v)) return bad(); out.u = (uint32_t)v; return true;
      case T_INT64: case T_UINT64: if (!get_varint(p, e, v)) return bad(); out.u = v; return true;
      case T_SINT64: if (!get_varint(p, e, v)) return bad(); out.u = (uint64_t)unzigzag64(v); return true;
      case T_FIXED32: case T_FLOAT: {
        if (e - p < 4) return bad();
        uint32_t x; memcpy(&x, p, 4); p += 4; out.u = x; return true;
      }
      case T_SFIXED32: {
        if (e - p < 4) return bad();
        uint32_t x; memcpy(&x, p, 4); p += 4; out.u = (uint64_t)(int64_t)(int32_t)x; return true;
      }
      case T_FIXED64: case T_SFIXED64: case T_DOUBLE: {
        if (e - p < 8) return bad();
        memcpy(&out.u, p, 8); p += 8; return true;
      }
      case T_STRING: case T_BYTES: {
        if (!get_varint(p, e, v)) return bad();
        if ((uint64_t)(e - p) < v) return bad();
        if (type == T_STRING && !utf8_valid(p, (size_t)v))
          return err.fail(ORC_INVALID_UTF8, "proto: field " + f.name + " contains invalid UTF-8");
        out.s.assign((const char*)p, (size_t)v);
        p += v;
        return true;
      }
      case T_MESSAGE: {
        if (!get_varint(p, e, v)) return bad();
        if ((uint64_t)(e - p) < v) return bad();
        if (!out.m) {
          out.m = std::make_shared<DynMsg>();
          out.m->d = &S.msgs[f.msg];
        }
        if (!message(p, p + v, *out.m, depth - 1)) return false;
        p += v;
        return true;
      }
      default: mismatch = true; return true;
    }
  }

  bool message(const uint8_t* p, const uint8_t* e, DynMsg& m, int depth) {
    if (depth < 0) return err.fail(ORC_DEPTH, "proto: exceeded max recursion depth") ;
    while (p < e) {
      uint64_t tag;
      if (!get_varint(p, e, tag)) return bad();
      uint64_t num = tag >> 3;
      int wt = (int)(tag & 7);
      if (num == 0 || num > 0x1FFFFFFF) return bad();
      if (wt == 4) return bad();  // stray end-group
      const FieldDesc* f = m.d->find_number((int32_t)num);
      if (!f) {
        if (!skip_value(p, e, (uint32_t)num, wt, depth)) return false;
        continue;
      }
      if (f->is_map) {
        if (wt != 2) {
          if (!skip_value(p, e, (uint32_t)num, wt, depth)) return false;
          continue;
        }
        uint64_t len;
        if (!get_varint(p, e, len)) return bad();
        if ((uint64_t)(e - p) < len) return bad();
        const uint8_t* q = p;
        const uint8_t* qe = p + len;
        p += len;
        const MsgDesc& ed = S.msgs[f->msg];
        const FieldDesc& kf = *ed.find_number(1);
        const FieldDesc& vf = *ed.find_number(2);
        Val key, val;
        if (vf.type == T_MESSAGE) {
          val.m = std::make_shared<DynMsg>();
          val.m->d = &S.msgs[vf.msg];
        }
        while (q < qe) {
          uint64_t t2;
          if (!get_varint(q, qe, t2)) return bad();
          uint64_t n2 = t2 >> 3;
          int w2 = (int)(t2 & 7);
          if (n2 == 0 || n2 > 0x1FFFFFFF) return bad();
          if (w2 == 4) return bad();
          bool mm = true;
          if (n2 == 1) {
            if (!scalar(q, qe, kf, kf.type, w2, key, mm, depth)) return false;
          } else if (n2 == 2) {
            if (!scalar(q, qe, vf, vf.type, w2, val, mm, depth)) return false;
          }
          if (mm && !skip_value(q, qe, (uint32_t)n2, w2, depth)) return false;
        }
        FieldVal& fv = m.known[f->number];
        bool replaced = false;
        for (auto& kv : fv.map) {
          bool same = kf.type == T_STRING ? kv.first.s == key.s : kv.first.u == key.u;
          if (same) {
            kv.second = val;
            replaced = true;
            break;
          }
        }
        if (!replaced) fv.map.push_back({key, val});
        continue;
      }
      if (f->repeated) {
        bool packable = f->type != T_STRING && f->type != T_BYTES && f->type != T_MESSAGE && f->type != T_GROUP;
        if (packable && wt == 2) {
          uint64_t len;
          if (!get_varint(p, e, len)) return bad();
          if ((uint64_t)(e - p) < len) return bad();
          const uint8_t* q = p;
          const uint8_t* qe = p + len;
          p += len;
          FieldVal& fv = m.known[f->number];
          while (q < qe) {
            Val v;
            bool mm;
            if (!scalar(q, qe, *f, f->type, wire_type_of(f->type), v, mm, depth)) return false;
            fv.list.push_back(v);
          }
          continue;
        }
        Val v;
        bool mm;
        if (!scalar(p, e, *f, f->type, wt, v, mm, depth)) return false;
        if (mm) {
          if (!skip_value(p, e, (uint32_t)num, wt, depth)) return false;
          continue;
        }
        m.known[f->number].list.push_back(v);
        continue;
      }
      // singular
      {
        // a oneof member clears its siblings
        auto it = m.known.find(f->number);
        Val v;
        if (f->type == T_MESSAGE && it != m.known.end() && !it->second.list.empty()) v = it->second.list[0];  // merge
        bool mm;
        if (!scalar(p, e, *f, f->type, wt, v, mm, depth)) return false;
        if (mm) {
          if (!skip_value(p, e, (uint32_t)num, wt, depth)) return false;
          continue;
        }
        if (f->oneof_index >= 0) {
          for (auto& g : m.d->fields)
            if (g.oneof_index == f->oneof_index && g.number != f->number) m.known.erase(g.number);
        }
        FieldVal& fv = m.known[f->number];
        fv.list.clear();
        fv.list.push_back(v);
      }
    }
    return true;
  }
};

}  // namespace orc
