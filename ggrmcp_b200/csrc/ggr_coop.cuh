// ggr_coop.cuh - warp-cooperative reply side: one warp per item, lanes work on different fields.
//
// The per-thread walker (ggr_decode.cuh) keeps one lane busy per message and the other 31 mostly
// idle on divergent paths.  Here the warp first lays the message out as a table of field entries
// in shared memory and then processes entries, not messages:
//   P1 discover : level by level (root, its sub-messages, theirs, ...) one lane scans one message's
//                 top-level tags - length-delimited payloads are skipped in O(1) - and appends
//                 one entry per field, linked to its parent message
//   P2 size     : one lane per entry computes the JSON size of leaf values (string escape scan,
//                 digit counts, ...): the bulk of the byte work, evenly spread
//   P3 totals   : bottom-up, one lane per message adds up its children
//   P4 offsets  : top-down, one lane per message hands out output offsets to its children
//   P5 write    : one lane per entry writes its text at its offset (write kernel only)
// Only the regular case is handled (fields in declaration order, no maps, table fits); anything
// else - including every malformed item - is left to the general per-thread kernels, which are
// the reference for semantics.  The code is written lane-by-lane against a small shared-state
// struct so that the host simulation can run the same phases with the lanes in sequence.
#pragma once
#include "ggr_decode.cuh"

#define GGR_COOP_ENTRIES 320
#define GGR_COOP_LEVELS 24
#define GGR_MODE_COOP 2u
#define GGR_MODE_PENDING 0xFFu

#define CF_MSG 0x0001u         /* entry is a (non-WKT) sub-message: has children */
#define CF_ARR_FIRST 0x0002u   /* first element of a repeated field: writes name and '[' */
#define CF_ARR_LAST 0x0004u    /* last element: writes ']' */
#define CF_ARR_ELEM 0x0008u    /* element of a repeated field */
#define CF_PACKED 0x0010u      /* packed repeated scalars: the entry holds the whole run */
#define CF_FIRST 0x0020u       /* first field written inside its parent: no leading comma */
#define CF_TIMESTAMP 0x0040u   /* google.protobuf.Timestamp leaf */

struct CoopEnt {            // 32 bytes
  u32 vpos;                 // position right after the tag (length prefix / scalar bytes start)
  u32 vend;                 // end of the value
  u32 gfield;               // global field index (0xFFFFFFFF for the root)
  u16 parent, next;         // parent message entry; next sibling (0xFFFF = none)
  u32 size;                 // full text size: separator + name + brackets + value
  u32 off;                  // output offset of the full text
  u16 first_child, last_child;
  u16 flags, level;
  u32 msg;                  // CF_MSG: message type index
};

struct CoopShared {
  CoopEnt ent[GGR_COOP_ENTRIES];
  u32 n_ent;
  u32 bail;                 // nonzero: leave the item to the general kernels
  u32 level_beg[GGR_COOP_LEVELS + 2];
};

GGR_DEV u32 coop_alloc(CoopShared& S) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(&S.n_ent, 1u);
#else
  return S.n_ent++;
#endif
}

// P1, one lane: scan the top-level fields of message entry `me` and append its children.
GGR_DEV void coop_scan_message(CoopShared& S, const DecCtx& cx, u32 me) {
  const Tables& T = cx.T;
  CoopEnt m = S.ent[me];
  MsgD md = ggr_msg(T, m.msg);
  if (md.wkt != GGR_WKT_NONE) { S.bail = 1; return; }
  Rd r;
  r.init(cx.in, m.vpos, m.vend);
  if (m.gfield != 0xFFFFFFFFu) {  // a field entry starts at its length prefix
    u64 len;
    if (!rd_varint(r, m.vend, &len) || len != (u64)(m.vend - r.pos)) { S.bail = 1; return; }
  }
  i32 last_decl = -1;
  u32 open = 0;        // emit index + 1 of the repeated field currently being collected
  u32 oneofs = 0;
  u32 prev = 0xFFFFu;  // previous child entry
  bool any = false;
  while (r.pos < m.vend) {
    u64 tag;
    if (!rd_varint(r, m.vend, &tag)) { S.bail = 1; return; }
    u64 num64 = tag >> 3;
    u32 wt = (u32)(tag & 7);
    if (num64 == 0 || num64 > 0x1FFFFFFFull || wt == 4 || wt > 5) { S.bail = 1; return; }
    u32 num = (u32)num64;
    i32 ei = find_field(T, md, num);
    if (ei < 0) {
      if (!rd_skip_value(r, m.vend, num, wt)) { S.bail = 1; return; }
      continue;
    }
    FieldD f = ggr_field(T, md.field_first + (u32)ei);
    bool packed_in = (f.flags & GF_PACKABLE) && wt == 2;
    if (wt != f.wt && !packed_in) {
      if (!rd_skip_value(r, m.vend, num, wt)) { S.bail = 1; return; }
      continue;
    }
    if (f.flags & GF_MAP) { S.bail = 1; return; }
    u32 vpos = r.pos;
    u32 flags = 0;
    bool repeated = (f.flags & GF_REPEATED) != 0;
    if (repeated) {
      flags |= CF_ARR_ELEM;
      if (packed_in) {  // an empty packed run contributes nothing (and opens nothing)
        Rd t = r;
        u64 len;
        if (!rd_varint(t, m.vend, &len) || len > (u64)(m.vend - t.pos)) { S.bail = 1; return; }
        if (len == 0) {
          if (open != (u32)ei + 1 && (i32)f.decl_index <= last_decl) { S.bail = 1; return; }
          r = t;
          continue;
        }
      }
      if (open != (u32)ei + 1) {
        if ((i32)f.decl_index <= last_decl) { S.bail = 1; return; }
        if (open && prev != 0xFFFFu) S.ent[prev].flags |= CF_ARR_LAST;
        last_decl = (i32)f.decl_index;
        open = (u32)ei + 1;
        flags |= CF_ARR_FIRST;
      }
      if (packed_in) flags |= CF_PACKED;
    } else {
      if (open && prev != 0xFFFFu) S.ent[prev].flags |= CF_ARR_LAST;
      open = 0;
      if ((i32)f.decl_index <= last_decl) { S.bail = 1; return; }
      if (f.oneof >= 0) {
        u32 bit = 1u << (f.oneof & 31);
        if (oneofs & bit) { S.bail = 1; return; }
        oneofs |= bit;
      }
      last_decl = (i32)f.decl_index;
      if (f.kind != GK_MESSAGE && !(f.flags & GF_PRESENCE)) {
        // implicit presence: zero values are not written
        Rd t = r;
        bool z;
        if (f.wt == 2) {
          u64 len;
          if (!rd_varint(t, m.vend, &len) || len > (u64)(m.vend - t.pos)) { S.bail = 1; return; }
          z = len == 0;
        } else {
          Cnt c;
          c.pos = 0;
          if (scalar_value<Cnt, false>(c, cx, t, m.vend, f.kind, f.child, false, &z) != GST_OK) { S.bail = 1; return; }
        }
        if (z) {
          if (!rd_skip_value(r, m.vend, num, wt)) { S.bail = 1; return; }
          continue;
        }
      }
    }
    if (!rd_skip_value(r, m.vend, num, wt)) { S.bail = 1; return; }
    u32 slot = coop_alloc(S);
    if (slot >= GGR_COOP_ENTRIES) { S.bail = 1; return; }
    CoopEnt e;
    e.vpos = vpos;
    e.vend = r.pos;
    e.gfield = md.field_first + (u32)ei;
    e.parent = (u16)me;
    e.next = 0xFFFFu;
    e.size = 0;
    e.off = 0;
    e.first_child = e.last_child = 0xFFFFu;
    e.level = (u16)(m.level + 1);
    e.msg = 0;
    if (!any) flags |= CF_FIRST;
    if (f.kind == GK_MESSAGE) {
      MsgD cd = ggr_msg(T, (u32)f.child);
      if (cd.wkt == GGR_WKT_TIMESTAMP) flags |= CF_TIMESTAMP;
      else if (cd.wkt != GGR_WKT_NONE) { S.bail = 1; return; }
      else {
        flags |= CF_MSG;
        e.msg = (u32)f.child;
      }
    }
    e.flags = (u16)flags;
    S.ent[slot] = e;
    if (prev == 0xFFFFu) S.ent[me].first_child = (u16)slot;
    else S.ent[prev].next = (u16)slot;
    prev = slot;
    any = true;
  }
  if (r.pos != m.vend) { S.bail = 1; return; }
  if (open && prev != 0xFFFFu) S.ent[prev].flags |= CF_ARR_LAST;
  S.ent[me].last_child = (u16)prev;
}

// Text that surrounds an entry's value: separator, key, brackets.
GGR_DEV u32 coop_prefix_len(const DecCtx& cx, const CoopEnt& e, u32 name_len) {
  u32 n = 0;
  // a comma precedes everything except the first entry written inside its parent message
  // (CF_FIRST is only ever set on a message's first child, which for an array is its first element)
  if (!(e.flags & CF_FIRST)) n += (cx.flags & GGR_F_COMMA_SPACE) ? 2u : 1u;
  if (!(e.flags & CF_ARR_ELEM) || (e.flags & CF_ARR_FIRST)) n += name_len;
  if (e.flags & CF_ARR_FIRST) n += 1;
  return n;
}
template <class W>
GGR_DEV void coop_put_prefix(W& w, const DecCtx& cx, const CoopEnt& e, const FieldD& f) {
  if (!(e.flags & CF_FIRST)) {
    w.put1(',');
    if (cx.flags & GGR_F_COMMA_SPACE) w.put1(' ');
  }
  if (!(e.flags & CF_ARR_ELEM) || (e.flags & CF_ARR_FIRST)) put_pool(w, cx.T.pool, f.name_off, f.name_len);
  if (e.flags & CF_ARR_FIRST) w.put1('[');
}

// Value text of a leaf entry (scalar, packed run, timestamp) into any writer.
template <class W>
GGR_DEV int coop_leaf_value(W& w, const DecCtx& cx, const CoopEnt& e, const FieldD& f) {
  Rd r;
  r.init(cx.in, e.vpos, e.vend);
  if (e.flags & CF_TIMESTAMP) {
    u64 len;
    if (!rd_varint(r, e.vend, &len) || len > (u64)(e.vend - r.pos)) return GST_BAD_WIRE;
    i64 s, n;
    int st = read_timestamp_payload(r, r.pos + (u32)len, &s, &n);
    if (st != GST_OK) return st;
    return put_timestamp(w, s, n);
  }
  if (e.flags & CF_PACKED) {
    u64 len;
    if (!rd_varint(r, e.vend, &len) || len > (u64)(e.vend - r.pos)) return GST_BAD_WIRE;
    u32 lim = r.pos + (u32)len;
    u32 first = 1;
    while (r.pos < lim) {
      put_sep(w, cx, first);
      bool z;
      int st = scalar_value<W, true>(w, cx, r, lim, f.kind, f.child, false, &z);
      if (st != GST_OK) return st;
    }
    return r.pos == lim ? GST_OK : GST_BAD_WIRE;
  }
  bool z;
  return scalar_value<W, true>(w, cx, r, e.vend, f.kind, f.child, false, &z);
}

// P2, one lane: size of entry e (leaf: full text; message: prefix/suffix only, children added in P3)
GGR_DEV void coop_size_entry(CoopShared& S, const DecCtx& cx, u32 ei) {
  CoopEnt e = S.ent[ei];
  if (e.gfield == 0xFFFFFFFFu) {  // root: "{" ... "}"
    S.ent[ei].size = 2;
    return;
  }
  FieldD f = ggr_field(cx.T, e.gfield);
  u32 n = coop_prefix_len(cx, e, f.name_len);
  if (e.flags & CF_ARR_LAST) n += 1;
  if (e.flags & CF_MSG) {
    n += 2;
  } else {
    Cnt c;
    c.pos = 0;
    if (coop_leaf_value(c, cx, e, f) != GST_OK) {
      S.bail = 1;
      return;
    }
    n += c.pos;
  }
  S.ent[ei].size = n;
}

// P3, one lane: message entry adds the sizes of its children (children are complete: deeper level)
GGR_DEV void coop_total_message(CoopShared& S, u32 me) {
  u32 sum = 0;
  for (u32 c = S.ent[me].first_child; c != 0xFFFFu; c = S.ent[c].next) sum += S.ent[c].size;
  S.ent[me].size += sum;
}

// P4, one lane: message entry hands out offsets to its children
GGR_DEV void coop_offsets_message(CoopShared& S, const DecCtx& cx, u32 me) {
  CoopEnt m = S.ent[me];
  u32 pos = m.off;
  if (m.gfield != 0xFFFFFFFFu) {
    FieldD f = ggr_field(cx.T, m.gfield);
    pos += coop_prefix_len(cx, m, f.name_len);
  }
  pos += 1;  // '{'
  for (u32 c = m.first_child; c != 0xFFFFu; c = S.ent[c].next) {
    S.ent[c].off = pos;
    pos += S.ent[c].size;
  }
}

// P5, one lane: write entry e
GGR_DEV int coop_write_entry(CoopShared& S, const DecCtx& cx, u32 ei, u8* out) {
  CoopEnt e = S.ent[ei];
  Wr w;
  if (e.gfield == 0xFFFFFFFFu) {
    w.init(out, e.off);
    w.put1('{');
    w.finish();
    w.init(out, e.off + e.size - 1);
    w.put1('}');
    w.finish();
    return GST_OK;
  }
  FieldD f = ggr_field(cx.T, e.gfield);
  w.init(out, e.off);
  coop_put_prefix(w, cx, e, f);
  if (e.flags & CF_MSG) {
    w.put1('{');
    w.finish();
    u32 tail = (e.flags & CF_ARR_LAST) ? 2u : 1u;
    w.init(out, e.off + e.size - tail);
    w.put1('}');
    if (e.flags & CF_ARR_LAST) w.put1(']');
    w.finish();
    return GST_OK;
  }
  int st = coop_leaf_value(w, cx, e, f);
  if (e.flags & CF_ARR_LAST) w.put1(']');
  w.finish();
  if (st == GST_OK && w.pos != e.off + e.size) st = GST_INTERNAL;
  return st;
}

// The whole item, driven by `nlanes` lanes; `lane` is this lane's index.  On the device all 32
// lanes of a warp call this together (SYNC = __syncwarp); the host simulation calls the phase
// helpers lane after lane instead (see coop_run_host below).
#if defined(__CUDA_ARCH__)
#define GGR_COOP_SYNC() __syncwarp()
#else
#define GGR_COOP_SYNC() do { } while (0)
#endif

// Returns true when the item was handled; *size gets the output size.  With out != nullptr the
// text is also written at out[out_off ..).
GGR_DEV bool coop_decode_item(CoopShared& S, const DecCtx& cx, u32 root_msg, u32 start, u32 end, u32 lane, u32 nlanes,
                              u8* out, u32 out_off, u32* size, int* wstatus) {
  if (lane == 0) {
    S.n_ent = 1;
    S.bail = 0;
    CoopEnt r0;
    r0.vpos = start; r0.vend = end; r0.gfield = 0xFFFFFFFFu; r0.parent = 0xFFFFu; r0.next = 0xFFFFu;
    r0.size = 0; r0.off = out_off; r0.first_child = r0.last_child = 0xFFFFu; r0.flags = CF_MSG | CF_FIRST; r0.level = 0;
    r0.msg = root_msg;
    S.ent[0] = r0;
    S.level_beg[0] = 0;
    S.level_beg[1] = 1;
  }
  GGR_COOP_SYNC();
  // P1
  u32 nlev = 0;
  for (;;) {
    u32 b = S.level_beg[nlev], e = S.level_beg[nlev + 1];
    if (b == e) break;
    for (u32 i = b + lane; i < e; i += nlanes)
      if (S.ent[i].flags & CF_MSG) coop_scan_message(S, cx, i);
    GGR_COOP_SYNC();
    if (S.bail) return false;
    nlev++;
    if (lane == 0) S.level_beg[nlev + 1] = S.n_ent < GGR_COOP_ENTRIES ? S.n_ent : GGR_COOP_ENTRIES;
    GGR_COOP_SYNC();
    if (nlev >= GGR_COOP_LEVELS) return false;
  }
  u32 n = S.n_ent;
  if (n > GGR_COOP_ENTRIES) return false;
  // P2
  for (u32 i = lane; i < n; i += nlanes) coop_size_entry(S, cx, i);
  GGR_COOP_SYNC();
  if (S.bail) return false;
  // P3 bottom-up
  for (i32 l = (i32)nlev - 1; l >= 0; l--) {
    u32 b = S.level_beg[l], e = S.level_beg[l + 1];
    for (u32 i = b + lane; i < e; i += nlanes)
      if (S.ent[i].flags & CF_MSG) coop_total_message(S, i);
    GGR_COOP_SYNC();
  }
  *size = S.ent[0].size;
  if (!out) return true;
  // P4 top-down
  for (u32 l = 0; l < nlev; l++) {
    u32 b = S.level_beg[l], e = S.level_beg[l + 1];
    for (u32 i = b + lane; i < e; i += nlanes)
      if (S.ent[i].flags & CF_MSG) coop_offsets_message(S, cx, i);
    GGR_COOP_SYNC();
  }
  // P5
  int st = GST_OK;
  for (u32 i = lane; i < n; i += nlanes) {
    int s2 = coop_write_entry(S, cx, i, out);
    if (s2 != GST_OK) st = s2;
  }
  *wstatus = st;
  return true;
}
