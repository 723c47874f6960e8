// ggr_tables.h - flat, pointer-free descriptor tables shared by the host-side schema compiler
// and the device kernels.  One blob per registered schema, uploaded once to HBM and read through
// the read-only path by every thread.
//
// These tables play the role protoreflect.MessageDescriptor plays for the reference
// (types.MethodInfo.InputDescriptor / OutputDescriptor, /root/reference/pkg/types/service.go:29-30):
// everything protojson/dynamicpb look up reflectively per field is resolved here at
// registration time (tags, wire types, packedness, presence, oneof membership, JSON key text,
// key -> field hash tables, enum name <-> number tables).
#pragma once
#include <stdint.h>

#define GGR_SCHEMA_MAGIC 0x47475231u /* "GGR1" */

// FieldDescriptorProto.Type values (descriptor.proto)
enum {
  GK_DOUBLE = 1, GK_FLOAT = 2, GK_INT64 = 3, GK_UINT64 = 4, GK_INT32 = 5, GK_FIXED64 = 6,
  GK_FIXED32 = 7, GK_BOOL = 8, GK_STRING = 9, GK_GROUP = 10, GK_MESSAGE = 11, GK_BYTES = 12,
  GK_UINT32 = 13, GK_ENUM = 14, GK_SFIXED32 = 15, GK_SFIXED64 = 16, GK_SINT32 = 17, GK_SINT64 = 18
};

enum { GGR_WKT_NONE = 0, GGR_WKT_TIMESTAMP = 1, GGR_WKT_DURATION = 2, GGR_WKT_WRAPPER = 3, GGR_WKT_EMPTY = 4, GGR_WKT_FIELDMASK = 5, GGR_WKT_UNSUPPORTED = 99 };

// GgrField.flags
#define GF_REPEATED 0x01u
#define GF_PACKED 0x02u
#define GF_MAP 0x04u
#define GF_PRESENCE 0x08u   /* explicit presence: message, oneof member, proto3 optional */
#define GF_PACKABLE 0x10u   /* repeated scalar that may arrive packed on the wire */

// GgrMsg.flags
#define GM_MAP_ENTRY 0x01u
#define GM_DECL_IS_EMIT 0x02u /* declaration order == wire emit order */

struct GgrSchemaHdr {  // 64 bytes
  uint32_t magic;
  uint32_t total_bytes;
  uint32_t n_msgs, msgs_off;       // GgrMsg[]
  uint32_t n_fields, fields_off;   // GgrField[]; a message's fields are contiguous, in EMIT order
  uint32_t n_enums, enums_off;     // GgrEnum[]
  uint32_t n_evals, evals_off;     // GgrEnumValue[]; an enum's values are contiguous, sorted by number
  uint32_t n_hash, hash_off;       // GgrHashEnt[] (all key tables and enum-name tables)
  uint32_t n_u16, u16_off;         // uint16_t[] (declaration-order permutations, number LUTs)
  uint32_t pool_bytes, pool_off;   // byte pool: key text, `"jsonName":` text, enum names
};

struct GgrToolsTrailer {  // the last 16 bytes of the blob (blob + total_bytes - 16)
  uint32_t hash_first, hash_mask;  // tool name -> method index (GgrHashEnt table)
  uint32_t methods_first;          // u16[methods_first + m] = input message of method m, 0xFFFF = streaming
  uint32_t n_methods;
};

struct GgrMsg {  // 32 bytes
  uint32_t field_first;   // index of the first GgrField
  uint16_t n_fields;
  uint8_t wkt;
  uint8_t flags;
  uint32_t key_hash_first;  // index of the first GgrHashEnt of the JSON-key table
  uint32_t key_hash_mask;   // table size - 1 (power of two)
  uint32_t decl_first;      // u16[decl_first + d] = emit index of the d-th declared field
  uint32_t lut_first;       // u16[lut_first + number] = emit index + 1 (0 = none) for number < lut_n
  uint32_t lut_n;
  uint32_t n_oneofs;
};

struct GgrField {  // 32 bytes
  uint32_t number;
  uint32_t tag;        // (number << 3) | wire type used when this field is written (LEN when packed)
  uint8_t tag_len;     // varint length of tag
  uint8_t kind;        // GK_*
  uint8_t flags;       // GF_*
  uint8_t wt;          // element wire type (0 varint, 1 fixed64, 2 len, 5 fixed32)
  int16_t oneof;       // real oneof index, -1 if none
  uint16_t decl_index; // declaration index within the message
  int32_t child;       // message index (GK_MESSAGE) / enum index (GK_ENUM) / -1
  uint32_t name_off;   // pool offset of the text `"jsonName":`
  uint16_t name_len;
  uint16_t pad0;
  uint32_t pad1;
};

struct GgrEnum {  // 16 bytes
  uint32_t val_first;
  uint32_t n_vals;
  uint32_t hash_first;  // name -> number table
  uint32_t hash_mask;
};

struct GgrEnumValue {  // 16 bytes; per enum sorted by number, one entry per distinct number
  int32_t number;      // (the first declared name of each number: protoreflect ByNumber)
  uint32_t name_off;
  uint32_t name_len;
  uint32_t pad;
};

struct GgrHashEnt {  // 32 bytes; name_len == 0xFFFFFFFF marks an empty slot
  uint32_t hash;      // ggr::key_hash of the name bytes (word-wise, see ggr_json_in.cuh)
  uint32_t name_off;
  uint32_t name_len;
  int32_t value;      // key tables: emit index of the field; enum tables: the enum number
  uint32_t w[4];      // first 16 bytes of the name, zero padded (hits need no pool access)
};

// batch flags the kernels look at (mirror include/ggrmcp_b200.h)
#define GGR_DF_COMMA_SPACE 0x1u
#define GGR_DF_GRPC_FRAME 0x2u
#define GGR_FRAME_BYTES 5u

// per-item status (mirrors ggr_status in include/ggrmcp_b200.h)
enum {
  GST_OK = 0, GST_SYNTAX = 1, GST_UNKNOWN_FIELD = 2, GST_INVALID_VALUE = 3, GST_RANGE = 4,
  GST_INVALID_UTF8 = 5, GST_DUPLICATE = 6, GST_ONEOF = 7, GST_DEPTH = 8, GST_TOO_LARGE = 9,
  GST_BAD_WIRE = 10, GST_UNSUPPORTED = 11, GST_NO_SPACE = 12,
  GST_INTERNAL = 13  /* size pass and write pass disagreed: engine bug, never expected */
};
