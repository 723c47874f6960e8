/*
 * ggr_oracle.h - C interface of the CPU ORACLE for the ggRMCP tools/call transcode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in ggrmcp_b200/ (the product) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / the timed CPU baseline.
 *
 * The oracle restates, in plain DOM-style C++, what the reference does per request
 * (/root/reference/pkg/server/handler.go:81-139,215-271,290-297 and
 *  /root/reference/pkg/grpc/reflection.go:333-391), including the third-party pieces the
 * reference calls and does not vendor (google.golang.org/protobuf v1.36.6 protojson/dynamicpb/
 * proto, Go 1.23 encoding/json + strconv; go.mod:3,11-12).  See oracle/README.md for the
 * pinning status of each function.
 */
#ifndef GGR_ORACLE_H_
#define GGR_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-item status categories (shared numbering with include/ggrmcp_b200.h). */
enum {
  ORC_OK = 0,
  ORC_SYNTAX = 1,         /* JSON tokenizer error / unexpected token            */
  ORC_UNKNOWN_FIELD = 2,  /* protojson: unknown field "x"                       */
  ORC_INVALID_VALUE = 3,  /* protojson: invalid value for <kind> field          */
  ORC_RANGE = 4,          /* well-known-type value out of range                 */
  ORC_INVALID_UTF8 = 5,
  ORC_DUPLICATE = 6,      /* duplicate field / duplicate map key                */
  ORC_ONEOF = 7,          /* oneof already set                                  */
  ORC_DEPTH = 8,
  ORC_TOO_LARGE = 9,
  ORC_BAD_WIRE = 10,      /* proto.Unmarshal: cannot parse invalid wire-format  */
  ORC_UNSUPPORTED = 11,
  ORC_NO_SPACE = 12
};

/* flags */
#define ORC_F_COMMA_SPACE 0x1u   /* protojson detrand bit: ", " instead of ","          */
#define ORC_F_GO_LEGACY_ORDER 0x2u /* wire field order = Go order.LegacyFieldOrder      */
#define ORC_F_GRPC_FRAME 0x4u      /* grpc-go rpc_util.go msgHeader: 1 byte compressed flag + big-endian uint32 length in front of
                                      the request wire; replies arrive framed (what conn.Invoke adds / strips, reflection.go:367-376) */

typedef struct orc_schema orc_schema;

orc_schema* orc_schema_new(const uint8_t* fds, size_t n, char* err, size_t errcap);
/* naming: 0 reflection route (full service name), 1 FileDescriptorSet route (pkg/descriptors/loader.go:221-235) */
orc_schema* orc_schema_new2(const uint8_t* fds, size_t n, int naming, char* err, size_t errcap);
void orc_schema_free(orc_schema*);
int32_t orc_message_index(const orc_schema*, const char* full_name);
int32_t orc_method_count(const orc_schema*);
/* tool name as types.GenerateToolName builds it (pkg/types/service.go:53-61), reflection route */
const char* orc_method_tool_name(const orc_schema*, int32_t method);
const char* orc_method_path(const orc_schema*, int32_t method); /* "/pkg.Svc/Method" */
int32_t orc_method_input(const orc_schema*, int32_t method);
int32_t orc_method_output(const orc_schema*, int32_t method);

void orc_free(void* p);

/* A5+A6: protojson.Unmarshal(json) into a dynamic message, then proto.Marshal (canonical order). */
int orc_encode(const orc_schema*, int32_t msg, const uint8_t* json, size_t n, uint32_t flags,
               uint8_t** out, size_t* out_n, char* err, size_t errcap);
/* A7+A8: proto.Unmarshal(wire) into a dynamic message, then protojson.Marshal. */
int orc_decode(const orc_schema*, int32_t msg, const uint8_t* wire, size_t n, uint32_t flags,
               uint8_t** out, size_t* out_n, char* err, size_t errcap);
/* A1+A2 on a bare JSON value: encoding/json decode into interface{} then json.Marshal. */
int orc_canon_json(const uint8_t* json, size_t n, uint8_t** out, size_t* out_n, char* err,
                   size_t errcap);

/*
 * Whole request side (A1,A3,A2,A4,A5,A6): JSON-RPC body -> tool + canonical args + wire.
 * kind: 0 = invoke (wire valid), 1 = JSON-RPC error response (resp holds the full body),
 *       2 = tool-call error result (isError:true; resp holds the full body)
 */
typedef struct {
  int32_t kind;
  int32_t status;      /* ORC_* of the transcode when kind==2 */
  int32_t method;      /* index of the resolved method, -1 if none */
  uint8_t* args;       /* canonical arguments string (A2) */
  size_t args_n;
  uint8_t* wire;       /* request wire bytes (A6) */
  size_t wire_n;
  uint8_t* id;         /* id re-printed as encoding/json would (e.g. 2 or "abc") */
  size_t id_n;
  uint8_t* resp;       /* full HTTP body for kind 1/2 (ends with '\n') */
  size_t resp_n;
} orc_request_out;
int orc_request(const orc_schema*, const uint8_t* body, size_t n, uint32_t flags,
                orc_request_out* out);
void orc_request_out_free(orc_request_out*);

/* Whole response side (A7,A8,A10): reply wire + id token -> HTTP body. Returns ORC_* status;
 * on a marshal error the body is the isError:true variant. */
int orc_response(const orc_schema*, int32_t msg, const uint8_t* wire, size_t n,
                 const uint8_t* id, size_t id_n, uint32_t flags, uint8_t** out, size_t* out_n);

/* Batch forms used to time the CPU baseline (threads >= 1).  in_off/out_off have n+1 entries. */
int orc_encode_batch(const orc_schema*, int64_t n, const int32_t* msg, const uint8_t* in,
                     const uint64_t* in_off, uint8_t* out, uint64_t out_cap, uint64_t* out_off,
                     int32_t* status, uint32_t flags, int threads);
int orc_decode_batch(const orc_schema*, int64_t n, const int32_t* msg, const uint8_t* in,
                     const uint64_t* in_off, uint8_t* out, uint64_t out_cap, uint64_t* out_off,
                     int32_t* status, uint32_t flags, int threads);
/* request bodies -> wire (envelope included) and reply wire + ids -> response bodies */
int orc_request_batch(const orc_schema*, int64_t n, const uint8_t* in, const uint64_t* in_off,
                      uint8_t* out, uint64_t out_cap, uint64_t* out_off, int32_t* method,
                      uint8_t* ids, uint64_t ids_cap, uint64_t* ids_off, int32_t* status,
                      uint32_t flags, int threads);
int orc_response_batch(const orc_schema*, int64_t n, const int32_t* msg, const uint8_t* in,
                       const uint64_t* in_off, const uint8_t* ids, const uint64_t* ids_off,
                       uint8_t* out, uint64_t out_cap, uint64_t* out_off, int32_t* status,
                       uint32_t flags, int threads);

/* number formatting helpers exposed for the float tests */
int orc_format_float(double v, int bits, char* out, size_t cap); /* Go json/protojson style */

#ifdef __cplusplus
}
#endif
#endif
