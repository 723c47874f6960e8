// ggr_oracle.cc - C entry points of the CPU ORACLE (TEST INFRASTRUCTURE, see ggr_oracle.h).
//
// Pipeline restated here, step by step as the reference performs it per request:
//   request : handlePost decode            /root/reference/pkg/server/handler.go:81-88
//             ValidateRequest              /root/reference/pkg/mcp/validation.go:24-61
//             handleToolsCall              /root/reference/pkg/server/handler.go:215-271
//             ValidateToolCallParams       /root/reference/pkg/mcp/validation.go:96-125
//             InvokeMethodByTool           /root/reference/pkg/grpc/discovery.go:346-375
//             InvokeMethod (request half)  /root/reference/pkg/grpc/reflection.go:333-376
//   response: InvokeMethod (reply half)    /root/reference/pkg/grpc/reflection.go:373-391
//             result wrapping + encode     /root/reference/pkg/server/handler.go:265-270,290-297
#include "ggr_oracle.h"

#include <atomic>
#include <thread>

#include "orc_gojson.h"
#include "orc_protojson.h"

using namespace orc;

struct orc_schema {
  Schema S;
};

static uint8_t* dup_bytes(const Bytes& b, size_t* n) {
  uint8_t* p = (uint8_t*)malloc(b.size() + 1);
  memcpy(p, b.data(), b.size());
  p[b.size()] = 0;
  if (n) *n = b.size();
  return p;
}

// ---------------------------------------------------------------------------------------------
static int encode_impl(const Schema& S, int32_t msg, const uint8_t* json, size_t n, uint32_t flags, Bytes& out,
                       std::string* emsg) {
  if (msg < 0 || msg >= (int32_t)S.msgs.size()) {
    if (emsg) *emsg = "bad message index";
    return ORC_UNSUPPORTED;
  }
  DynMsg m;
  m.d = &S.msgs[msg];
  Err err;
  // reflection.go:354: "" and "{}" bypass the parser
  bool bypass = n == 0 || (n == 2 && json[0] == '{' && json[1] == '}');
  if (!bypass) {
    PJUnmarshal u(S, json, n, err);
    if (!u.run(m)) {
      if (emsg) *emsg = "proto: " + err.msg;
      return err.code ? err.code : ORC_SYNTAX;
    }
  }
  WireMarshal w(S, flags);
  if (flags & ORC_F_GRPC_FRAME) {  // [upstream grpc-go rpc_util.go msgHeader]: payload format byte, then the length
    Bytes body;
    w.message(body, m);
    const uint32_t len = (uint32_t)body.size();
    out.push_back((char)0);
    out.push_back((char)(len >> 24));
    out.push_back((char)(len >> 16));
    out.push_back((char)(len >> 8));
    out.push_back((char)len);
    out += body;
    return ORC_OK;
  }
  w.message(out, m);
  return ORC_OK;
}

static int decode_impl(const Schema& S, int32_t msg, const uint8_t* wire, size_t n, uint32_t flags, Bytes& out,
                       std::string* emsg) {
  if (msg < 0 || msg >= (int32_t)S.msgs.size()) {
    if (emsg) *emsg = "bad message index";
    return ORC_UNSUPPORTED;
  }
  if (flags & ORC_F_GRPC_FRAME) {  // [upstream grpc-go rpc_util.go parser.recvMsg]: header, then exactly `length` bytes
    if (n < 5) {
      if (emsg) *emsg = "grpc: message header truncated";
      return ORC_BAD_WIRE;
    }
    const uint32_t len = ((uint32_t)wire[1] << 24) | ((uint32_t)wire[2] << 16) | ((uint32_t)wire[3] << 8) | (uint32_t)wire[4];
    if (wire[0] == 1) {
      if (emsg) *emsg = "grpc: compressed message, no decompressor on this path";
      return ORC_UNSUPPORTED;
    }
    if (wire[0] != 0 || (size_t)len != n - 5) {
      if (emsg) *emsg = "grpc: message header does not match the payload";
      return ORC_BAD_WIRE;
    }
    return decode_impl(S, msg, wire + 5, n - 5, flags & ~ORC_F_GRPC_FRAME, out, emsg);
  }
  DynMsg m;
  m.d = &S.msgs[msg];
  Err err;
  WireUnmarshal u(S, err);
  if (!u.message(wire, wire + n, m, u.depth_limit)) {
    if (emsg) *emsg = err.msg;
    return err.code ? err.code : ORC_BAD_WIRE;
  }
  PJMarshal pm(S, flags, err);
  if (!pm.message(m)) {
    if (emsg) *emsg = "proto: " + err.msg;
    return err.code ? err.code : ORC_UNSUPPORTED;
  }
  out = std::move(pm.out);
  return ORC_OK;
}

// ---------------- mcp.SanitizeError (/root/reference/pkg/mcp/validation.go:235-271) ----------------
static std::string sanitize_error(std::string msg) {
  static const char* pats[] = {"password", "token", "key", "secret", "credential", "auth"};
  auto is_space = [](unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r'; };
  for (const char* pat : pats) {
    size_t pl = strlen(pat);
    std::string outp;
    size_t i = 0;
    while (i < msg.size()) {
      bool hit = false;
      if (i + pl <= msg.size()) {
        hit = true;
        for (size_t k = 0; k < pl; k++) {
          char c = msg[i + k];
          if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
          if (c != pat[k]) {
            hit = false;
            break;
          }
        }
      }
      if (hit) {
        i += pl;
        while (i < msg.size() && !is_space((unsigned char)msg[i])) i++;
        outp += "[REDACTED]";
      } else {
        outp.push_back(msg[i++]);
      }
    }
    msg = outp;
  }
  // SanitizeString: drop [\x00-\x1F\x7F], cap at 1024 bytes, TrimSpace
  std::string s;
  for (unsigned char c : msg)
    if (!(c < 0x20 || c == 0x7F)) s.push_back((char)c);
  if (s.size() > 1024) s.resize(1024);
  size_t a = 0, b = s.size();
  while (a < b && s[a] == ' ') a++;
  while (b > a && s[b - 1] == ' ') b--;
  return s.substr(a, b - a);
}

// ---------------- envelope writers (mcp/types.go:49-54,128-165; handler.go:290-317) ----------------
static void id_token(Bytes& out, const JVal* id) {
  if (!id || id->t == JVal::Null) out += "null";
  else if (id->t == JVal::Str) go_json_string(out, (const uint8_t*)id->s.data(), id->s.size());
  else out += format_float_go(id->n, 64);
}
static Bytes error_body(int code, const std::string& message, const Bytes& id_tok) {
  Bytes b = "{\"jsonrpc\":\"2.0\",\"error\":{\"code\":";
  b += std::to_string(code);
  b += ",\"message\":";
  go_json_string(b, (const uint8_t*)message.data(), message.size());
  b += "},\"id\":";
  b += id_tok;
  b += "}\n";
  return b;
}
static Bytes result_body(const Bytes& text, bool is_error, const Bytes& id_tok) {
  Bytes b = "{\"jsonrpc\":\"2.0\",\"result\":{\"content\":[{\"type\":\"text\"";
  if (!text.empty()) {  // `json:"text,omitempty"`
    b += ",\"text\":";
    go_json_string(b, (const uint8_t*)text.data(), text.size());
  }
  b += "}]";
  if (is_error) b += ",\"isError\":true";
  b += "},\"id\":";
  b += id_tok;
  b += "}\n";
  return b;
}

// ---------------- validators (/root/reference/pkg/mcp/validation.go) ----------------
static bool validate_depth(const JVal& v, int depth, int maxd) {
  if (depth > maxd) return false;
  if (v.t == JVal::Obj) {
    for (auto& kv : v.as_map())
      if (!validate_depth(*kv.second, depth + 1, maxd)) return false;
  } else if (v.t == JVal::Arr) {
    for (auto& c : v.a)
      if (!validate_depth(c, depth + 1, maxd)) return false;
  }
  return true;
}
static int64_t calc_size(const JVal& v) {
  switch (v.t) {
    case JVal::Str: return (int64_t)v.s.size();
    case JVal::Obj: {
      int64_t s = 0;
      for (auto& kv : v.as_map()) s += (int64_t)kv.first.size() + calc_size(*kv.second);
      return s;
    }
    case JVal::Arr: {
      int64_t s = 0;
      for (auto& c : v.a) s += calc_size(c);
      return s;
    }
    default: return 8;
  }
}
// validateParams on a map value; returns "" when fine
static std::string validate_params(const JVal& obj) {
  if (!validate_depth(obj, 0, 10)) return "object nesting too deep (max 10)";
  if (calc_size(obj) > 1024 * 1024) return "object too large (max 1048576 bytes)";
  return "";
}
static std::string validate_arguments(const JVal& a) {
  switch (a.t) {
    case JVal::Obj: return validate_params(a);
    case JVal::Arr:
      for (size_t i = 0; i < a.a.size(); i++) {
        std::string e = validate_arguments(a.a[i]);
        if (!e.empty()) return "argument[" + std::to_string(i) + "]: " + e;
      }
      return "";
    case JVal::Str:
      if (a.s.size() > 1024) return "string too long (max 1024)";
      return "";
    default: return "";
  }
}
static bool charset_ok(const Bytes& s, bool tool) {
  if (s.empty()) return false;
  for (unsigned char c : s) {
    bool ok = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' ||
              (tool ? c == '.' : c == '/');
    if (!ok) return false;
  }
  return true;
}
static bool fold_eq(const Bytes& k, const char* name) {
  size_t n = strlen(name);
  if (k.size() != n) return false;
  for (size_t i = 0; i < n; i++) {
    char c = k[i];
    if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    if (c != name[i]) return false;
  }
  return true;
}

// ---------------- batch forms ----------------
// Threads take blocks of items from a shared counter; every thread appends its outputs to an arena of its own
// (nothing allocated by one thread is freed by another) and the packed result is assembled by the same
// threads, block by block, once the offsets are known.  The all-core figure of bench.py is this code: it has
// to scale, or the GPU/CPU ratio is flattered (round-1 review: 13-25x on 128 threads).
struct Arena {
  Bytes buf;
};
struct Piece {
  uint32_t thread;
  uint64_t pos, len;
};
template <class Fn>
static void par_blocks(int64_t n, int threads, int64_t block, Fn fn) {  // fn(thread, i0, i1)
  if (threads <= 1 || n <= block) {
    fn(0, (int64_t)0, n);
    return;
  }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++)
    ts.emplace_back([&, t]() {
      while (true) {
        int64_t i0 = next.fetch_add(block);
        if (i0 >= n) break;
        fn(t, i0, i0 + block < n ? i0 + block : n);
      }
    });
  for (auto& t : ts) t.join();
}
struct BatchOut {
  std::vector<Arena> arenas;
  std::vector<Piece> pieces;
  int threads;
  BatchOut(int64_t n, int th) : arenas((size_t)(th > 1 ? th : 1)), pieces((size_t)n), threads(th > 1 ? th : 1) {}
  void put(int thread, int64_t i, const void* p, size_t len) {
    Bytes& b = arenas[(size_t)thread].buf;
    pieces[(size_t)i] = Piece{(uint32_t)thread, (uint64_t)b.size(), (uint64_t)len};
    b.append((const char*)p, len);
  }
  int pack(int64_t n, uint8_t* out, uint64_t cap, uint64_t* off) {
    uint64_t pos = 0;
    for (int64_t i = 0; i < n; i++) {
      off[i] = pos;
      pos += pieces[(size_t)i].len;
    }
    off[n] = pos;
    if (pos > cap) return ORC_NO_SPACE;
    par_blocks(n, threads, 1024, [&](int, int64_t i0, int64_t i1) {
      for (int64_t i = i0; i < i1; i++) {
        const Piece& pc = pieces[(size_t)i];
        if (pc.len) memcpy(out + off[i], arenas[pc.thread].buf.data() + pc.pos, pc.len);
      }
    });
    return ORC_OK;
  }
};


extern "C" {

orc_schema* orc_schema_new(const uint8_t* fds, size_t n, char* err, size_t errcap) {
  orc_schema* s = new orc_schema();
  SchemaBuilder b(s->S);
  if (!b.build(fds, n)) {
    set_err(err, errcap, b.err);
    delete s;
    return nullptr;
  }
  return s;
}
// naming: 0 = reflection route (full service name), 1 = FileDescriptorSet route (last package segment + service)
orc_schema* orc_schema_new2(const uint8_t* fds, size_t n, int naming, char* err, size_t errcap) {
  orc_schema* s = new orc_schema();
  SchemaBuilder b(s->S);
  b.short_service_names = naming == 1;
  if (!b.build(fds, n)) {
    set_err(err, errcap, b.err);
    delete s;
    return nullptr;
  }
  return s;
}
void orc_schema_free(orc_schema* s) { delete s; }
int32_t orc_message_index(const orc_schema* s, const char* full_name) {
  auto it = s->S.msg_by_name.find(full_name);
  return it == s->S.msg_by_name.end() ? -1 : it->second;
}
int32_t orc_method_count(const orc_schema* s) { return (int32_t)s->S.methods.size(); }
const char* orc_method_tool_name(const orc_schema* s, int32_t m) { return s->S.methods[m].tool_name.c_str(); }
const char* orc_method_path(const orc_schema* s, int32_t m) { return s->S.methods[m].path.c_str(); }
int32_t orc_method_input(const orc_schema* s, int32_t m) { return s->S.methods[m].input; }
int32_t orc_method_output(const orc_schema* s, int32_t m) { return s->S.methods[m].output; }
void orc_free(void* p) { free(p); }

int orc_encode(const orc_schema* s, int32_t msg, const uint8_t* json, size_t n, uint32_t flags, uint8_t** out,
               size_t* out_n, char* err, size_t errcap) {
  Bytes b;
  std::string em;
  int rc = encode_impl(s->S, msg, json, n, flags, b, &em);
  if (rc != ORC_OK) {
    set_err(err, errcap, em);
    *out = nullptr;
    *out_n = 0;
    return rc;
  }
  *out = dup_bytes(b, out_n);
  return ORC_OK;
}

int orc_decode(const orc_schema* s, int32_t msg, const uint8_t* wire, size_t n, uint32_t flags, uint8_t** out,
               size_t* out_n, char* err, size_t errcap) {
  Bytes b;
  std::string em;
  int rc = decode_impl(s->S, msg, wire, n, flags, b, &em);
  if (rc != ORC_OK) {
    set_err(err, errcap, em);
    *out = nullptr;
    *out_n = 0;
    return rc;
  }
  *out = dup_bytes(b, out_n);
  return ORC_OK;
}

int orc_canon_json(const uint8_t* json, size_t n, uint8_t** out, size_t* out_n, char* err, size_t errcap) {
  GoJsonParser p(json, n);
  JVal v;
  *out = nullptr;
  *out_n = 0;
  if (!p.value(v)) {
    set_err(err, errcap, "invalid JSON");
    return ORC_SYNTAX;
  }
  p.ws();
  if (p.p != p.e) {  // json.Unmarshal (unlike Decoder.Decode) rejects trailing data
    set_err(err, errcap, "invalid character after top-level value");
    return ORC_SYNTAX;
  }
  Bytes b;
  if (!go_json_marshal(b, v)) {
    set_err(err, errcap, "unsupported value");
    return ORC_INVALID_VALUE;
  }
  *out = dup_bytes(b, out_n);
  return ORC_OK;
}

int orc_format_float(double v, int bits, char* out, size_t cap) {
  std::string s = format_float_go(v, bits);
  set_err(out, cap, s);
  return (int)s.size();
}

void orc_request_out_free(orc_request_out* o) {
  free(o->args);
  free(o->wire);
  free(o->id);
  free(o->resp);
  memset(o, 0, sizeof *o);
}

int orc_request(const orc_schema* sc, const uint8_t* body, size_t n, uint32_t flags, orc_request_out* o) {
  const Schema& S = sc->S;
  memset(o, 0, sizeof *o);
  o->method = -1;
  auto finish_err = [&](int code, const std::string& msg, const Bytes& idt) {
    o->kind = 1;
    Bytes b = error_body(code, msg, idt);
    o->resp = dup_bytes(b, &o->resp_n);
    o->id = dup_bytes(idt, &o->id_n);
    return 0;
  };
  // ---- handlePost: json.NewDecoder(r.Body).Decode(&req) ----
  GoJsonParser p(body, n);
  JVal top;
  bool ok = p.value(top);  // trailing data after the first value is left unread by Decoder.Decode
  // struct decode of mcp.JSONRPCRequest
  bool type_err = false;
  bool have_jsonrpc = false, have_method = false;
  Bytes jsonrpc, method;
  bool params_set = false;  // map non-nil
  JVal params;
  params.t = JVal::Obj;
  const JVal* idv = nullptr;
  if (ok) {
    if (top.t == JVal::Obj) {
      for (auto& kv : top.mem) {
        const Bytes& k = kv.first;
        const JVal& v = kv.second;
        // exact name first, then ASCII case fold [upstream encoding/json decode.go object()]
        int which = -1;
        static const char* names[] = {"jsonrpc", "method", "params", "id"};
        for (int i = 0; i < 4; i++)
          if (k == names[i]) which = i;
        if (which < 0)
          for (int i = 0; i < 4; i++)
            if (fold_eq(k, names[i])) {
              which = i;
              break;
            }
        switch (which) {
          case 0:
            if (v.t == JVal::Str) { jsonrpc = v.s; have_jsonrpc = true; }
            else if (v.t != JVal::Null) type_err = true;
            break;
          case 1:
            if (v.t == JVal::Str) { method = v.s; have_method = true; }
            else if (v.t != JVal::Null) type_err = true;
            break;
          case 2:
            if (v.t == JVal::Obj) {
              params_set = true;
              for (auto& m : v.mem) params.mem.push_back(m);  // decoding into a non-nil map merges
            } else if (v.t == JVal::Null) {
              params_set = false;
              params.mem.clear();
            } else type_err = true;
            break;
          case 3:
            // RequestID.UnmarshalJSON: string or float64 only (types.go:25-30); null -> error
            if (v.t == JVal::Str || v.t == JVal::Num) idv = &v;
            else type_err = true;
            break;
          default: break;
        }
      }
    } else if (top.t != JVal::Null) {
      type_err = true;
    }
  }
  (void)have_jsonrpc;
  (void)have_method;
  if (!ok || type_err) return finish_err(-32700, "Parse error", "null");
  Bytes idt;
  id_token(idt, idv);
  // ---- ValidateRequest ----
  {
    std::string first;
    auto add = [&](const std::string& m) {
      if (first.empty()) first = m;
    };
    if (jsonrpc != "2.0") add("must be '2.0'");
    if (method.empty()) add("is required");
    else if (method.size() > 1024) add("must be less than 1024 characters");
    if (!method.empty() && !charset_ok(method, false)) add("contains invalid characters");
    if (!idv) add("is required");
    if (params_set) {
      std::string e = validate_params(params);
      if (!e.empty()) add(e);
    }
    if (!first.empty()) return finish_err(-32600, sanitize_error("validation errors: " + first), idt);
  }
  // ---- handleRequest ----
  if (method != "tools/call") {
    if (method == "initialize" || method == "tools/list" || method == "prompts/list" || method == "resources/list") {
      o->kind = 3;  // not the transcode path
      o->id = dup_bytes(idt, &o->id_n);
      return 0;
    }
    return finish_err(-32601, sanitize_error("method not found: " + method), idt);
  }
  // ---- handleToolsCall: ValidateToolCallParams ----
  const JVal* name = params_set ? params.get("name") : nullptr;
  const JVal* args = params_set ? params.get("arguments") : nullptr;
  {
    std::string first;
    auto add = [&](const std::string& m) {
      if (first.empty()) first = m;
    };
    if (!name) add("is required");
    else if (name->t != JVal::Str) add("must be a string");
    else if (name->s.empty()) add("cannot be empty");
    else if (name->s.size() > 128) add("must be less than 128 characters");
    else if (!charset_ok(name->s, true)) add("contains invalid characters");
    if (args) {
      std::string e = validate_arguments(*args);
      if (!e.empty()) add(e);
    }
    if (!first.empty()) {
      std::string em = "invalid parameters: validation errors: " + first;
      int code = em.find("not found") != std::string::npos ? -32601 : -32602;
      return finish_err(code, sanitize_error(em), idt);
    }
  }
  o->id = dup_bytes(idt, &o->id_n);
  // ---- json.Marshal(args) ----
  Bytes args_json;
  if (args && args->t != JVal::Null) {
    if (!go_json_marshal(args_json, *args)) return finish_err(-32603, "failed to marshal arguments", idt);
  }
  o->args = dup_bytes(args_json, &o->args_n);
  auto finish_tool_err = [&](int status, const std::string& em) {
    o->kind = 2;
    o->status = status;
    Bytes b = result_body("Error invoking method: " + sanitize_error(em), true, idt);
    o->resp = dup_bytes(b, &o->resp_n);
    return 0;
  };
  // ---- InvokeMethodByTool ----
  auto it = S.method_by_tool.find(name->s);
  if (it == S.method_by_tool.end()) return finish_tool_err(ORC_UNSUPPORTED, "tool " + name->s + " not found");
  const MethodDesc& md = S.methods[it->second];
  o->method = it->second;
  if (md.client_streaming || md.server_streaming) return finish_tool_err(ORC_UNSUPPORTED, "streaming methods are not supported");
  // ---- InvokeMethod request half ----
  Bytes wire;
  std::string em;
  int rc = encode_impl(S, md.input, (const uint8_t*)args_json.data(), args_json.size(), flags, wire, &em);
  if (rc != ORC_OK) return finish_tool_err(rc, "failed to invoke method: failed to parse input JSON: " + em);
  o->kind = 0;
  o->wire = dup_bytes(wire, &o->wire_n);
  return 0;
}

int orc_response(const orc_schema* sc, int32_t msg, const uint8_t* wire, size_t n, const uint8_t* id, size_t id_n,
                 uint32_t flags, uint8_t** out, size_t* out_n) {
  Bytes text;
  std::string em;
  Bytes idt((const char*)id, id_n);
  int rc = decode_impl(sc->S, msg, wire, n, flags, text, &em);
  Bytes body;
  if (rc == ORC_OK) body = result_body(text, false, idt);
  else if (rc == ORC_BAD_WIRE || rc == ORC_DEPTH || (rc == ORC_INVALID_UTF8 && em.rfind("proto: field", 0) == 0))
    body = result_body("Error invoking method: " + sanitize_error("failed to invoke method: gRPC call failed: " + em), true, idt);
  else
    body = result_body("Error invoking method: " + sanitize_error("failed to invoke method: failed to marshal output to JSON: " + em), true, idt);
  *out = dup_bytes(body, out_n);
  return rc;
}

int orc_encode_batch(const orc_schema* s, int64_t n, const int32_t* msg, const uint8_t* in, const uint64_t* in_off,
                     uint8_t* out, uint64_t out_cap, uint64_t* out_off, int32_t* status, uint32_t flags, int threads) {
  BatchOut B(n, threads);
  par_blocks(n, threads, 256, [&](int t, int64_t i0, int64_t i1) {
    Bytes o;
    for (int64_t i = i0; i < i1; i++) {
      o.clear();
      status[i] = encode_impl(s->S, msg[i], in + in_off[i], (size_t)(in_off[i + 1] - in_off[i]), flags, o, nullptr);
      B.put(t, i, o.data(), status[i] == ORC_OK ? o.size() : 0);
    }
  });
  return B.pack(n, out, out_cap, out_off);
}
int orc_decode_batch(const orc_schema* s, int64_t n, const int32_t* msg, const uint8_t* in, const uint64_t* in_off,
                     uint8_t* out, uint64_t out_cap, uint64_t* out_off, int32_t* status, uint32_t flags, int threads) {
  BatchOut B(n, threads);
  par_blocks(n, threads, 256, [&](int t, int64_t i0, int64_t i1) {
    Bytes o;
    for (int64_t i = i0; i < i1; i++) {
      o.clear();
      status[i] = decode_impl(s->S, msg[i], in + in_off[i], (size_t)(in_off[i + 1] - in_off[i]), flags, o, nullptr);
      B.put(t, i, o.data(), status[i] == ORC_OK ? o.size() : 0);
    }
  });
  return B.pack(n, out, out_cap, out_off);
}
int orc_request_batch(const orc_schema* s, int64_t n, const uint8_t* in, const uint64_t* in_off, uint8_t* out,
                      uint64_t out_cap, uint64_t* out_off, int32_t* method, uint8_t* ids, uint64_t ids_cap,
                      uint64_t* ids_off, int32_t* status, uint32_t flags, int threads) {
  BatchOut B(n, threads), I(n, threads);
  par_blocks(n, threads, 256, [&](int t, int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++) {
      orc_request_out o;
      orc_request(s, in + in_off[i], (size_t)(in_off[i + 1] - in_off[i]), flags, &o);
      method[i] = o.method;
      if (o.kind == 0) {
        status[i] = ORC_OK;
        B.put(t, i, o.wire, o.wire_n);
      } else {
        status[i] = o.kind == 2 ? o.status : -o.kind;
        B.put(t, i, nullptr, 0);
      }
      I.put(t, i, o.id, o.id ? o.id_n : 0);
      orc_request_out_free(&o);
    }
  });
  int rc = B.pack(n, out, out_cap, out_off);
  if (rc != ORC_OK) return rc;
  return I.pack(n, ids, ids_cap, ids_off);
}
int orc_response_batch(const orc_schema* s, int64_t n, const int32_t* msg, const uint8_t* in, const uint64_t* in_off,
                       const uint8_t* ids, const uint64_t* ids_off, uint8_t* out, uint64_t out_cap, uint64_t* out_off,
                       int32_t* status, uint32_t flags, int threads) {
  BatchOut B(n, threads);
  par_blocks(n, threads, 256, [&](int t, int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++) {
      uint8_t* o = nullptr;
      size_t on = 0;
      status[i] = orc_response(s, msg[i], in + in_off[i], (size_t)(in_off[i + 1] - in_off[i]), ids + ids_off[i],
                               (size_t)(ids_off[i + 1] - ids_off[i]), flags, &o, &on);
      B.put(t, i, o, on);
      free(o);
    }
  });
  return B.pack(n, out, out_cap, out_off);
}

}  // extern "C"
