// ggr_schema.h - host-side schema compiler: serialized FileDescriptorSet -> device table blob.
//
// Input is what the reference already has in hand at discovery time:
//   * the .binpb bytes read by /root/reference/pkg/descriptors/loader.go:33-64, or
//   * FileDescriptorProto bytes from the reflection route
//     (/root/reference/pkg/grpc/reflection.go:235-243) wrapped in a FileDescriptorSet.
// The descriptor walk mirrors the recursion-by-index of
// /root/reference/pkg/tools/builder.go:162-260 (visited set -> indices instead of pointers).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "ggr_tables.h"

namespace ggr {

enum WireOrder { ORDER_FIELD_NUMBER = 0, ORDER_GO_LEGACY = 1 };

struct MethodInfo {            // host mirror of types.MethodInfo (pkg/types/service.go:15-43)
  std::string name;            // "SayHello"
  std::string full_name;       // "hello.HelloService.SayHello"
  std::string service_name;    // "hello.HelloService"
  std::string tool_name;       // "hello_helloservice_sayhello" (GenerateToolName, service.go:53-61)
  std::string grpc_path;       // "/hello.HelloService/SayHello" (reflection.go:367)
  std::string input_type, output_type;
  int32_t input_msg = -1, output_msg = -1;
  bool client_streaming = false, server_streaming = false;
};

struct CompiledSchema {
  std::vector<uint8_t> blob;                  // GgrSchemaHdr + tables, 16-byte aligned sections
  std::map<std::string, int32_t> msg_index;   // full name -> message index
  std::vector<std::string> msg_names;
  std::vector<MethodInfo> methods;
  std::map<std::string, int32_t> tool_index;  // tool name -> method index
  uint32_t max_msg_fields = 0;
};

// Returns false and fills *err on malformed or unresolvable input.
// short_service_names: the FileDescriptorSet route's tool names (pkg/descriptors/loader.go:221-235)
bool compile_schema(const uint8_t* fds, size_t n, WireOrder order, CompiledSchema* out, std::string* err, bool short_service_names = false);

uint32_t key_hash(const uint8_t* p, size_t n);

}  // namespace ggr
