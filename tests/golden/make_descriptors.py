#!/usr/bin/env python3
"""Build the FileDescriptorSet fixtures used by the oracle, the engine and the tests.

There is no protoc in the build image, so the descriptors are assembled by hand with
python-protobuf's descriptor_pb2.  They restate (not copy) the schemas the reference ships:

  * hello.proto     - /root/reference/examples/hello-service/proto/hello.proto:8-25
  * complex.proto   - /root/reference/tests/testdata/complex.proto:1-97
  * bench.proto     - synthetic descriptors named in SURVEY.md section 8(d)
                      (bench.Flat, bench.Blob, bench.All, bench.Echo services)

Output: tests/golden/schemas.binpb (one serialized google.protobuf.FileDescriptorSet).
Run:    python tests/golden/make_descriptors.py
"""
import os
import sys

from google.protobuf import descriptor_pb2 as dpb
from google.protobuf import timestamp_pb2, duration_pb2, wrappers_pb2, empty_pb2, struct_pb2, field_mask_pb2

F = dpb.FieldDescriptorProto

TYPES = {
    "double": F.TYPE_DOUBLE, "float": F.TYPE_FLOAT, "int64": F.TYPE_INT64, "uint64": F.TYPE_UINT64,
    "int32": F.TYPE_INT32, "fixed64": F.TYPE_FIXED64, "fixed32": F.TYPE_FIXED32, "bool": F.TYPE_BOOL,
    "string": F.TYPE_STRING, "bytes": F.TYPE_BYTES, "uint32": F.TYPE_UINT32, "sfixed32": F.TYPE_SFIXED32,
    "sfixed64": F.TYPE_SFIXED64, "sint32": F.TYPE_SINT32, "sint64": F.TYPE_SINT64,
}


def camel(name):
    """protoc's default json_name: drop '_' and upper-case the following letter."""
    out, up = [], False
    for ch in name:
        if ch == "_":
            up = True
        elif up:
            out.append(ch.upper())
            up = False
        else:
            out.append(ch)
    return "".join(out)


def add_field(msg, name, number, typ, *, repeated=False, oneof=None, json_name=None,
              packed=None, proto3_optional=False):
    f = msg.field.add()
    f.name = name
    f.number = number
    f.label = F.LABEL_REPEATED if repeated else F.LABEL_OPTIONAL
    if typ in TYPES:
        f.type = TYPES[typ]
    elif typ.startswith("enum:"):
        f.type = F.TYPE_ENUM
        f.type_name = typ[5:]
    else:
        f.type = F.TYPE_MESSAGE
        f.type_name = typ
    f.json_name = json_name if json_name is not None else camel(name)
    if oneof is not None:
        f.oneof_index = oneof
    if packed is not None:
        f.options.packed = packed
    if proto3_optional:
        f.proto3_optional = True
    return f


def add_map(msg, pkg_msg_fqn, name, number, ktype, vtype):
    """map<k,v> name = number;  ->  nested XxxEntry message with map_entry option."""
    entry_name = camel("_" + name) + "Entry"  # protoc: CamelCase(name) + "Entry"
    entry_name = entry_name[0].upper() + entry_name[1:]
    e = msg.nested_type.add()
    e.name = entry_name
    e.options.map_entry = True
    add_field(e, "key", 1, ktype)
    add_field(e, "value", 2, vtype)
    add_field(msg, name, number, pkg_msg_fqn + "." + entry_name, repeated=True)


def add_method(svc, name, inp, out):
    m = svc.method.add()
    m.name = name
    m.input_type = inp
    m.output_type = out


def hello_file():
    fd = dpb.FileDescriptorProto(name="hello.proto", package="hello", syntax="proto3")
    m = fd.message_type.add(name="HelloRequest")
    add_field(m, "name", 1, "string")
    add_field(m, "email", 2, "string")
    m = fd.message_type.add(name="HelloReply")
    add_field(m, "message", 1, "string")
    s = fd.service.add(name="HelloService")
    add_method(s, "SayHello", ".hello.HelloRequest", ".hello.HelloReply")
    return fd


def complex_file():
    P = ".com.example.complex"
    fd = dpb.FileDescriptorProto(name="complex.proto", package="com.example.complex", syntax="proto3")
    fd.dependency.append("google/protobuf/timestamp.proto")
    e = fd.enum_type.add(name="UserType")
    for i, n in enumerate(["USER_TYPE_UNSPECIFIED", "STANDARD", "PREMIUM", "ADMIN"]):
        e.value.add(name=n, number=i)
    m = fd.message_type.add(name="UserProfile")
    add_field(m, "user_id", 1, "string")
    add_field(m, "display_name", 2, "string")
    add_field(m, "email", 3, "string")
    add_field(m, "user_type", 4, "enum:" + P + ".UserType")
    add_field(m, "last_login", 5, ".google.protobuf.Timestamp")
    m = fd.message_type.add(name="GetUserProfileRequest")
    add_field(m, "user_id", 1, "string")
    m = fd.message_type.add(name="GetUserProfileResponse")
    add_field(m, "profile", 1, P + ".UserProfile")
    m = fd.message_type.add(name="StructuredMetadata")
    add_map(m, P + ".StructuredMetadata", "data", 1, "string", "string")
    m = fd.message_type.add(name="Document")
    add_field(m, "document_id", 1, "string")
    add_field(m, "title", 2, "string")
    add_field(m, "content", 3, "string")
    m.oneof_decl.add(name="metadata")
    add_field(m, "simple_summary", 4, "string", oneof=0)
    add_field(m, "structured_metadata_wrapper", 5, P + ".StructuredMetadata", oneof=0)
    m = fd.message_type.add(name="CreateDocumentRequest")
    add_field(m, "document", 1, P + ".Document")
    m = fd.message_type.add(name="CreateDocumentResponse")
    add_field(m, "document_id", 1, "string")
    add_field(m, "success", 2, "bool")
    m = fd.message_type.add(name="Node")
    add_field(m, "id", 1, "string")
    add_field(m, "value", 2, "string")
    add_field(m, "children", 3, P + ".Node", repeated=True)
    m = fd.message_type.add(name="ProcessNodeRequest")
    add_field(m, "root_node", 1, P + ".Node")
    m = fd.message_type.add(name="ProcessNodeResponse")
    add_field(m, "processed_summary", 1, "string")
    add_field(m, "total_nodes", 2, "int32")
    s = fd.service.add(name="UserProfileService")
    add_method(s, "GetUserProfile", P + ".GetUserProfileRequest", P + ".GetUserProfileResponse")
    s = fd.service.add(name="DocumentService")
    add_method(s, "CreateDocument", P + ".CreateDocumentRequest", P + ".CreateDocumentResponse")
    s = fd.service.add(name="NodeService")
    add_method(s, "ProcessNode", P + ".ProcessNodeRequest", P + ".ProcessNodeResponse")
    return fd


SCALARS = ["int32", "int64", "uint32", "uint64", "sint32", "sint64", "fixed32", "fixed64",
           "sfixed32", "sfixed64", "float", "double", "bool", "string", "bytes"]


def bench_file():
    P = ".bench"
    fd = dpb.FileDescriptorProto(name="bench.proto", package="bench", syntax="proto3")
    fd.dependency.append("google/protobuf/timestamp.proto")
    fd.dependency.append("google/protobuf/duration.proto")
    fd.dependency.append("complex.proto")
    e = fd.enum_type.add(name="Color")
    for n, v in [("COLOR_UNSPECIFIED", 0), ("RED", 1), ("GREEN", 2), ("BLUE", 3), ("NEGATIVE", -5),
                 ("BIG", 1 << 30)]:
        e.value.add(name=n, number=v)

    m = fd.message_type.add(name="Flat")
    for i in range(1, 9):
        add_field(m, "a%d" % i, i, "int32")
    for i in range(1, 5):
        add_field(m, "s%d" % i, 8 + i, "string")

    m = fd.message_type.add(name="Blob")
    add_field(m, "data", 1, "bytes")
    add_field(m, "name", 2, "string")

    m = fd.message_type.add(name="Inner")
    add_field(m, "x", 1, "int32")
    add_field(m, "y", 2, "string")

    m = fd.message_type.add(name="All")
    A = P + ".All"
    for i, t in enumerate(SCALARS):
        add_field(m, "f_" + t, 1 + i, t)
    add_field(m, "f_enum", 16, "enum:" + P + ".Color")
    add_field(m, "f_msg", 17, P + ".Inner")
    for i, t in enumerate(SCALARS):
        add_field(m, "r_" + t, 21 + i, t, repeated=True)
    add_field(m, "r_enum", 36, "enum:" + P + ".Color", repeated=True)
    add_field(m, "r_msg", 37, P + ".Inner", repeated=True)
    add_field(m, "r_unpacked", 38, "int32", repeated=True, packed=False)
    add_map(m, A, "m_str_int32", 41, "string", "int32")
    add_map(m, A, "m_int32_str", 42, "int32", "string")
    add_map(m, A, "m_int64_msg", 43, "int64", P + ".Inner")
    add_map(m, A, "m_bool_double", 44, "bool", "double")
    add_map(m, A, "m_uint64_bytes", 45, "uint64", "bytes")
    add_map(m, A, "m_str_enum", 46, "string", "enum:" + P + ".Color")
    add_map(m, A, "m_sint32_float", 47, "sint32", "float")
    add_map(m, A, "m_fixed64_sfixed32", 48, "fixed64", "sfixed32")
    m.oneof_decl.add(name="choice")
    add_field(m, "o_int32", 51, "int32", oneof=0)
    add_field(m, "o_string", 52, "string", oneof=0)
    add_field(m, "o_msg", 53, P + ".Inner", oneof=0)
    add_field(m, "o_bool", 54, "bool", oneof=0)
    add_field(m, "o_enum", 55, "enum:" + P + ".Color", oneof=0)
    # proto3 `optional`: explicit presence through synthetic oneofs (declared after real oneofs)
    m.oneof_decl.add(name="_opt_int32")
    m.oneof_decl.add(name="_opt_string")
    m.oneof_decl.add(name="_opt_bool")
    add_field(m, "opt_int32", 61, "int32", oneof=1, proto3_optional=True)
    add_field(m, "opt_string", 62, "string", oneof=2, proto3_optional=True)
    add_field(m, "opt_bool", 63, "bool", oneof=3, proto3_optional=True)
    add_field(m, "ts", 71, ".google.protobuf.Timestamp")
    add_field(m, "r_ts", 72, ".google.protobuf.Timestamp", repeated=True)
    add_field(m, "recursive", 81, A)
    add_field(m, "custom", 90, "string", json_name="CustomJSON")
    # declaration order != field-number order (protojson emits declaration order, wire is by number)
    add_field(m, "z_last", 100, "string")
    add_field(m, "late_low", 18, "int32")

    # request/response wrappers and echo services so the 4 KB shape exists in both directions
    m = fd.message_type.add(name="FlatReply")
    add_field(m, "flat", 1, P + ".Flat")
    s = fd.service.add(name="BenchService")
    add_method(s, "EchoFlat", P + ".Flat", P + ".Flat")
    add_method(s, "EchoNode", ".com.example.complex.Node", ".com.example.complex.Node")
    add_method(s, "EchoAll", P + ".All", P + ".All")
    add_method(s, "GetBlob", P + ".Inner", P + ".Blob")
    return fd


def mixed_file():
    """configs[4] (mixed replay): 28 generated unary methods Call00..Call27 over messages that mix every scalar
    kind, enums, nested / repeated messages, maps, oneofs, proto3 optional and Timestamp - drawn from a seeded
    generator (seed 0xB2000005, the config's seed) so the fixture is reproducible."""
    import random
    rng = random.Random(0xB2000005)
    P = ".mixed"
    fd = dpb.FileDescriptorProto(name="mixed.proto", package="mixed", syntax="proto3")
    fd.dependency.append("google/protobuf/timestamp.proto")
    e = fd.enum_type.add(name="Level")
    for n, v in [("LEVEL_UNSPECIFIED", 0), ("LOW", 1), ("MID", 2), ("HIGH", 3), ("MAX", 100), ("BELOW", -1)]:
        e.value.add(name=n, number=v)
    words = ["id", "name", "count", "total", "flag", "score", "ratio", "data", "tag", "kind", "note", "owner", "size", "rank",
             "price", "stamp", "label", "path", "mode", "limit", "offset", "token", "state", "weight", "key", "group", "zone"]
    # a few shared leaf messages
    subs = []
    for k in range(6):
        m = fd.message_type.add(name="Part%d" % k)
        names = rng.sample(words, 4)
        for i, nm in enumerate(names):
            add_field(m, "%s_%d" % (nm, i), i + 1, rng.choice(["string", "int32", "int64", "bool", "double", "uint32", "bytes", "sint32"]))
        subs.append(P + ".Part%d" % k)

    def gen_message(name):
        m = fd.message_type.add(name=name)
        fq = P + "." + name
        n_fields = rng.randint(4, 14)
        numbers = list(range(1, n_fields + 1))
        if rng.random() < 0.3:  # gaps and a high field number
            numbers[-1] = rng.choice([100, 536, 2047, 70000])
        used = set()
        oneof_left = 0
        have_oneof = False
        decl = []
        for i in range(n_fields):
            while True:
                nm = rng.choice(words) + rng.choice(["", "", "_x", "_value", "_list", "_id"])
                if nm not in used:
                    used.add(nm)
                    break
            decl.append((nm, numbers[i]))
        if rng.random() < 0.25:
            rng.shuffle(decl)  # declaration order != field number order
        for nm, num in decl:
            u = rng.random()
            if oneof_left > 0:
                add_field(m, nm, num, rng.choice(["string", "int32", "bool", rng.choice(subs), "enum:" + P + ".Level"]), oneof=0)
                oneof_left -= 1
            elif u < 0.40:
                add_field(m, nm, num, rng.choice(SCALARS))
            elif u < 0.46:
                add_field(m, nm, num, "enum:" + P + ".Level")
            elif u < 0.54:
                add_field(m, nm, num, rng.choice(subs))
            elif u < 0.66:
                add_field(m, nm, num, rng.choice(SCALARS), repeated=True)
            elif u < 0.70:
                add_field(m, nm, num, "int32", repeated=True, packed=False)
            elif u < 0.76:
                add_field(m, nm, num, rng.choice(subs), repeated=True)
            elif u < 0.86:
                k, v = rng.choice([("string", "string"), ("string", "int32"), ("int32", "string"), ("string", rng.choice(subs)),
                                   ("uint64", "bool"), ("string", "double")])
                add_map(m, fq, nm, num, k, v)
            elif u < 0.92 and not have_oneof:
                m.oneof_decl.add(name="choice")
                have_oneof = True
                oneof_left = rng.randint(1, 2)
                add_field(m, nm, num, rng.choice(["string", "int64"]), oneof=0)
            elif u < 0.96:
                add_field(m, nm, num, ".google.protobuf.Timestamp")
            else:
                add_field(m, nm, num, rng.choice(["string", "int32", "bool"]))
        return fq

    s = fd.service.add(name="MixedService")
    for k in range(28):
        add_method(s, "Call%02d" % k, gen_message("Req%02d" % k), gen_message("Rep%02d" % k))
    return fd


def wkt_file():
    """well-known types next to Timestamp: Duration, the nine wrappers and Empty as singular fields,
    list elements, map values, oneof members and as the request / reply message of a method; Struct is there to be
    refused"""
    P = ".wkt"
    G = ".google.protobuf."
    fd = dpb.FileDescriptorProto(name="wkt.proto", package="wkt", syntax="proto3")
    for dep in ("timestamp", "duration", "wrappers", "empty", "struct", "field_mask"):
        fd.dependency.append("google/protobuf/%s.proto" % dep)
    m = fd.message_type.add(name="Wkt")
    W = P + ".Wkt"
    add_field(m, "d", 1, G + "Duration")
    add_field(m, "r_d", 2, G + "Duration", repeated=True)
    add_map(m, W, "m_d", 3, "string", G + "Duration")
    add_field(m, "e", 4, G + "Empty")
    add_field(m, "bv", 5, G + "BoolValue")
    add_field(m, "i32", 6, G + "Int32Value")
    add_field(m, "i64", 7, G + "Int64Value")
    add_field(m, "u32", 8, G + "UInt32Value")
    add_field(m, "u64", 9, G + "UInt64Value")
    add_field(m, "fv", 10, G + "FloatValue")
    add_field(m, "dv", 11, G + "DoubleValue")
    add_field(m, "sv", 12, G + "StringValue")
    add_field(m, "byv", 13, G + "BytesValue")
    add_field(m, "r_sv", 14, G + "StringValue", repeated=True)
    add_map(m, W, "m_i64", 15, "int32", G + "Int64Value")
    m.oneof_decl.add(name="choice")
    add_field(m, "o_d", 16, G + "Duration", oneof=0)
    add_field(m, "o_sv", 17, G + "StringValue", oneof=0)
    add_field(m, "r_e", 18, G + "Empty", repeated=True)
    add_field(m, "ts", 19, G + "Timestamp")
    add_field(m, "r_i32", 20, G + "Int32Value", repeated=True)
    add_field(m, "name", 21, "string")
    add_field(m, "inner", 22, W)
    add_field(m, "fm", 23, G + "FieldMask")
    add_field(m, "r_fm", 24, G + "FieldMask", repeated=True)
    m = fd.message_type.add(name="HasStruct")
    add_field(m, "s", 1, G + "Struct")
    add_field(m, "x", 2, "int32")
    s = fd.service.add(name="WktService")
    add_method(s, "Ping", G + "Empty", G + "Duration")
    add_method(s, "Wait", G + "Duration", G + "Empty")
    add_method(s, "Rename", G + "StringValue", G + "Int64Value")
    add_method(s, "EchoWkt", W, W)
    add_method(s, "UpdateMask", G + "FieldMask", G + "FieldMask")
    return fd


def build_set():
    fds = dpb.FileDescriptorSet()
    ts = fds.file.add()
    timestamp_pb2.DESCRIPTOR.CopyToProto(ts)
    du = fds.file.add()
    duration_pb2.DESCRIPTOR.CopyToProto(du)
    # source_code_info is not needed on the hot path; strip to keep the fixture small
    ts.ClearField("source_code_info")
    du.ClearField("source_code_info")
    for mod in (wrappers_pb2, empty_pb2, struct_pb2, field_mask_pb2):
        f = fds.file.add()
        mod.DESCRIPTOR.CopyToProto(f)
        f.ClearField("source_code_info")
    fds.file.append(hello_file())
    fds.file.append(complex_file())
    fds.file.append(bench_file())
    fds.file.append(mixed_file())
    fds.file.append(wkt_file())
    return fds


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "schemas.binpb")
    data = build_set().SerializeToString(deterministic=True)
    with open(out, "wb") as fh:
        fh.write(data)
    print("wrote %s (%d bytes)" % (out, len(data)))


if __name__ == "__main__":
    sys.exit(main())
