"""Seeded generators of random messages for the fixture schemas, built on python-protobuf (upb).

Test infrastructure: python-protobuf is the SECONDARY oracle of SURVEY.md section 8(c) - an
independent implementation of the protobuf JSON mapping that is authoritative for canonical wire
bytes (deterministic serialization = ascending field number, sorted map keys) and for JSON after
re-parsing.
"""
import math
import os
import random
import struct

from google.protobuf import descriptor_pb2, descriptor_pool, json_format, message_factory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_pool():
    fds = descriptor_pb2.FileDescriptorSet()
    with open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb") as fh:
        fds.ParseFromString(fh.read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    return pool


_POOL = None


def cls(name):
    global _POOL
    if _POOL is None:
        _POOL = load_pool()
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(name))


ASCII = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 _-.,:;!?@#$%^*()[]{}+=~|'/"
SPECIAL = ['"', "\\", "\n", "\t", "\r", "\b", "\f", "\x01", "\x1f", "<", ">", "&", " ", " ", "\x7f", "/"]
MULTI = ["é", "ö", "ñ", "张", "三", "日本", "😀", "𝄞", " ", "�", "€"]


def rand_string(rng, lo=0, hi=24, p_special=0.15, p_multi=0.15):
    n = rng.randint(lo, hi)
    out = []
    for _ in range(n):
        r = rng.random()
        if r < p_special:
            out.append(rng.choice(SPECIAL))
        elif r < p_special + p_multi:
            out.append(rng.choice(MULTI))
        else:
            out.append(rng.choice(ASCII))
    return "".join(out)


def rand_int(rng, bits, signed):
    r = rng.random()
    if r < 0.3:
        v = rng.randint(0, 127)
    elif r < 0.5:
        v = rng.randint(0, (1 << (bits - 1)) - 1)
    elif r < 0.6:
        v = rng.choice([0, 1, (1 << (bits - 1)) - 1, (1 << bits) - 1 if not signed else -(1 << (bits - 1)), 1 << 31,
                        (1 << 32) - 1, 1 << 53, (1 << 53) + 1])
    else:
        v = rng.getrandbits(bits)
    if signed:
        v &= (1 << bits) - 1
        if v >= 1 << (bits - 1):
            v -= 1 << bits
        if rng.random() < 0.25:
            v = -abs(v) if v != -(1 << (bits - 1)) else v
    else:
        v &= (1 << bits) - 1
    return v


def rand_double(rng):
    r = rng.random()
    if r < 0.15:
        return float(rng.randint(-1000, 1000))
    if r < 0.3:
        return rng.choice([0.0, -0.0, 1e-7, 1e-6, 1e21, 1e20, 0.1, 1 / 3, 5e-324, 1.7976931348623157e308, 2.5e-5,
                           123456789012345680000.0, 9007199254740993.0, float("inf"), float("-inf"), float("nan"),
                           2.0 ** -1074, 2.0 ** 52, 1e22, 1e23, 4.35e-7])
    if r < 0.6:
        return struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
    return rng.uniform(-1e6, 1e6) * 10 ** rng.randint(-12, 12)


def rand_float(rng):
    r = rng.random()
    if r < 0.15:
        return float(rng.randint(-1000, 1000))
    if r < 0.3:
        return rng.choice([0.0, -0.0, 1e-7, 1e-6, 1e21, 0.1, 3.4028235e38, 1.401298464324817e-45, float("inf"),
                           float("-inf"), float("nan"), 16777216.0, 0.3])
    if r < 0.6:
        return struct.unpack("<f", struct.pack("<I", rng.getrandbits(32)))[0]
    return struct.unpack("<f", struct.pack("<f", rng.uniform(-1e6, 1e6) * 10 ** rng.randint(-8, 8)))[0]


def rand_scalar(rng, fd):
    from google.protobuf.descriptor import FieldDescriptor as FD
    t = fd.type
    if t in (FD.TYPE_INT32, FD.TYPE_SINT32, FD.TYPE_SFIXED32):
        return rand_int(rng, 32, True)
    if t in (FD.TYPE_INT64, FD.TYPE_SINT64, FD.TYPE_SFIXED64):
        return rand_int(rng, 64, True)
    if t in (FD.TYPE_UINT32, FD.TYPE_FIXED32):
        return rand_int(rng, 32, False)
    if t in (FD.TYPE_UINT64, FD.TYPE_FIXED64):
        return rand_int(rng, 64, False)
    if t == FD.TYPE_BOOL:
        return rng.random() < 0.5
    if t == FD.TYPE_FLOAT:
        return rand_float(rng)
    if t == FD.TYPE_DOUBLE:
        return rand_double(rng)
    if t == FD.TYPE_STRING:
        return rand_string(rng)
    if t == FD.TYPE_BYTES:
        return bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 20)))
    if t == FD.TYPE_ENUM:
        if rng.random() < 0.1:
            return rng.choice([7, -9, 1234567])  # unknown number (proto3 open enum)
        return rng.choice([v.number for v in fd.enum_type.values])
    raise AssertionError(t)


def fill(rng, msg, depth=0, p_field=0.35, floats=True):
    """Populate msg (a python protobuf message) at random."""
    from google.protobuf.descriptor import FieldDescriptor as FD
    d = msg.DESCRIPTOR
    if d.full_name == "google.protobuf.Timestamp":
        msg.seconds = rng.choice([0, 1704110400, -62135596800, 253402300799, rng.randint(-62135596800, 253402300799)])
        msg.nanos = rng.choice([0, 0, 500000000, 123000, 999999999, 1, rng.randint(0, 999999999)])
        return
    if d.full_name == "google.protobuf.Duration":
        sign = rng.choice([1, 1, -1])
        msg.seconds = sign * rng.choice([0, 0, 1, 3600, 315576000000, rng.randint(0, 315576000000)])
        msg.nanos = sign * rng.choice([0, 0, 500000000, 123000, 999999999, 1, rng.randint(0, 999999999)])
        return
    if d.full_name == "google.protobuf.FieldMask":
        seg = lambda: rng.choice(["a", "foo", "foo_bar", "x1", "user_id", "display_name", "b2_c"])
        for _ in range(rng.randint(0, 4)):
            msg.paths.append(".".join(seg() for _ in range(rng.randint(1, 3))))
        return
    for fd in d.fields:
        if rng.random() > p_field:
            continue
        if not floats and fd.type in (FD.TYPE_FLOAT, FD.TYPE_DOUBLE):
            continue
        if fd.message_type is not None and fd.message_type.GetOptions().map_entry:
            kf = fd.message_type.fields_by_name["key"]
            vf = fd.message_type.fields_by_name["value"]
            if not floats and vf.type in (FD.TYPE_FLOAT, FD.TYPE_DOUBLE):
                continue
            m = getattr(msg, fd.name)
            for _ in range(rng.randint(1, 4)):
                k = rand_scalar(rng, kf)
                if vf.type == FD.TYPE_MESSAGE:
                    if depth < 3:
                        fill(rng, m[k], depth + 1, p_field, floats)
                    else:
                        m[k].SetInParent()
                else:
                    m[k] = rand_scalar(rng, vf)
        elif fd.label == FD.LABEL_REPEATED:
            lst = getattr(msg, fd.name)
            for _ in range(rng.randint(1, 4)):
                if fd.type == FD.TYPE_MESSAGE:
                    sub = lst.add()
                    if depth < 3:
                        fill(rng, sub, depth + 1, p_field, floats)
                else:
                    lst.append(rand_scalar(rng, fd))
        elif fd.type == FD.TYPE_MESSAGE:
            sub = getattr(msg, fd.name)
            sub.SetInParent()
            if depth < 3:
                fill(rng, sub, depth + 1, p_field * 0.6, floats)
        else:
            setattr(msg, fd.name, rand_scalar(rng, fd))


def random_message(name, seed, **kw):
    rng = random.Random(seed)
    m = cls(name)()
    fill(rng, m, **kw)
    return m


def wire(m):
    return m.SerializeToString(deterministic=True)


def to_json(m, proto_names=False):
    return json_format.MessageToJson(m, indent=None, ensure_ascii=False, preserving_proto_field_name=proto_names)
