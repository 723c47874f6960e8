"""Wire-level mutators used by the parity tests: reorder / duplicate / split fields, inject unknown
fields, truncate - everything a non-canonical (but legal, or illegal) backend reply can look like."""
import random


def read_varint(b, i):
    v = 0
    s = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return v, i


def put_varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def split_fields(b):
    """-> list of (number, wire_type, raw_bytes_including_tag) for a well-formed message payload."""
    out = []
    i = 0
    while i < len(b):
        s = i
        tag, i = read_varint(b, i)
        wt = tag & 7
        if wt == 0:
            _, i = read_varint(b, i)
        elif wt == 1:
            i += 8
        elif wt == 5:
            i += 4
        elif wt == 2:
            n, i = read_varint(b, i)
            i += n
        else:
            raise ValueError("group")
        out.append((tag >> 3, wt, b[s:i]))
    return out


def shuffle(b, rng):
    f = split_fields(b)
    rng.shuffle(f)
    return b"".join(x[2] for x in f)


def duplicate_some(b, rng):
    f = split_fields(b)
    out = []
    for x in f:
        out.append(x)
        if rng.random() < 0.3:
            out.insert(rng.randrange(len(out)), x)
    return b"".join(x[2] for x in out)


def inject_unknown(b, rng):
    f = [x[2] for x in split_fields(b)]
    for _ in range(rng.randint(1, 3)):
        num = rng.choice([19, 200, 1000, 99999, 536870911])
        kind = rng.randint(0, 3)
        if kind == 0:
            raw = put_varint(num << 3) + put_varint(rng.getrandbits(rng.choice([7, 31, 63])))
        elif kind == 1:
            raw = put_varint((num << 3) | 1) + bytes(rng.getrandbits(8) for _ in range(8))
        elif kind == 2:
            n = rng.randint(0, 40)
            raw = put_varint((num << 3) | 2) + put_varint(n) + bytes(rng.getrandbits(8) for _ in range(n))
        else:
            raw = put_varint((num << 3) | 5) + bytes(rng.getrandbits(8) for _ in range(4))
        f.insert(rng.randrange(len(f) + 1), raw)
    return b"".join(f)


def truncate(b, rng):
    if len(b) < 2:
        return b
    return b[: rng.randrange(1, len(b))]


def corrupt(b, rng):
    if not b:
        return b
    a = bytearray(b)
    for _ in range(rng.randint(1, 3)):
        a[rng.randrange(len(a))] = rng.getrandbits(8)
    return bytes(a)
