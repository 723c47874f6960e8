"""Small parity run for compute-sanitizer (scripts/sanitize.sh): a few hundred items of every shape
through the host entry points, compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import benchgen
import cases
import ggrmcp_b200
import ggrmcp_b200.engine
import orc

fds = open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb").read()
eng = ggrmcp_b200.Engine(0)
schema = eng.register(fds)
O = orc.Schema(fds)
ok = 0
# random + edge cases
enc = [(n, j) for n, j in cases.random_encode_cases(8)] + [(n, j) for n, j, _ in cases.ENCODE_EDGE]
ids = np.array([schema.message(n) for n, _ in enc], np.int32)
data, off = ggrmcp_b200.engine.pack([j for _, j in enc])
out, ooff, st = eng.encode_batch(schema, ids, data, off)
for i, (n, j) in enumerate(enc):
    ost, ow, _ = O.encode(n, j)
    if st[i] == 11 and ost == 0:
        continue  # documented gap
    assert (ost == 0) == (st[i] == 0), (n, j[:80], ost, st[i])
    if ost == 0:
        assert bytes(out[int(ooff[i]):int(ooff[i + 1])]) == ow, (n, j[:80])
        ok += 1
dec = [(n, w) for n, w in cases.random_decode_cases(8)]
ids = np.array([schema.message(n) for n, _ in dec], np.int32)
data, off = ggrmcp_b200.engine.pack([w for _, w in dec])
out, ooff, st = eng.decode_batch(schema, ids, data, off)
for i, (n, w) in enumerate(dec):
    ost, oj, _ = O.decode(n, w, 0)
    if st[i] == 11 and ost == 0:
        continue  # documented gap
    assert (ost == 0) == (st[i] == 0), (n, w.hex()[:80], ost, st[i])
    if ost == 0:
        assert bytes(out[int(ooff[i]):int(ooff[i + 1])]) == oj, (n, w.hex()[:80])
        ok += 1
# bench shapes
wl = benchgen.nested(192, schema.message)
o1, f1, s1 = eng.encode_batch(schema, wl.req_msg, wl.req_json, wl.req_off)
o2, f2, s2 = eng.decode_batch(schema, wl.rep_msg, wl.rep_wire, wl.rep_off)
r1, rf1, rs1 = O.encode_batch(wl.req_msg, wl.req_json, wl.req_off, threads=2)
r2, rf2, rs2 = O.decode_batch(wl.rep_msg, wl.rep_wire, wl.rep_off, threads=2, cap=int(len(wl.rep_wire) * 3 + 64 * wl.n + 4096))
assert int((s1 != 0).sum()) == 0 and int((s2 != 0).sum()) == 0
assert bytes(o1[: int(f1[-1])]) == bytes(r1) and bytes(o2[: int(f2[-1])]) == bytes(r2)
ok += 2 * wl.n
# HTTP-body boundary: request bodies in, result bodies out
from ggrmcp_b200.engine import pack
tool_of = {}
for mi in reversed(schema.methods()):
    tool_of[mi["input_msg"]] = mi["tool_name"].encode()
jb = wl.req_json.tobytes()
bodies = [b'{"jsonrpc":"2.0","id":%d,"method":"tools/call","params":{"name":"%s","arguments":%s}}'
          % (i, tool_of[int(wl.req_msg[i])], jb[int(wl.req_off[i]):int(wl.req_off[i + 1])]) for i in range(wl.n)]
bodies += [b'{"jsonrpc":"2.0","id":1.5,"method":"tools/call","params":{"name":"x","arguments":{}}}', b"{", b""]
data, off = pack(bodies)
wout, woff, method, span, st = eng.request_batch(schema, data, off)
for i in range(wl.n):
    assert st[i] == 0 and bytes(wout[int(woff[i]):int(woff[i + 1])]) == bytes(r1[int(rf1[i]):int(rf1[i + 1])])
assert list(st[wl.n:]) == [11, 11, 11]
ids, ioff = pack([b"%d" % i for i in range(wl.n)])
bout, boff, st = eng.decode_wrap_batch(schema, wl.rep_msg, wl.rep_wire, wl.rep_off, ids, ioff)
names = {schema.message(n): n for n in ("com.example.complex.Node", "com.example.complex.GetUserProfileResponse")}
rw = wl.rep_wire.tobytes()
for i in range(0, wl.n, 7):
    rc, body = O.response(names[int(wl.rep_msg[i])], rw[int(wl.rep_off[i]):int(wl.rep_off[i + 1])], b"%d" % i)
    assert st[i] == 0 and bytes(bout[int(boff[i]):int(boff[i + 1])]) == body
ok += 2 * wl.n
print("ok items", ok)
