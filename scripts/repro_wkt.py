"""debug aid: WKT_ENCODE through the engine under the small_chunks switches; prints the items that differ from the oracle"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1 and sys.argv[1] == "small":
    os.environ.update({"GGR_CHUNK_ITEMS": "128", "GGR_SLOTS": "2", "GGR_LOCKSTEP_MIN_BYTES": "0"})
else:
    os.environ.update({"GGR_LOCKSTEP_MIN_BYTES": "0"})
import cases, orc, ggrmcp_b200
from ggrmcp_b200.engine import pack
fds = open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb").read()
eng = ggrmcp_b200.Engine(0)
schema = eng.register(fds)
O = orc.Schema(fds)
def run(items, tag):
    ids = np.array([schema.message(n) for n, _ in items], np.int32)
    data, off = pack([j for _, j in items])
    out, ooff, st = eng.encode_batch(schema, ids, data, off)
    bad = 0
    for i, (n, j) in enumerate(items):
        ost, ow, _ = O.encode(n, j)
        got = bytes(out[int(ooff[i]):int(ooff[i + 1])])
        if (ost == 0) != (st[i] == 0) or (ost == 0 and got != ow):
            bad += 1
            print(tag, "DIFF", i, n, j, ost, int(st[i]), ow.hex(), got.hex())
    print(tag, "items", len(items), "bad", bad)
sel = sys.argv[2] if len(sys.argv) > 2 else "all"
W = cases.WKT_ENCODE
if sel == "all":
    for rep in range(3):
        run(W, "full%d" % rep)
    run(W[-1:], "last1")
    run(W[-3:], "last3")
    run(W[-20:], "last20")
    run(W[:-1] + W[-1:] * 5, "dup5")
else:
    run(W, "full")
