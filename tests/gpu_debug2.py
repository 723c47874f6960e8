import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cases, orc, ggrmcp_b200
from ggrmcp_b200.engine import pack, unpack
fds = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schemas.binpb"), "rb").read()
eng = ggrmcp_b200.Engine(0); sch = eng.register(fds); O = orc.Schema(fds)
items = cases.random_decode_cases(120)[512:590]
ids = np.array([sch.message(n) for n, _ in items], np.int32)
data, off = pack([b for _, b in items])
out, ooff, st = eng.decode_batch(sch, ids, data, off)
eo = unpack(out, ooff)
bad = [i for i, (n, b) in enumerate(items) if O.decode(n, b)[0] == 0 and int(st[i]) == 0 and eo[i] != O.decode(n, b)[1]]
print("bad", bad, "statuses", [int(x) for x in st])
