"""ad-hoc GPU debugging helper (not a test)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cases, orc, ggrmcp_b200
from ggrmcp_b200.engine import pack, unpack
fds = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schemas.binpb"), "rb").read()
eng = ggrmcp_b200.Engine(0); sch = eng.register(fds); O = orc.Schema(fds)
items = cases.random_decode_cases(120)
ol = [O.decode(n, b) for n, b in items]

def bad_set(sub):
    ids = np.array([sch.message(items[i][0]) for i in sub], np.int32)
    data, off = pack([items[i][1] for i in sub])
    out, ooff, st = eng.decode_batch(sch, ids, data, off)
    eo = unpack(out, ooff)
    return [i for k, i in enumerate(sub) if ol[i][0] == 0 and int(st[k]) == 0 and eo[k] != ol[i][1]]

full = list(range(len(items)))
print("full:", bad_set(full)[:10])
lo, hi = 0, len(items)
# shrink from the left
step = 1024
while step >= 1:
    while lo + step <= 589 and 589 in bad_set(list(range(lo + step, hi))):
        lo += step
    step //= 2
step = 1024
while step >= 1:
    while hi - step > 589 and 589 in bad_set(list(range(lo, hi - step))):
        hi -= step
    step //= 2
print("minimal window", lo, hi, "bad:", bad_set(list(range(lo, hi))))
sub = list(range(lo, hi))
# try dropping single items
need = []
for x in list(sub):
    if x == 589: continue
    t = [i for i in sub if i != x]
    if 589 not in bad_set(t):
        need.append(x)
print("items whose removal fixes 589:", need)
for x in need[:6] + [589]:
    print(x, items[x][0], "oracle st", ol[x][0], items[x][1].hex()[:300])
