"""debug aid: where the time of the mixed replay goes.  Decodes / encodes subsets of the workload (by item size, by the tier that
takes them) through the host entry points and prints wall times (copies included; the per-thread kernels dominate)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import benchgen, ggrmcp_b200
from ggrmcp_b200.engine import pack
fds = open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb").read()
eng = ggrmcp_b200.Engine(0)
schema = eng.register(fds)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wl = benchgen.mixed(n, schema.message)
L = ggrmcp_b200.engine._load()
L.ggr_debug_paths.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
wlen = np.diff(wl.rep_off.astype(np.int64)); jlen = np.diff(wl.req_off.astype(np.int64))
print("items", n, "reply wire bytes", int(wlen.sum()), "mean", wlen.mean(), "request json bytes", int(jlen.sum()))
print("reply size percentiles", np.percentile(wlen, [50, 90, 99, 99.9, 100]).astype(int))

def subset(data, off, msg, idx):
    parts = [data[int(off[i]):int(off[i + 1])] for i in idx]
    d, o = pack([bytes(p) for p in parts])
    return d, o, msg[idx].copy()

def timed(fn, *a, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(*a); best = min(best, time.perf_counter() - t0)
    return best * 1e3, r

# whole batch, then read which tier took what
ms, (out, ooff, st) = timed(eng.decode_batch, schema, wl.rep_msg, wl.rep_wire, wl.rep_off)
mode = np.zeros(n, np.uint32); cnt = np.zeros(3, np.uint32)
L.ggr_debug_paths(eng.h, n, mode.ctypes.data, cnt.ctypes.data)
print("decode all: %.1f ms; modes" % ms, dict(zip(*np.unique(mode, return_counts=True))), "errors", int((st != 0).sum()))
pend = np.nonzero(mode != 2)[0]
for lo, hi in ((0, 640), (640, 4096), (4096, 16384), (16384, 1 << 30)):
    sel = pend[(wlen[pend] >= lo) & (wlen[pend] < hi)]
    if len(sel) == 0: continue
    d, o, m = subset(wl.rep_wire, wl.rep_off, wl.rep_msg, sel)
    ms, _ = timed(eng.decode_batch, schema, m, d, o)
    print("decode per-thread items with wire in [%d, %d): %d items, %d bytes: %.1f ms" % (lo, hi, len(sel), len(d), ms))
    slow = sel[mode[sel] == 1]
    print("   of them in slow mode (declaration order != wire order):", len(slow))
big = int(np.argmax(wlen))
for tag, sel in (("largest item alone", np.array([big])), ("largest per-thread item alone", np.array([pend[np.argmax(wlen[pend])]]))):
    d, o, m = subset(wl.rep_wire, wl.rep_off, wl.rep_msg, sel)
    ms, _ = timed(eng.decode_batch, schema, m, d, o)
    print("decode %s (%d bytes, mode %d, msg %d): %.2f ms" % (tag, wlen[sel[0]], mode[sel[0]], wl.rep_msg[sel[0]], ms))
sel = np.nonzero(mode == 2)[0]
d, o, m = subset(wl.rep_wire, wl.rep_off, wl.rep_msg, sel)
ms, _ = timed(eng.decode_batch, schema, m, d, o)
print("decode lock-step items only: %d items %d bytes: %.1f ms" % (len(sel), len(d), ms))

ms, (out, ooff, st) = timed(eng.encode_batch, schema, wl.req_msg, wl.req_json, wl.req_off)
L.ggr_debug_paths(eng.h, n, None, cnt.ctypes.data)
print("encode all: %.1f ms; lists {lock-step, left by walker, per-thread}" % ms, cnt.tolist(), "errors", int((st != 0).sum()))
for lo, hi in ((0, 1024), (1024, 4096), (4096, 16384), (16384, 1 << 30)):
    sel = np.nonzero((jlen >= lo) & (jlen < hi))[0]
    d, o, m = subset(wl.req_json, wl.req_off, wl.req_msg, sel)
    ms, _ = timed(eng.encode_batch, schema, m, d, o)
    L.ggr_debug_paths(eng.h, len(sel), None, cnt.ctypes.data)
    print("encode items with json in [%d, %d): %d items, %d bytes: %.1f ms lists %s" % (lo, hi, len(sel), len(d), ms, cnt.tolist()))
sel = np.array([int(np.argmax(jlen))])
d, o, m = subset(wl.req_json, wl.req_off, wl.req_msg, sel)
ms, _ = timed(eng.encode_batch, schema, m, d, o)
print("encode largest item alone (%d bytes): %.2f ms" % (jlen[sel[0]], ms))
