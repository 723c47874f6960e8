"""which kernel tier takes the items of a workload (debug aid): python scripts/debug_paths.py [nested|mixed|flat] [items]"""
import ctypes as C
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, ggrmcp_b200, bench
kind = sys.argv[1] if len(sys.argv) > 1 else "nested"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 151552
eng = ggrmcp_b200.Engine(0)
schema = eng.register(bench.load_fds())
R = bench.Resident(torch, eng, schema, kind, n, 0, torch.device("cuda", 0))
R.warm(3)
L = ggrmcp_b200.engine._load()
mode = np.zeros(n, np.uint32); cnt = np.zeros(3, np.uint32)
L.ggr_debug_paths.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
L.ggr_debug_paths(eng.h, n, mode.ctypes.data, cnt.ctypes.data)
wlen = np.diff(R.wl.rep_off.astype(np.int64))
print("reply modes:", dict(zip(*np.unique(mode, return_counts=True))))
pend = mode != 2
print("reply per-thread items:", int(pend.sum()), "of", n, "wire bytes min/mean/max", (wlen[pend].min(), wlen[pend].mean(), wlen[pend].max()) if pend.any() else None)
big = wlen >= 640
print("  of those routed to the lock-step tier (>= 640 B):", int((pend & big).sum()), "sizes", np.percentile(wlen[pend & big], [0, 50, 100]) if (pend & big).any() else None)
print("request lists {lock-step, left by walker, per-thread}:", cnt.tolist())
