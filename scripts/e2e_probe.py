"""Timing probe for the host-buffer entry points: copies alone vs encode / decode calls."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ggrmcp_b200, benchgen
from ggrmcp_b200.engine import _load
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
eng = ggrmcp_b200.Engine(0)
schema = eng.register(open(os.path.join(ROOT, "tests/golden/schemas.binpb"), "rb").read())
wl = benchgen.nested(n, schema.message)
L = _load()
def pinned(a):
    t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
    t.numpy()[:] = a.view(np.uint8).reshape(-1)
    return t
h_req, h_off, h_msg = pinned(wl.req_json), pinned(wl.req_off), pinned(wl.req_msg)
cap = int(len(wl.req_json) * 1.2) + 4096
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
h_oo = torch.empty((n + 1) * 8, dtype=torch.uint8).pin_memory()
h_st = torch.empty(n * 4, dtype=torch.uint8).pin_memory()
d = torch.empty(len(wl.req_json), dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h_req, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(h_req, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("H2D %.1f MB in %.2f ms = %.1f GB/s" % (len(wl.req_json) / 1e6, (t1 - t0) * 1e3, len(wl.req_json) / (t1 - t0) / 1e9))
t0 = time.perf_counter(); h_out[: len(wl.req_json)].copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("D2H %.2f ms = %.1f GB/s" % ((t1 - t0) * 1e3, len(wl.req_json) / (t1 - t0) / 1e9))
def enc():
    rc = L.ggr_encode_batch(eng.h, schema.h, n, h_msg.data_ptr(), h_req.data_ptr(), h_off.data_ptr(), h_out.data_ptr(), cap,
                            h_oo.data_ptr(), h_st.data_ptr(), 0)
    assert rc == 0, rc
for _ in range(2): enc()
ts = []
for _ in range(4):
    t0 = time.perf_counter(); enc(); ts.append((time.perf_counter() - t0) * 1e3)
print("encode_batch(host) ms:", [round(t, 2) for t in ts], "slots", os.environ.get("GGR_SLOTS"), "chunk", os.environ.get("GGR_CHUNK_ITEMS"))
