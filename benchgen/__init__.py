"""Synthetic workloads for bench.py and the full-size parity tests.

Benchmark infrastructure - neither the product nor the oracle.  gen.cc is compiled with g++ into
benchgen/libggr_benchgen.so (built by __graft_entry__.build(); the .so travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libggr_benchgen.so")
SEEDS = {2: 0xB2000002, 3: 0xB2000003, 4: 0xB2000004, 5: 0xB2000005}
_lib = None


def build(force=False):
    src = os.path.join(HERE, "gen.cc")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB, src])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.ggr_gen_flat.argtypes = [C.c_uint64, C.c_int64, vp, C.c_uint64, vp, vp, C.c_uint64, vp]
        L.ggr_gen_nested.argtypes = [C.c_uint64, C.c_int64, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp]
        L.ggr_gen_blob.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, vp, C.c_uint64, vp]
        _lib = L
    return _lib


class Workload:
    """req_json/req_off/req_msg: request side (canonical arguments); rep_wire/rep_off/rep_msg: reply side."""

    def __init__(self, name):
        self.name = name
        self.req_json = self.req_off = self.req_msg = None
        self.rep_wire = self.rep_off = self.rep_msg = None

    @property
    def n(self):
        return len(self.rep_msg) if self.rep_msg is not None else len(self.req_msg)


def flat(n, msg_index, seed=SEEDS[2], first=0):
    """Config 2.  msg_index: callable full_name -> message index of the engine/oracle in use.
    `first` offsets the item indices (rank sharding: rank r generates items [r*n, (r+1)*n))."""
    L = _load()
    jc, wc = n * 700 + 4096, n * 320 + 4096
    j = np.empty(jc, np.uint8); jo = np.empty(n + 1, np.uint64)
    w = np.empty(wc, np.uint8); wo = np.empty(n + 1, np.uint64)
    rc = L.ggr_gen_flat(seed + 0x100000001B3 * first, n, j.ctypes.data, jc, jo.ctypes.data, w.ctypes.data, wc, wo.ctypes.data)
    assert rc == 0
    wl = Workload("flat")
    wl.req_json, wl.req_off = j[: int(jo[n])].copy(), jo
    wl.rep_wire, wl.rep_off = w[: int(wo[n])].copy(), wo
    wl.req_msg = np.full(n, msg_index("bench.Flat"), np.int32)
    wl.rep_msg = wl.req_msg.copy()
    return wl


def nested(n, msg_index, seed=SEEDS[3], first=0):
    """Config 3."""
    L = _load()
    jc, wc = n * 9000 + 4096, n * 7000 + 4096
    j = np.empty(jc, np.uint8); jo = np.empty(n + 1, np.uint64)
    w = np.empty(wc, np.uint8); wo = np.empty(n + 1, np.uint64)
    kind = np.empty(n, np.int32)
    rc = L.ggr_gen_nested(seed + 0x100000001B3 * first, n, j.ctypes.data, jc, jo.ctypes.data, kind.ctypes.data, w.ctypes.data, wc,
                          wo.ctypes.data)
    assert rc == 0
    wl = Workload("nested")
    wl.req_json, wl.req_off = j[: int(jo[n])].copy(), jo
    wl.rep_wire, wl.rep_off = w[: int(wo[n])].copy(), wo
    P = "com.example.complex."
    req = np.array([msg_index(P + "ProcessNodeRequest"), msg_index(P + "CreateDocumentRequest")], np.int32)
    rep = np.array([msg_index(P + "Node"), msg_index(P + "GetUserProfileResponse")], np.int32)
    wl.req_msg, wl.rep_msg = req[kind], rep[kind]
    wl.kind = kind
    return wl


def blob(n, msg_index, payload=65536, seed=SEEDS[4], first=0):
    """Config 4 (reply side only)."""
    L = _load()
    wc = n * (payload + 64) + 4096
    w = np.empty(wc, np.uint8); wo = np.empty(n + 1, np.uint64)
    rc = L.ggr_gen_blob(seed + 0x100000001B3 * first, n, payload, w.ctypes.data, wc, wo.ctypes.data)
    assert rc == 0
    wl = Workload("blob")
    wl.rep_wire, wl.rep_off = w[: int(wo[n])].copy(), wo
    wl.rep_msg = np.full(n, msg_index("bench.Blob"), np.int32)
    return wl
