"""ctypes binding of the CPU oracle (oracle/libggr_oracle.so) - test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libggr_oracle.so")

OK, SYNTAX, UNKNOWN_FIELD, INVALID_VALUE, RANGE, INVALID_UTF8, DUPLICATE, ONEOF, DEPTH, TOO_LARGE, \
    BAD_WIRE, UNSUPPORTED, NO_SPACE = range(13)
F_COMMA_SPACE = 1
F_GO_LEGACY_ORDER = 2


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


class RequestOut(C.Structure):
    _fields_ = [("kind", C.c_int32), ("status", C.c_int32), ("method", C.c_int32),
                ("args", C.c_void_p), ("args_n", C.c_size_t),
                ("wire", C.c_void_p), ("wire_n", C.c_size_t),
                ("id", C.c_void_p), ("id_n", C.c_size_t),
                ("resp", C.c_void_p), ("resp_n", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_schema_new.restype = C.c_void_p
        L.orc_schema_new.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_schema_new2.restype = C.c_void_p
        L.orc_schema_new2.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_schema_free.argtypes = [C.c_void_p]
        L.orc_message_index.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_message_index.restype = C.c_int32
        L.orc_method_count.argtypes = [C.c_void_p]
        for fn in ("orc_method_tool_name", "orc_method_path"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int32]
            getattr(L, fn).restype = C.c_char_p
        for fn in ("orc_method_input", "orc_method_output"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int32]
            getattr(L, fn).restype = C.c_int32
        L.orc_free.argtypes = [C.c_void_p]
        for fn in ("orc_encode", "orc_decode"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t, C.c_uint32,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        L.orc_canon_json.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                     C.c_char_p, C.c_size_t]
        L.orc_request.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(RequestOut)]
        L.orc_request_out_free.argtypes = [C.POINTER(RequestOut)]
        L.orc_response.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                   C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        vp = C.c_void_p
        L.orc_encode_batch.argtypes = [vp, C.c_int64, vp, vp, vp, vp, C.c_uint64, vp, vp, C.c_uint32, C.c_int]
        L.orc_decode_batch.argtypes = L.orc_encode_batch.argtypes
        L.orc_request_batch.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp, vp,
                                        C.c_uint32, C.c_int]
        L.orc_response_batch.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_uint64, vp, vp,
                                         C.c_uint32, C.c_int]
        L.orc_format_float.argtypes = [C.c_double, C.c_int, C.c_char_p, C.c_size_t]
        _lib = L
    return _lib


def _take(ptr, n):
    data = C.string_at(ptr.value, n.value) if ptr.value else b""
    if ptr.value:
        lib().orc_free(ptr)
    return data


class Schema:
    def __init__(self, fds_bytes, naming=0):
        err = C.create_string_buffer(256)
        self.h = lib().orc_schema_new2(fds_bytes, len(fds_bytes), naming, err, 256)
        if not self.h:
            raise ValueError(err.value.decode())
        self._idx = {}

    def msg(self, name):
        if name not in self._idx:
            i = lib().orc_message_index(self.h, name.encode())
            if i < 0:
                raise KeyError(name)
            self._idx[name] = i
        return self._idx[name]

    def methods(self):
        L = lib()
        return [dict(tool=L.orc_method_tool_name(self.h, i).decode(), path=L.orc_method_path(self.h, i).decode(),
                     input=L.orc_method_input(self.h, i), output=L.orc_method_output(self.h, i))
                for i in range(L.orc_method_count(self.h))]

    def _call(self, fn, name, data, flags):
        out, n = C.c_void_p(), C.c_size_t()
        err = C.create_string_buffer(512)
        idx = name if isinstance(name, int) else self.msg(name)
        rc = fn(self.h, idx, data, len(data), flags, C.byref(out), C.byref(n), err, 512)
        return rc, _take(out, n), err.value.decode(errors="replace")

    def encode(self, name, json_bytes, flags=0):
        return self._call(lib().orc_encode, name, json_bytes, flags)

    def decode(self, name, wire, flags=0):
        return self._call(lib().orc_decode, name, wire, flags)

    def request(self, body, flags=0):
        o = RequestOut()
        lib().orc_request(self.h, body, len(body), flags, C.byref(o))
        g = lambda p, n: C.string_at(p, n) if p else b""
        r = dict(kind=o.kind, status=o.status, method=o.method, args=g(o.args, o.args_n), wire=g(o.wire, o.wire_n),
                 id=g(o.id, o.id_n), resp=g(o.resp, o.resp_n))
        lib().orc_request_out_free(C.byref(o))
        return r

    def response(self, name, wire, id_tok, flags=0):
        out, n = C.c_void_p(), C.c_size_t()
        idx = name if isinstance(name, int) else self.msg(name)
        rc = lib().orc_response(self.h, idx, wire, len(wire), id_tok, len(id_tok), flags, C.byref(out), C.byref(n))
        return rc, _take(out, n)

    def _batch(self, fn, msg_ids, data, off, flags, threads, cap):
        n = len(msg_ids)
        msg_ids = np.ascontiguousarray(msg_ids, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(cap, dtype=np.uint8)
        out_off = np.empty(n + 1, dtype=np.uint64)
        status = np.empty(n, dtype=np.int32)
        rc = fn(self.h, n, msg_ids.ctypes.data, data.ctypes.data, off.ctypes.data, out.ctypes.data, cap,
                out_off.ctypes.data, status.ctypes.data, flags, threads)
        if rc != 0:
            raise MemoryError("oracle batch output does not fit (cap=%d)" % cap)
        return out[: int(out_off[n])], out_off, status

    def request_batch(self, bodies, off, threads=1):
        """HTTP request bodies -> (wire bytes, offsets, method[n], id tokens, id offsets, status[n])"""
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        bodies = np.ascontiguousarray(bodies, dtype=np.uint8)
        cap = int(len(bodies) + 64 * n + 64)
        out, out_off = np.empty(cap, np.uint8), np.empty(n + 1, np.uint64)
        ids, ids_off = np.empty(cap, np.uint8), np.empty(n + 1, np.uint64)
        method, status = np.empty(n, np.int32), np.empty(n, np.int32)
        rc = lib().orc_request_batch(self.h, n, bodies.ctypes.data, off.ctypes.data, out.ctypes.data, cap, out_off.ctypes.data,
                                     method.ctypes.data, ids.ctypes.data, cap, ids_off.ctypes.data, status.ctypes.data, 0, threads)
        if rc != 0:
            raise MemoryError("oracle batch output does not fit")
        return out[: int(out_off[n])], out_off, method, ids[: int(ids_off[n])], ids_off, status

    def response_batch(self, msg_ids, data, off, ids, ids_off, flags=0, threads=1, cap=None):
        """reply wire bytes + id tokens -> complete result bodies"""
        n = len(msg_ids)
        msg_ids = np.ascontiguousarray(msg_ids, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        ids = np.ascontiguousarray(ids, dtype=np.uint8)
        ids_off = np.ascontiguousarray(ids_off, dtype=np.uint64)
        cap = int(cap or len(data) * 8 + 256 * n + 64)
        out, out_off, status = np.empty(cap, np.uint8), np.empty(n + 1, np.uint64), np.empty(n, np.int32)
        rc = lib().orc_response_batch(self.h, n, msg_ids.ctypes.data, data.ctypes.data, off.ctypes.data, ids.ctypes.data, ids_off.ctypes.data,
                                      out.ctypes.data, cap, out_off.ctypes.data, status.ctypes.data, flags, threads)
        if rc != 0:
            raise MemoryError("oracle batch output does not fit")
        return out[: int(out_off[n])], out_off, status

    def encode_batch(self, msg_ids, data, off, flags=0, threads=1, cap=None):
        cap = cap or int(len(data) + 64 * len(msg_ids) + 64)
        return self._batch(lib().orc_encode_batch, msg_ids, data, off, flags, threads, cap)

    def decode_batch(self, msg_ids, data, off, flags=0, threads=1, cap=None):
        cap = cap or int(len(data) * 8 + 64 * len(msg_ids) + 64)
        return self._batch(lib().orc_decode_batch, msg_ids, data, off, flags, threads, cap)


def canon_json(b):
    out, n = C.c_void_p(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = lib().orc_canon_json(b, len(b), C.byref(out), C.byref(n), err, 256)
    return rc, _take(out, n)


def format_float(v, bits=64):
    buf = C.create_string_buffer(64)
    lib().orc_format_float(v, bits, buf, 64)
    return buf.value.decode()


def load_schema():
    with open(os.path.join(ROOT, "tests", "golden", "schemas.binpb"), "rb") as fh:
        return Schema(fh.read())
