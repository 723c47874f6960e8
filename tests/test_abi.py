"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/ggrmcp_b200.h declares, and refuses to run without a CUDA device (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ggrmcp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ggr_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    import ggrmcp_b200
    path = ggrmcp_b200.lib_path()
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def test_status_strings():
    import ggrmcp_b200
    lib = ctypes.CDLL(ggrmcp_b200.lib_path())
    lib.ggr_status_string.restype = ctypes.c_char_p
    assert lib.ggr_status_string(0) == b"ok"
    assert lib.ggr_status_string(2) == b"unknown_field"
    assert [lib.ggr_status_string(i).decode() for i in range(14)] == ggrmcp_b200.STATUS_NAMES


def test_no_cpu_fallback():
    import torch
    import ggrmcp_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ggrmcp_b200.EngineError):
        ggrmcp_b200.Engine(0)


def test_product_does_not_touch_oracle():
    """the product path must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "ggrmcp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("#include", "import ", "from ")):
                        assert "oracle" not in s and "orc" not in s.split(), (f, s)
                assert "libggr_oracle" not in text and "hostsim" not in text.replace("tests/hostsim", ""), f
