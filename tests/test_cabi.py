"""The C-ABI shared library loads without a GPU and exports every symbol include/unimatch_sm100.h declares."""
import ctypes
import os
import re

from unimatch_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "unimatch_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(um_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(ops.LIB_PATH)
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(ops.SYMBOLS) == names


def test_abi_version_and_build_info():
    assert ops.LIB.um_abi_version() == 1
    info = ops.build_info()
    assert "sm_100a" in info


def test_bad_arguments_are_reported_without_a_gpu():
    # argument validation happens before any CUDA call, so it can run here
    rc = ops.LIB.um_flow_warp(None, None, None, 1, 4, 4, 2, None)
    assert rc == -22
    assert b"um_flow_warp" in ops.LIB.um_last_error()
    g = ops.AttnGeom(10, 10, 3, 2, 0, 0, 0)      # 10 not divisible by 3
    one = ctypes.c_void_p(16)
    rc = ops.LIB.um_window_attention(one, one, one, one, 2, 0, 128, 128, 128, 128, ctypes.byref(g), None, 0, 0, None)
    assert rc == -22
