"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/svbrdf_b200.h
declares.  No kernel is launched here (argument validation returns before touching the device)."""
import ctypes
import os
import re
import subprocess

import pytest

from svbrdf_diff_renderer_b200 import _native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "svbrdf_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svbrdf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    assert set(_declared()) == set(nv.EXPORTS)


def test_library_builds_and_exports_every_symbol():
    path = nv.build_native()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in include/svbrdf_b200.h but not exported"
    syms = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (svbrdf_[a-z0-9_]+)", syms))
    assert exported == set(_declared()), "exported C symbols differ from the header"


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", nv.build_native()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_abi_version_and_errors():
    L = nv.lib()
    assert L.svbrdf_abi_version() == 1
    assert L.svbrdf_error_string(0) == b"success"
    assert b"bad argument" in L.svbrdf_error_string(-1)
    assert b"unsupported" in L.svbrdf_error_string(-2)
    # 64 epochs x 4096 rows of [loss, dpow x3] partials + 64 finish tickets (+ pad)
    assert L.svbrdf_workspace_bytes(1024, 1024) == 64 * 4096 * 16 + 64 * 4 + 16
    assert L.svbrdf_workspace_bytes(8192, 8192) == 64 * 4096 * 16 + 64 * 4 + 16
    assert L.svbrdf_workspace_bytes(0, 5) == 0


def test_null_arguments_are_rejected_without_touching_the_device():
    L = nv.lib()
    null = ctypes.c_void_p(0)
    assert L.svbrdf_render_fwd(None, null, null, null) == -1
    g = nv.Geom(0, 0, 0, 1.0, 16, 16, 0, 9, 0)
    assert L.svbrdf_render_fwd(ctypes.byref(g), null, null, null) == -1
    g = nv.Geom(8, 8, 8, 1.0, 0, 16, 0, 9, 0)          # res = 0
    assert L.svbrdf_render_fwd(ctypes.byref(g), ctypes.c_void_p(8), ctypes.c_void_p(8), null) == -1
    g = nv.Geom(8, 8, 8, 1.0, 16, 16, 0, 9, 0)
    assert L.svbrdf_render_bwd(ctypes.byref(g), null, null, null, null, null, null) == -1
    assert L.svbrdf_l2_grad(ctypes.byref(g), null, null, 0, 9, null, null, null, null, null) == -1
    assert L.svbrdf_adam_apply(null, null, null, null, 10, None, null) == -1
    a = nv.Adam(0.01, 0.9, 0.999, 1e-8, 0)              # step must be >= 1
    assert L.svbrdf_l2_adam_step(ctypes.byref(g), ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), 0,
                                 ctypes.byref(a), null, null, ctypes.c_void_p(8), null) == -1
    with pytest.raises(RuntimeError, match="bad argument"):
        nv.check(-1, "x")
