"""TEST HARNESS ONLY: ctypes front-end of ``emu.cpp`` (host build of the kernel math).

Lets the CPU test-suite check the analytic backward and the fused Adam update of
``csrc/svbrdf_core.cuh`` against the oracle without a GPU.  Never imported by the
product package.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_BUILD = os.path.join(_ROOT, "tests", "_build")
_SO = os.path.join(_BUILD, "libsvbrdf_hostemu.so")
_SRC = [os.path.join(_HERE, "emu.cpp"),
        os.path.join(_ROOT, "svbrdf_diff_renderer_b200", "csrc", "svbrdf_core.cuh")]

_SO_NOISE = os.path.join(_BUILD, "libsvbrdf_hostemu_mufu_noise.so")
_libs = {}


def build(force=False, noise=False):
    """noise=True: every MUFU-approximated function gets a pseudo-random error of up to 2^-22
    (the documented bound of the device units) — checks the tolerances against device-like math."""
    os.makedirs(_BUILD, exist_ok=True)
    so = _SO_NOISE if noise else _SO
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in _SRC):
        return so
    # -ffp-contract=off: fused multiply-adds only where the source says fma()
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC"]
    cmd += (["-DSV_EMU_MUFU_NOISE"] if noise else []) + ["-x", "c++", _SRC[0], "-o", so]
    subprocess.run(cmd, check=True)
    return so


def lib(noise=False):
    if noise not in _libs:
        _libs[noise] = ctypes.CDLL(build(noise=noise))
    return _libs[noise]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


MODE_RENDER, MODE_VJP, MODE_L2, MODE_ADAM, MODE_VJP_L2 = 0, 1, 2, 3, 4


def adam_scalars(step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch/optim/adam.py:531-547 scalar prep, in python doubles like torch."""
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    return np.array([1 - b1, b2, 1 - b2, lr / bc1, 1.0 / (bc2 ** 0.5), eps], dtype=np.float64)


def run(mode, tex, cam, light, power, size, res, io=None, dtype=np.float32, colocated=None, outer_clamp=False,
        n_total=None, row0=0, m=None, v=None, adam=None, noise=False, packed=False):
    """Run the host emulation.  ``tex`` [9,H,W]; returns dict with out / grad_tex / grad_pow / loss.

    In MODE_ADAM ``tex``, ``m``, ``v`` are updated in place (pass arrays of ``dtype``).
    """
    fn = lib(noise).emu_run_f32 if dtype == np.float32 else lib(noise).emu_run_f64
    if packed:                      # T = V2: the generic code instantiated for texel pairs (float only)
        assert dtype == np.float32
        fn = lib(noise).emu_run_v2
    keep_inplace = mode == MODE_ADAM
    tex_a = tex if keep_inplace else np.ascontiguousarray(tex, dtype=dtype)
    assert tex_a.dtype == dtype and tex_a.flags["C_CONTIGUOUS"]
    cam_a = np.ascontiguousarray(cam, dtype=dtype)
    light_a = np.ascontiguousarray(light, dtype=dtype)
    pw_a = np.ascontiguousarray(power, dtype=dtype)
    n = cam_a.shape[0]
    rows, w = tex_a.shape[-2], tex_a.shape[-1]
    if colocated is None:
        colocated = bool(np.array_equal(cam_a, light_a))
    io_a = None if io is None else np.ascontiguousarray(io, dtype=dtype)
    out = np.zeros((n, 3, rows, w), dtype=dtype) if mode == MODE_RENDER else None
    grad_tex = np.zeros((9, rows, w), dtype=dtype) if mode in (MODE_VJP, MODE_L2, MODE_VJP_L2) else None
    grad_pow = np.zeros(3, dtype=dtype) if mode != MODE_RENDER else None
    loss = ctypes.c_double(0.0)
    adam_a = None if adam is None else np.ascontiguousarray(adam, dtype=np.float64)
    fn(ctypes.c_int(int(colocated)), ctypes.c_int(mode), _ptr(tex_a), _ptr(cam_a), _ptr(light_a), _ptr(pw_a),
       ctypes.c_float(size), ctypes.c_int(res), ctypes.c_int(row0), ctypes.c_int(rows), ctypes.c_int(w),
       ctypes.c_int(n), ctypes.c_int(n if n_total is None else n_total), _ptr(io_a), _ptr(out), _ptr(grad_tex),
       _ptr(grad_pow), ctypes.byref(loss), ctypes.c_int(int(outer_clamp)), _ptr(m), _ptr(v), _ptr(adam_a))
    return {"out": out, "grad_tex": grad_tex, "grad_pow": grad_pow, "loss": loss.value}
