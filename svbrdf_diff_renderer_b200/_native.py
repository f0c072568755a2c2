"""ctypes binding of ``libsvbrdf_b200.so`` (C ABI: ``include/svbrdf_b200.h``).

The library is built in-tree by ``build_native()`` (``nvcc -gencode
arch=compute_100a,code=sm_100a``) and loaded from
``svbrdf_diff_renderer_b200/csrc/``.  There is no CPU fallback: if the library is
missing, or a tensor is not a contiguous CUDA fp32 tensor, the call raises.
PyTorch is used for device memory and streams only.
"""

from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

import torch as th

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(CSRC, "libsvbrdf_b200.so")
SOURCES = [os.path.join(CSRC, "svbrdf_kernels.cu"), os.path.join(CSRC, "svbrdf_maps.cu")]
HEADERS = [os.path.join(CSRC, "svbrdf_core.cuh"), os.path.join(CSRC, "svbrdf_host.h"), os.path.join(_ROOT, "include", "svbrdf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # no implicit mul+add contraction: every FMA in the kernels is written as one (Fm<T>::fma), so the forward pass is
    # rounded identically in every kernel (a rendered target is reproduced bit for bit by the L2 forward) and the host
    # emulation in tests/hostemu follows the same arithmetic; costs 5 of 1254 instructions per tile
    "-fmad=false",
    "--shared", "-Xcompiler", "-fPIC",
    "-diag-suppress", "128",
]

TARGET_F32, TARGET_U8, TARGET_F16 = 0, 1, 2

EXPORTS = (
    "svbrdf_abi_version", "svbrdf_error_string", "svbrdf_workspace_bytes", "svbrdf_render_fwd", "svbrdf_render_bwd",
    "svbrdf_l2_grad", "svbrdf_l2_adam_step", "svbrdf_l2_adam_run", "svbrdf_adam_apply",
    "svbrdf_l2_grad_push", "svbrdf_reduce_adam_push",
    "svbrdf_maps_encode_u8", "svbrdf_maps_decode_u8", "svbrdf_lanczos4_tables", "svbrdf_resize_lanczos4_u8",
    "svbrdf_render_norm_l2_fwd", "svbrdf_render_norm_l2_bwd",
)


class Geom(ctypes.Structure):
    """``svbrdf_geom_t``."""
    _fields_ = [
        ("camera_pos", ctypes.c_void_p), ("light_pos", ctypes.c_void_p), ("light_pow", ctypes.c_void_p),
        ("size", ctypes.c_float), ("res", ctypes.c_int32), ("rows", ctypes.c_int32), ("row_offset", ctypes.c_int32),
        ("n_lights", ctypes.c_int32), ("plane_stride", ctypes.c_int64),
    ]


class Peers(ctypes.Structure):
    """``svbrdf_peers_t``."""
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("chunk", ctypes.c_int64),
                ("recv", ctypes.c_void_p * 8), ("tex", ctypes.c_void_p * 8), ("tex_multicast", ctypes.c_void_p),
                ("pull_tex", ctypes.c_int32)]


class Adam(ctypes.Structure):
    """``svbrdf_adam_t``."""
    _fields_ = [("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double), ("eps", ctypes.c_double),
                ("step", ctypes.c_int64)]


def nvcc_path() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if os.path.exists(cand):
        return cand
    found = shutil.which("nvcc")
    if not found:
        raise RuntimeError("nvcc not found: cannot build libsvbrdf_b200.so")
    return found


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library (building it if the sources are newer).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("SVBRDF_B200_LIB")        # development: an alternative build of the same ABI
    if override:
        L = ctypes.CDLL(override)
    else:
        if needs_build():
            build_native()
        L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    gp, ap = ctypes.POINTER(Geom), ctypes.POINTER(Adam)
    L.svbrdf_abi_version.restype = ctypes.c_int
    L.svbrdf_abi_version.argtypes = []
    L.svbrdf_error_string.restype = ctypes.c_char_p
    L.svbrdf_error_string.argtypes = [ctypes.c_int]
    L.svbrdf_workspace_bytes.restype = ctypes.c_size_t
    L.svbrdf_workspace_bytes.argtypes = [i32, i32]
    L.svbrdf_render_fwd.argtypes = [gp, vp, vp, vp]
    L.svbrdf_render_bwd.argtypes = [gp, vp, vp, vp, vp, vp, vp]
    L.svbrdf_l2_grad.argtypes = [gp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.svbrdf_l2_adam_step.argtypes = [gp, vp, vp, vp, vp, i32, ap, vp, vp, vp, vp]
    L.svbrdf_l2_adam_run.argtypes = [gp, vp, vp, vp, vp, i32, ap, i32, vp, vp, vp, vp]
    L.svbrdf_adam_apply.argtypes = [vp, vp, vp, vp, ctypes.c_size_t, ap, vp]
    pp = ctypes.POINTER(Peers)
    L.svbrdf_l2_grad_push.argtypes = [gp, vp, vp, i32, i32, pp, vp, vp, vp]
    L.svbrdf_reduce_adam_push.argtypes = [pp, i64, vp, vp, ap, vp]
    f3 = ctypes.POINTER(ctypes.c_float)
    L.svbrdf_render_norm_l2_fwd.argtypes = [gp, vp, f3, f3, vp, i32, vp, vp, vp, vp]
    L.svbrdf_render_norm_l2_bwd.argtypes = [gp, vp, f3, vp, vp, i32, vp, vp, vp, vp, vp]
    L.svbrdf_maps_encode_u8.argtypes = [vp, i64, i32, i32, i32, vp, vp]
    L.svbrdf_maps_decode_u8.argtypes = [vp, i32, i32, vp, i64, vp]
    L.svbrdf_lanczos4_tables.argtypes = [i32, i32, vp, vp]          # host pointers (numpy buffers)
    L.svbrdf_resize_lanczos4_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp]
    for name in EXPORTS[3:]:
        getattr(L, name).restype = ctypes.c_int
    if L.svbrdf_abi_version() != 1:
        raise RuntimeError("libsvbrdf_b200.so: ABI version mismatch")
    _lib = L
    return L


def check(code: int, what: str):
    if code != 0:
        msg = lib().svbrdf_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def dev_f32(t: th.Tensor, what: str) -> th.Tensor:
    """The native path only takes contiguous CUDA fp32 tensors — no silent CPU route."""
    if not isinstance(t, th.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (svbrdf_diff_renderer_b200 has no CPU path)")
    if t.dtype != th.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{what}: expected a contiguous tensor")
    return t


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(th.cuda.current_stream(device).cuda_stream)


def ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def workspace(res: int, rows: int, device) -> th.Tensor:
    nbytes = lib().svbrdf_workspace_bytes(res, rows)
    # zero-initialised once: the trailing finish counter must be 0 before the first launch (kernels leave it 0)
    return th.zeros(max(nbytes // 4, 4), dtype=th.float32, device=device)


def target_dtype_code(t: th.Tensor) -> int:
    if t.dtype == th.float32:
        return TARGET_F32
    if t.dtype == th.uint8:
        return TARGET_U8
    if t.dtype == th.float16:
        return TARGET_F16
    raise RuntimeError(f"targets: unsupported dtype {t.dtype} (float32, float16 or uint8)")
