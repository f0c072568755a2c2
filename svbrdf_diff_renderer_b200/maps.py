"""Texture-map hand-off between resolutions, on the device.

The reference carries maps from one ``optim_perpixel`` call to the next (256 -> 512 -> 1024,
``/root/reference/run.py:55-56``) through 8-bit PNG files and ``cv2.resize(INTER_LANCZOS4)`` on the host
(``SvbrdfIO.save_textures_th`` / ``load_textures_th``, ``/root/reference/src/svbrdf.py:150-189``;
``/root/reference/src/imageio.py:11-76``).  The functions here are that round trip without the files and without the
host: three native kernels (``svbrdf_maps_encode_u8`` -> ``svbrdf_resize_lanczos4_u8`` -> ``svbrdf_maps_decode_u8``)
that reproduce it bit for bit (tests/test_gpu_maps.py against ``oracle/maps_port.py`` and reference-generated vectors).

Byte planes are planar RGB order ``[dif r,g,b | nom x,y,z | rgh | spe r,g,b]``; ``planes_to_png_arrays`` /
``png_arrays_to_planes`` convert to and from what ``cv2.imwrite`` / ``cv2.imread`` handle (interleaved BGR).
No CPU fallback: CPU tensors raise.
"""

from __future__ import annotations

import ctypes
import functools

import numpy as np
import torch as th

from . import _native as nv

PLANES = 10


def _dev_u8(t: th.Tensor, what: str) -> th.Tensor:
    if not isinstance(t, th.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (svbrdf_diff_renderer_b200 has no CPU path)")
    if t.dtype != th.uint8 or not t.is_contiguous():
        raise RuntimeError(f"{what}: expected a contiguous uint8 tensor")
    return t


@functools.lru_cache(maxsize=64)
def _host_tables(src: int, dst: int):
    """cv2's Lanczos-4 fixed-point tables for one axis, computed by the library on the host (cached per size pair)."""
    tap = np.empty(dst, np.int32)
    coef = np.empty((dst, 8), np.int16)
    nv.check(nv.lib().svbrdf_lanczos4_tables(src, dst, tap.ctypes.data_as(ctypes.c_void_p), coef.ctypes.data_as(ctypes.c_void_p)),
             "svbrdf_lanczos4_tables")
    return tap, coef


_dev_tables = {}


def _tables(src: int, dst: int, device):
    key = (src, dst, str(device))
    if key not in _dev_tables:
        tap, coef = _host_tables(src, dst)
        _dev_tables[key] = (th.from_numpy(tap).to(device), th.from_numpy(coef).to(device))
    return _dev_tables[key]


def encode_u8(textures: th.Tensor, clamp: bool = True) -> th.Tensor:
    """``[1,9,r,c]`` or ``[9,r,c]`` float32 parameters -> ``[10,r,c]`` uint8: the bytes ``save_textures_th`` writes
    (svbrdf.py:168-184, imageio.py:52-71).  ``clamp`` applies the caller's ``textures.clamp(-1,1)`` (scripts.py:91)."""
    t = textures[0] if textures.dim() == 4 else textures
    t = nv.dev_f32(t.detach(), "textures")
    if t.dim() != 3 or t.shape[0] != 9:
        raise RuntimeError(f"textures must be [9,rows,cols], got {tuple(t.shape)}")
    out = th.empty((PLANES, t.shape[1], t.shape[2]), dtype=th.uint8, device=t.device)
    nv.check(nv.lib().svbrdf_maps_encode_u8(nv.ptr(t), 0, t.shape[1], t.shape[2], 1 if clamp else 0, nv.ptr(out), nv.stream_ptr(t.device)),
             "svbrdf_maps_encode_u8")
    return out


def decode_u8(planes: th.Tensor) -> th.Tensor:
    """``[10,r,c]`` uint8 -> ``[1,9,r,c]`` float32 parameters: what ``load_textures_th`` returns (svbrdf.py:150-166)."""
    b = _dev_u8(planes, "planes")
    if b.dim() != 3 or b.shape[0] != PLANES:
        raise RuntimeError(f"planes must be [10,rows,cols], got {tuple(b.shape)}")
    out = th.empty((1, 9, b.shape[1], b.shape[2]), dtype=th.float32, device=b.device)
    nv.check(nv.lib().svbrdf_maps_decode_u8(nv.ptr(b), b.shape[1], b.shape[2], nv.ptr(out), 0, nv.stream_ptr(b.device)), "svbrdf_maps_decode_u8")
    return out


def resize_lanczos4_u8(planes: th.Tensor, rows: int, cols: int) -> th.Tensor:
    """``cv2.resize(plane, (cols, rows), interpolation=cv2.INTER_LANCZOS4)`` on every plane of ``[P,h,w]`` uint8
    (imageio.py:75-76), bit-identical to OpenCV."""
    b = _dev_u8(planes, "planes")
    if b.dim() != 3:
        raise RuntimeError(f"planes must be [P,rows,cols], got {tuple(b.shape)}")
    p, h, w = b.shape
    out = th.empty((p, rows, cols), dtype=th.uint8, device=b.device)
    xt, xc = _tables(w, cols, b.device)
    yt, yc = _tables(h, rows, b.device)
    nv.check(nv.lib().svbrdf_resize_lanczos4_u8(nv.ptr(b), p, h, w, nv.ptr(out), rows, cols, nv.ptr(xt), nv.ptr(xc), nv.ptr(yt), nv.ptr(yc),
                                                nv.stream_ptr(b.device)), "svbrdf_resize_lanczos4_u8")
    return out


def handoff(textures: th.Tensor, res: int, clamp: bool = True) -> th.Tensor:
    """``save_textures_th(textures.clamp(-1,1), d); load_textures_th(d, res)`` on the device: ``[1,9,r,r]`` ->
    ``[1,9,res,res]``, the initial maps of the next resolution (run.py:55-56)."""
    return decode_u8(resize_lanczos4_u8(encode_u8(textures, clamp), res, res))


def planes_to_png_arrays(planes: th.Tensor):
    """Device byte planes -> the four host arrays ``cv2.imwrite`` takes (BGR interleaved; roughness single channel)."""
    b = planes.cpu().numpy()
    bgr = lambda a: np.ascontiguousarray(a[::-1].transpose(1, 2, 0))  # noqa: E731
    return {"dif": bgr(b[0:3]), "nom": bgr(b[3:6]), "rgh": np.ascontiguousarray(b[6]), "spe": bgr(b[7:10])}


def png_arrays_to_planes(arrays, device) -> th.Tensor:
    """Inverse of ``planes_to_png_arrays`` for decoded 8-bit PNGs (``cv2.imread`` output)."""
    rgb = lambda a: a[:, :, ::-1].transpose(2, 0, 1)  # noqa: E731
    rgh = arrays["rgh"]
    if rgh.ndim == 3:
        raise RuntimeError("rgh.png: 3-channel roughness goes through the host path (imageio.imread 'rough' averages in float)")
    b = np.concatenate([rgb(arrays["dif"]), rgb(arrays["nom"]), rgh[None], rgb(arrays["spe"])], 0)
    return th.from_numpy(np.ascontiguousarray(b)).to(device)
