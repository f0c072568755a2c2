"""ORACLE (test infrastructure, not product code): CPU restatement of the texture-map hand-off between
resolutions — the step on either side of ``optim_perpixel``.

The reference moves maps from one resolution to the next through 8-bit PNG files (run.py:55-56):

    save_textures_th   /root/reference/src/svbrdf.py:168-189   [1,9,r,r] -> nom/dif/spe/rgh.png
      imwrite          /root/reference/src/imageio.py:52-71    clip, BGR, (im*255).astype("uint8")  (truncation)
    load_textures_th   /root/reference/src/svbrdf.py:150-166   PNG -> resize to (res,res) -> [1,9,res,res]
      imread/imresize  /root/reference/src/imageio.py:11-76    cv2.resize(INTER_LANCZOS4) ON THE BYTES, /255, BGR->RGB,
                                                               "normal": *2-1 and renormalise

Third-party arithmetic: ``cv2.resize(..., INTER_LANCZOS4)`` on 8-bit data — OpenCV (container: 4.13.0),
modules/imgproc/src/resize.cpp: 8-tap separable filter, coefficients from ``interpolateLanczos4`` (double sin/cos,
float32 normalisation) rounded to 11-bit fixed point, int32 accumulation in both passes, one rounding shift by 22 bits,
replicated borders.  ``resize_lanczos4_u8`` below restates that algorithm; tests/test_oracle_maps.py pins it bit for
bit against cv2 itself and against the reference's save->load round trip (where /root/reference exists), and
tests/golden/maps_handoff_*.npz hold reference-generated vectors for the GPU box.

Plain numpy float32 with explicit operation order (no FMA): the CUDA kernels follow the same order.
Byte planes are planar RGB order: [dif r,g,b | nom x,y,z | rgh | spe r,g,b].
"""

from __future__ import annotations

import math

import numpy as np

F = np.float32
PLANES = 10


# ------------------------------------------------------------------------------------------------
# encode: svbrdf.py:168-184 + imageio.py:52-71
# ------------------------------------------------------------------------------------------------
def _quant(x01):
    """imageio.py:70: (im * 255).astype("uint8") after the flag's clip — truncation toward zero."""
    return (x01 * F(255)).astype(np.uint8)


def encode_maps_u8(tex):
    """``tex`` [9,r,r] float32 in the parameter range -> [10,r,r] uint8 (what the four PNGs hold, RGB order)."""
    t = np.asarray(tex, dtype=F)
    assert t.ndim == 3 and t.shape[0] == 9
    out = np.empty((PLANES,) + t.shape[1:], np.uint8)
    half = lambda a: ((a + F(1)) / F(2)).clip(F(0), F(1))                # svbrdf.py:171,173,174 + imageio.py:59,63
    out[0:3] = _quant(half(t[0:3]))
    out[6] = _quant(half(t[5]))
    out[7:10] = _quant(half(t[6:9]))
    # SvbrdfIO.reconstruct_normal, svbrdf.py:102-108 (clamp to 1, not 1-eps, unlike the renderer's)
    x, y = t[3].clip(F(-1), F(1)), t[4].clip(F(-1), F(1))
    z = np.sqrt(F(1) - (x * x + y * y).clip(F(0), F(1)))
    norm = np.sqrt(x * x + y * y + z * z)
    for k, c in enumerate((x, y, z)):
        n = c / norm
        out[3 + k] = _quant((n.clip(F(-1), F(1)) + F(1)) / F(2))          # imageio.py:66
    return out


# ------------------------------------------------------------------------------------------------
# cv2.resize(INTER_LANCZOS4) on uint8
# ------------------------------------------------------------------------------------------------
def lanczos4_coeffs(x):
    """cv::interpolateLanczos4: float32 fraction in, 8 float32 weights out (double trig, float32 normalisation)."""
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    x = F(x)
    c = np.zeros(8, F)
    if x < np.finfo(F).eps:
        c[3] = 1
        return c
    y0 = -(float(x) + 3) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    s = F(0)
    for i in range(8):
        y = -(float(x) + 3 - i) * math.pi * 0.25
        c[i] = F((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        s = F(s + c[i])
    return (c * (F(1) / s)).astype(F)


def lanczos4_tables(ssize, dsize):
    """Per destination index: first-tap source index (may be out of range: taps are clamped) and the 8 fixed-point
    weights ``saturate_cast<short>(w * 2048)`` (INTER_RESIZE_COEF_BITS = 11)."""
    scale = float(ssize) / dsize
    ofs = np.zeros(dsize, np.int32)
    co = np.zeros((dsize, 8), np.int16)
    for d in range(dsize):
        fx = F((d + 0.5) * scale - 0.5)
        sx = int(math.floor(fx))
        ofs[d] = sx - 3
        co[d] = np.rint(lanczos4_coeffs(F(fx - F(sx))) * F(2048)).clip(-32768, 32767).astype(np.int16)
    return ofs, co


def resize_lanczos4_u8(src, dh, dw):
    """``src`` [C,h,w] uint8 -> [C,dh,dw] uint8, bit-identical to cv2.resize(INTER_LANCZOS4) per plane."""
    src = np.asarray(src, np.uint8)
    c, h, w = src.shape
    if (dh, dw) == (h, w):
        return src.copy()
    xo, xa = lanczos4_tables(w, dw)
    yo, ya = lanczos4_tables(h, dh)
    taps = np.arange(8)
    ix = np.clip(xo[:, None] + taps[None, :], 0, w - 1)                   # [dw,8] replicated border
    iy = np.clip(yo[:, None] + taps[None, :], 0, h - 1)
    s = src.astype(np.int64)
    rows = (s[:, :, ix] * xa.astype(np.int64)[None, None]).sum(3).astype(np.int32).astype(np.int64)   # [C,h,dw] int32 (wraps like C)
    acc = (rows[:, iy, :] * ya.astype(np.int64)[None, :, :, None]).sum(2).astype(np.int32).astype(np.int64)
    return np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)     # FixedPtCast<int, uchar, 22>


# ------------------------------------------------------------------------------------------------
# decode: imageio.py:11-49 + svbrdf.py:155-166
# ------------------------------------------------------------------------------------------------
def decode_maps_u8(b):
    """[10,R,R] uint8 -> [9,R,R] float32 textures (diffuse | normal xy | roughness | specular)."""
    b = np.asarray(b, np.uint8)
    f = b.astype(F) / F(255)                                              # imageio.py:18-19
    out = np.empty((9,) + b.shape[1:], F)
    out[0:3] = f[0:3] * F(2) - F(1)                                       # svbrdf.py:161
    out[5] = f[6] * F(2) - F(1)
    out[6:9] = f[7:10] * F(2) - F(1)
    n = f[3:6] * F(2) - F(1)                                              # imageio.py:44-47
    norm = np.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2])
    out[3], out[4] = n[0] / norm, n[1] / norm
    return out


def handoff(tex, res_out):
    """save_textures_th(tex) -> load_textures_th(dir, res_out) without the files."""
    return decode_maps_u8(resize_lanczos4_u8(encode_maps_u8(tex), res_out, res_out))
