"""CPU: the map hand-off oracle (oracle/maps_port.py) against (1) the reference-generated golden vectors,
(2) OpenCV itself — the third-party code the reference calls (imageio.py:75-76) — and, where /root/reference exists,
(3) the live reference round trip through PNG files; plus the library's host-side Lanczos tables."""
import ctypes
import glob
import os

import numpy as np
import pytest

from oracle import maps_port as mp
from svbrdf_diff_renderer_b200 import _native as nv

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "maps_handoff_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_reference_round_trip(path):
    g = np.load(path)
    res_out = int(g["res_out"])
    planes = mp.encode_maps_u8(g["tex"])
    assert np.array_equal(planes, g["planes_u8"]), "encode differs from the bytes the reference wrote"
    up = mp.resize_lanczos4_u8(planes, res_out, res_out)
    assert np.array_equal(mp.decode_maps_u8(up), g["loaded"]), "resize+decode differs from what the reference loaded"
    assert np.array_equal(mp.handoff(g["tex"], res_out), g["loaded"])


def test_four_golden_cases_present():
    assert len(GOLD) == 4


@pytest.mark.parametrize("shape", [(3, 16, 16, 32, 32), (1, 37, 29, 64, 50), (3, 50, 50, 20, 30), (2, 33, 33, 100, 77), (3, 8, 8, 8, 8),
                                   (1, 9, 300, 21, 301), (3, 128, 128, 256, 256)])
def test_resize_bit_identical_to_opencv(shape):
    cv2 = pytest.importorskip("cv2")
    c, h, w, dh, dw = shape
    rng = np.random.default_rng(h * 1000 + w)
    src = rng.integers(0, 256, (c, h, w), dtype=np.uint8)
    src[0, :2] = 255                                     # saturation at the borders (overshoot of the negative lobes)
    src[0, 2:4] = 0
    ref = np.stack([cv2.resize(src[k], (dw, dh), interpolation=cv2.INTER_LANCZOS4) for k in range(c)])
    assert np.array_equal(mp.resize_lanczos4_u8(src, dh, dw), ref)


def test_library_host_tables_match_oracle():
    L = nv.lib()
    for s, d in ((256, 512), (512, 1024), (100, 37), (33, 100), (64, 64), (1, 5)):
        tap = np.empty(d, np.int32)
        co = np.empty((d, 8), np.int16)
        assert L.svbrdf_lanczos4_tables(s, d, tap.ctypes.data_as(ctypes.c_void_p), co.ctypes.data_as(ctypes.c_void_p)) == 0
        o, c = mp.lanczos4_tables(s, d)
        assert np.array_equal(tap, o) and np.array_equal(co, c)
    assert L.svbrdf_lanczos4_tables(0, 4, None, None) == -1
    # a 2x upscale uses two coefficient sets only (fractions 0.75 and 0.25), each summing to ~2048
    _, c = mp.lanczos4_tables(256, 512)
    assert len({tuple(r) for r in c.tolist()}) == 2 and all(abs(int(r.sum()) - 2048) <= 2 for r in c)


def test_encode_decode_edge_cases():
    t = np.zeros((9, 2, 4), np.float32)
    t[:, 0, 0] = -1
    t[:, 0, 1] = 1
    t[3, 0, 2], t[4, 0, 2] = 1, 1            # planar norm > 1: z clamps to 0
    t[:, 1, :] = np.float32(0.3)
    b = mp.encode_maps_u8(t)
    assert b[0, 0, 0] == 0 and b[0, 0, 1] == 255 and b[6, 0, 0] == 0 and b[9, 0, 1] == 255
    assert b[5, 0, 2] == 127                 # z = 0 -> (0+1)/2*255 = 127.5 truncated
    d = mp.decode_maps_u8(b)
    assert d.shape == (9, 2, 4) and np.isfinite(d).all()
    n = d[3:5, 1, 0]
    assert abs(float(n[0]) - 0.3 / np.sqrt(0.09 * 2 + (1 - 0.18))) < 0.01        # unit normal's x after 8-bit rounding


def test_pinned_against_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("needs /root/reference (build container only)")
    pytest.importorskip("cv2")
    import torch as th

    from oracle.make_golden_maps import reference_roundtrip
    from svbrdf_diff_renderer_b200 import synth
    for res, seed in ((64, 21), (96, 22)):
        tex = synth.random_textures(res, seed)
        tex = (tex + th.randn(tex.shape, generator=th.Generator().manual_seed(seed)) * 0.4).clamp(-1, 1)
        planes, loaded = reference_roundtrip(tex, 2 * res)
        assert np.array_equal(mp.encode_maps_u8(tex[0].numpy()), planes)
        assert np.array_equal(mp.handoff(tex[0].numpy(), 2 * res), loaded)
