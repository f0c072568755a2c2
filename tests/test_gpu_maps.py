"""GPU: device-side map hand-off (svbrdf_maps_encode_u8 / svbrdf_resize_lanczos4_u8 / svbrdf_maps_decode_u8) against the
oracle, the reference-generated golden vectors and OpenCV — bit-exact (byte/integer work; tolerance 0)."""
import glob
import os

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu

import svbrdf_diff_renderer_b200 as pkg  # noqa: E402
from oracle import maps_port as mp  # noqa: E402
from svbrdf_diff_renderer_b200 import maps, synth  # noqa: E402

DEV = th.device("cuda:0")
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "maps_handoff_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_golden_round_trip(path):
    g = np.load(path)
    res_out = int(g["res_out"])
    tex = th.from_numpy(g["tex"]).to(DEV)
    planes = maps.encode_u8(tex, clamp=False)
    assert np.array_equal(planes.cpu().numpy(), g["planes_u8"])
    up = maps.resize_lanczos4_u8(planes, res_out, res_out)
    assert np.array_equal(maps.decode_u8(up)[0].cpu().numpy(), g["loaded"])
    assert np.array_equal(maps.handoff(tex[None], res_out, clamp=False)[0].cpu().numpy(), g["loaded"])


@pytest.mark.parametrize("res", [25, 64, 250])          # 25: 625 texels, not a multiple of 4 -> scalar kernels
def test_encode_decode_against_oracle(res):
    tex = synth.random_textures(res, 5)
    tex = tex + th.randn(tex.shape, generator=th.Generator().manual_seed(res)) * 0.5     # out-of-range values: clamp path
    ref_planes = mp.encode_maps_u8(tex[0].clamp(-1, 1).numpy())
    planes = maps.encode_u8(tex.to(DEV), clamp=True)
    assert np.array_equal(planes.cpu().numpy(), ref_planes)
    rnd = np.random.default_rng(res).integers(0, 256, (10, res, res), dtype=np.uint8)
    assert np.array_equal(maps.decode_u8(th.from_numpy(rnd).to(DEV))[0].cpu().numpy(), mp.decode_maps_u8(rnd))


@pytest.mark.parametrize("shape", [(10, 24, 24, 48, 48), (1, 37, 29, 64, 50), (3, 50, 50, 20, 30), (2, 33, 33, 100, 77), (3, 8, 8, 8, 8),
                                   (1, 9, 300, 21, 301), (4, 130, 70, 131, 260)])
def test_resize_against_oracle(shape):
    c, h, w, dh, dw = shape
    src = np.random.default_rng(h * 1000 + w).integers(0, 256, (c, h, w), dtype=np.uint8)
    src[0, :2] = 255
    src[0, 2:4] = 0
    out = maps.resize_lanczos4_u8(th.from_numpy(src).to(DEV), dh, dw)
    assert np.array_equal(out.cpu().numpy(), mp.resize_lanczos4_u8(src, dh, dw))


def test_full_size_schedule_against_opencv():
    """The reference's schedule 256 -> 512 -> 1024 (run.py:55-56) at full size: the resize against OpenCV (what
    imageio.py:75-76 calls), the float steps against the oracle."""
    cv2 = pytest.importorskip("cv2")
    tex = synth.random_textures(256, 9).to(DEV)
    for res in (512, 1024):
        planes = maps.encode_u8(tex)
        host = planes.cpu().numpy()
        assert np.array_equal(host, mp.encode_maps_u8(tex[0].clamp(-1, 1).cpu().numpy()))
        up = maps.resize_lanczos4_u8(planes, res, res)
        ref = np.stack([cv2.resize(host[k], (res, res), interpolation=cv2.INTER_LANCZOS4) for k in range(10)])
        assert np.array_equal(up.cpu().numpy(), ref)
        nxt = maps.decode_u8(up)
        assert np.array_equal(nxt[0].cpu().numpy(), mp.decode_maps_u8(ref))
        assert th.equal(nxt, maps.handoff(tex, res))
        tex = nxt
    # size-independent properties: a same-size resize is a copy (cv::resize), re-quantising decoded maps moves a byte
    # by at most one level and never up (truncation), and the decoded normals lie inside the unit disc
    up1k = maps.encode_u8(tex)
    assert th.equal(maps.resize_lanczos4_u8(up1k, 1024, 1024), up1k)
    again = maps.encode_u8(maps.decode_u8(up1k))
    keep = [0, 1, 2, 6, 7, 8, 9]                                 # the normal planes are renormalised by the decoder
    d = up1k[keep].to(th.int16) - again[keep].to(th.int16)
    assert int(d.min()) >= 0 and int(d.max()) <= 1
    assert float((tex[0, 3] ** 2 + tex[0, 4] ** 2).max()) <= 1.0 + 1e-6


def test_svbrdfio_device_paths_write_and_read_the_same_files(tmp_path):
    """SvbrdfIO.save_textures_th / load_textures_th on a CUDA device go through the native codec; the files and the
    loaded maps are identical to the host (reference-behaviour) path."""
    import json
    cfg = {"reference_dir": "ref", "target_dir": "t", "optimize_dir": "o", "rerender_dir": "r"}
    (tmp_path / "a.json").write_text(json.dumps(cfg))
    io = pkg.SvbrdfIO(tmp_path / "a.json", DEV)
    tex = synth.random_textures(64, 3).to(DEV).clamp(-1, 1)
    io.save_textures_th(tex, tmp_path / "dev")
    io.save_textures_th(tex.cpu(), tmp_path / "host")
    for name in ("nom.png", "dif.png", "spe.png", "rgh.png", "tex.png"):
        assert (tmp_path / "dev" / name).read_bytes() == (tmp_path / "host" / name).read_bytes(), name
    for res in (64, 128, 100):
        a = io.load_textures_th(tmp_path / "dev", res)
        b = io.load_textures_th(tmp_path / "dev", res, on_device=False)
        assert a.is_cuda and th.equal(a, b.to(DEV)), res


def test_errors():
    with pytest.raises(RuntimeError, match="CUDA"):
        maps.encode_u8(th.zeros(1, 9, 8, 8))
    with pytest.raises(RuntimeError):
        maps.decode_u8(th.zeros(9, 8, 8, dtype=th.uint8, device=DEV))
    with pytest.raises(RuntimeError):
        maps.resize_lanczos4_u8(th.zeros(3, 8, 8, device=DEV), 16, 16)


def test_pyramid_on_device_equals_the_file_based_schedule(tmp_path):
    """run.py:55-56 in miniature (32 -> 64 -> 128): optim_perpixel_pyramid keeps the hand-off on the GPU; the reference
    recipe passes the maps through optimize_dir PNGs and tex_init="textures".  Same losses, same maps, bit for bit."""
    import json
    n = 9
    cl = synth.calibration(n)
    base = {"im_size": synth.IM_SIZE_CM, "idx": list(range(n)), "camera_pos": cl[0].tolist(), "light_pos": cl[1].tolist(),
            "light_pow": list(synth.LIGHT_POW)}
    # targets: one set of ground-truth maps rendered at 128 and saved; every stage loads them resized (imageio.py:14-15)
    (tmp_path / "gt.json").write_text(json.dumps(dict(base, reference_dir="gt/maps", target_dir="target", optimize_dir="x", rerender_dir="x/r")))
    io = pkg.SvbrdfIO(tmp_path / "gt.json", DEV)
    io.save_textures_th(synth.random_textures(128, 1).to(DEV).clamp(-1, 1), io.reference_dir)
    pkg.render(tmp_path / "gt.json", 128)
    stages_a, stages_b = [], []
    prev_dir = "none"
    for res in (32, 64, 128):
        for tag, stages in (("a", stages_a), ("b", stages_b)):
            jp = tmp_path / f"{tag}{res}.json"
            jp.write_text(json.dumps(dict(base, reference_dir=(prev_dir if tag == "b" else "unused"), target_dir="target",
                                          optimize_dir=f"{tag}/{res}", rerender_dir=f"{tag}/{res}/rerender")))
            stages.append((jp, res))
        prev_dir = f"b/{res}"
    outs_a = pkg.optim_perpixel_pyramid(stages_a, 0.02, 6, "const")
    outs_b = []
    for k, (jp, res) in enumerate(stages_b):
        outs_b.append(pkg.optim_perpixel(jp, res, 0.02, 6, "const" if k == 0 else "textures"))
    for a, b in zip(outs_a, outs_b):
        assert a.losses == b.losses
        assert th.equal(a.textures, b.textures)
    assert outs_a[-1].textures.shape == (1, 9, 128, 128) and outs_a[-1].losses[-1] < outs_a[-1].losses[0]
