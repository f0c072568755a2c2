"""CPU: host-side logic of the drop-in package — no kernels are launched."""
import json

import numpy as np
import pytest
import torch as th

import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import synth
from svbrdf_diff_renderer_b200.imageio import imread, imwrite


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Microfacet(16, 9, synth.IM_SIZE_CM, synth.calibration(9), th.device("cpu"))
    if not th.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pkg.optim_perpixel(None, 16, 0.01, 1, "random")


def test_cpu_tensors_are_rejected_by_the_native_binding():
    from svbrdf_diff_renderer_b200 import _native as nv
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        nv.dev_f32(th.zeros(4), "x")
    with pytest.raises(RuntimeError, match="unsupported dtype"):
        nv.target_dtype_code(th.zeros(2, dtype=th.float64))
    assert nv.target_dtype_code(th.zeros(2, dtype=th.uint8)) == nv.TARGET_U8


def test_synthetic_inputs_are_reproducible_and_in_range():
    a, b = synth.random_textures(32, 4), synth.random_textures(32, 4)
    assert th.equal(a, b) and a.shape == (1, 9, 32, 32) and a.dtype == th.float32
    assert float(a.min()) >= -1 and float(a.max()) <= 1
    assert not th.equal(a, synth.random_textures(32, 5))
    for n in (9, 64, 256):
        cl = synth.calibration(n)
        assert cl[0].shape == (n, 3) and th.equal(cl[0], cl[1]) and cl[2].tolist() == [1500.0] * 3
    off = synth.calibration(9, colocated=False)
    assert not th.equal(off[0], off[1]) and float(off[1][:, 2].min()) >= 4.0
    with pytest.raises(ValueError):
        synth.grid_lights(10)


def _write_config(tmp_path, n=9):
    cl = synth.calibration(n)
    cfg = {"reference_dir": "target/maps", "target_dir": "target", "optimize_dir": "optim", "rerender_dir": "optim/rerender",
           "im_size": synth.IM_SIZE_CM, "idx": list(range(n)), "camera_pos": cl[0].tolist(), "light_pos": cl[1].tolist(),
           "light_pow": list(synth.LIGHT_POW)}
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(cfg))
    return p


def test_svbrdfio_config_and_png_round_trip(tmp_path):
    io = pkg.SvbrdfIO(_write_config(tmp_path), th.device("cpu"))
    assert io.n_of_imgs == 9 and io.im_size == synth.IM_SIZE_CM
    assert th.equal(io.cl[0], synth.calibration(9)[0]) and io.cl[2].tolist() == [1500.0] * 3

    # images: float32 and uint8 ingest agree exactly (x/255)
    imgs = th.rand(9, 3, 16, 16)
    io.save_images_th(imgs, io.target_dir)
    assert (io.target_dir / "all.png").exists()
    f32 = io.load_images_th(io.target_dir, 16)
    u8 = io.load_images_th(io.target_dir, 16, as_uint8=True)
    assert f32.dtype == th.float32 and u8.dtype == th.uint8 and f32.shape == (9, 3, 16, 16)
    assert th.equal(u8.float() / 255, f32)
    assert float((f32 - imgs).abs().max()) <= 1 / 255 + 1e-6       # 8-bit quantisation (truncating, imageio.py:71)

    # textures: save -> load returns the quantised maps in the [1,9,R,R] channel order
    tex = synth.random_textures(16, 3)
    io.save_textures_th(tex, io.reference_dir)
    for name in ("nom.png", "dif.png", "spe.png", "rgh.png", "tex.png"):
        assert (io.reference_dir / name).exists()
    back = io.load_textures_th(io.reference_dir, 16)
    assert back.shape == (1, 9, 16, 16)
    assert float((back[:, [0, 1, 2, 5, 6, 7, 8]] - tex[:, [0, 1, 2, 5, 6, 7, 8]]).abs().max()) <= 2 / 255 + 1e-6
    assert float((back[:, 3:5] - tex[:, 3:5]).abs().max()) <= 0.02

    with pytest.raises(FileNotFoundError):
        pkg.SvbrdfIO(tmp_path / "missing.json", th.device("cpu"))


def test_imageio_flags(tmp_path):
    im = np.random.default_rng(0).random((8, 8, 3)).astype("float32")
    imwrite(im, tmp_path / "a.png", "srgb")
    back = imread(tmp_path / "a.png", "srgb")
    assert back.shape == (8, 8, 3) and np.abs(back - im).max() <= 1 / 255 + 1e-6
    rough = imread(tmp_path / "a.png", "rough")
    assert rough.shape == (8, 8)
    nrm = imread(tmp_path / "a.png", "normal")
    assert np.allclose(np.linalg.norm(nrm, axis=2), 1, atol=1e-5)
    up = imread(tmp_path / "a.png", "srgb", (16, 16))
    assert up.shape == (16, 16, 3)


def test_optim_base_class_api():
    base = pkg.Optim(th.device("cpu"), None)
    leaf = base.gradient(th.zeros(3))
    assert leaf.requires_grad and leaf.is_leaf
    lst = base.gradient([th.zeros(2), th.ones(2)])
    assert all(t.requires_grad for t in lst)
    with pytest.raises(NotImplementedError):
        base.load_targets(None)
    with pytest.raises(NotImplementedError):
        base.optim(1, 0.1, None, False)


def test_vggloss_drop_in_matches_oracle_on_cpu():
    """descriptor.VGGLoss is plain torch: on the CPU it must agree with the oracle's op-for-op restatement of the
    reference class (descriptor.py:7-79) — broadcast normalisation == per-image loop, same hooks, same weighting."""
    import pytest
    pytest.importorskip("torchvision")
    import torch as th
    from torchvision.models import vgg19

    from oracle import descriptor_port as dp
    from svbrdf_diff_renderer_b200.descriptor import VGGLoss
    th.manual_seed(11)
    mine = VGGLoss(th.device("cpu"), net=vgg19(weights=None).features)
    net = dp.seeded_vgg_features(11)
    x = th.rand(2, 3, 32, 32, generator=th.Generator().manual_seed(1))
    y = th.rand(2, 3, 32, 32, generator=th.Generator().manual_seed(2))
    assert th.equal(mine.normalize(x), dp.normalize(x))
    mine.load(y)
    ref = float(dp.vgg_loss(net, x, dp.feature_vector(net, dp.normalize(y))))
    assert float(mine(x)) == pytest.approx(ref, rel=1e-6)
    assert float(mine.forward_normalized(mine.normalize(x))) == pytest.approx(ref, rel=1e-6)
    assert mine.compute_feature_vector(mine.normalize(x), is_gram=True).numel() == sum((2 * c) ** 2 for c in (64, 64, 256, 512))


def test_map_plane_layout_round_trip():
    """maps.planes_to_png_arrays / png_arrays_to_planes: planar RGB byte planes <-> the BGR-interleaved arrays cv2 handles."""
    import numpy as np
    import torch as th

    from svbrdf_diff_renderer_b200 import maps
    planes = th.from_numpy(np.random.default_rng(0).integers(0, 256, (10, 6, 5), dtype=np.uint8))
    arrays = maps.planes_to_png_arrays(planes)
    assert arrays["dif"].shape == (6, 5, 3) and arrays["rgh"].shape == (6, 5)
    assert np.array_equal(arrays["dif"][:, :, 0], planes[2].numpy()) and np.array_equal(arrays["nom"][:, :, 2], planes[3].numpy())
    back = maps.png_arrays_to_planes(arrays, th.device("cpu"))
    assert th.equal(back, planes)
    import pytest
    with pytest.raises(RuntimeError, match="CUDA"):
        maps.encode_u8(th.zeros(9, 4, 4))


def test_texel_positions_are_bit_identical_to_the_reference_division():
    """The kernels form (j + 0.5)/R as a residual-corrected product with RN(1/R) (svbrdf_core.cuh texel_position_rcp);
    it must equal the reference's fp32 division (microfacet.py:16-19) bit for bit at EVERY resolution, power of two or
    not: all R <= 2048, then a spread of larger ones up to 16384."""
    import ctypes

    import numpy as np
    import torch as th

    from tests import hostemu
    L = hostemu.lib()
    sizes = list(range(1, 2049)) + [2049, 2500, 3000, 3071, 4095, 4096, 4097, 5000, 6144, 8191, 8192, 10000, 16383, 16384]
    for res in sizes:
        a = np.empty(2 * res, dtype=np.float32)
        b = np.empty(2 * res, dtype=np.float32)
        L.emu_positions(res, ctypes.c_float(6.848), a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(a, b), res
        if res in (40, 48, 1100, 4096):          # and the division form is torch's: the reference's own expression
            t = ((th.arange(res, dtype=th.float32) + 0.5) / res - 0.5) * 6.848
            assert np.array_equal(a[0::2], t.numpy()) and np.array_equal(a[1::2], -t.numpy())


def test_fixed_divisor_division_is_ieee_division():
    """svbrdf_core.cuh div_rn — q0 = RN(x r), rem = fma(-q0, b, x), q = fma(rem, r, q0) with r = RN(1/b) — is what the consumer
    kernels use for Normalize's division by std.  The host build of that code must agree with the IEEE division bit for bit:
    the torchvision constants, awkward divisors (all-ones significand, powers of two, just above/below them), random ones;
    numerators over the range a normalised image or its gradient can take."""
    import ctypes

    from tests import hostemu
    L = hostemu.lib()
    rng = np.random.default_rng(11)
    ones = np.float32(2.0) - np.float32(2.0 ** -23)
    divisors = [0.229, 0.224, 0.225, 0.255, 0.5, 1.0, 2.0, float(ones), float(ones) / 4, float(np.nextafter(np.float32(1), np.float32(2))),
                float(np.nextafter(np.float32(1), np.float32(0))), 3.0, 1e-3, 37.25, 1e4] + list(rng.uniform(0.01, 10.0, 25))
    xs = np.concatenate([
        rng.uniform(-1.0, 1.0, 400_000),
        rng.uniform(0.0, 1.0, 200_000) - 0.485,                       # a rendered value minus a mean
        rng.standard_normal(200_000) * 10.0 ** rng.uniform(-12, 2, 200_000),   # gradients over 14 decades
        [0.0, -0.0, 1.0, -1.0, 1e-20, -1e-20, 1e20],
    ]).astype(np.float32)
    a, b = np.empty_like(xs), np.empty_like(xs)
    for d in divisors:
        L.emu_fixed_div(xs.size, xs.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(d), a.ctypes.data_as(ctypes.c_void_p),
                        b.ctypes.data_as(ctypes.c_void_p))
        nz = xs != 0                                                   # a -0.0 numerator comes out as +0.0 (equal as a number)
        assert np.array_equal(a.view(np.uint32)[nz], b.view(np.uint32)[nz]), d
        assert np.array_equal(a, b) and np.array_equal(b, xs / np.float32(d))
