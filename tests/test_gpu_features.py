"""GPU: the fused consumer path (svbrdf_render_norm_l2_fwd / _bwd, Microfacet.eval_normalized, descriptor.VGGLoss,
SvbrdfOptim.optim_with_features) against torch on the same device, the oracle and the reference-generated goldens."""
import glob
import os

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu

import svbrdf_diff_renderer_b200 as pkg  # noqa: E402
from oracle import descriptor_port as dp  # noqa: E402
from svbrdf_diff_renderer_b200 import synth  # noqa: E402
from svbrdf_diff_renderer_b200.descriptor import MEAN, STD  # noqa: E402

DEV = th.device("cuda:0")
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "features_*.npz")))


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32)
    th.backends.cudnn.allow_tf32 = False
    th.backends.cuda.matmul.allow_tf32 = False
    yield
    th.backends.cudnn.allow_tf32, th.backends.cuda.matmul.allow_tf32 = old


def _setup(res, n, coloc=True, seed=1):
    cl = [c.to(DEV) for c in synth.calibration(n, coloc)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, DEV)
    gt, t0 = synth.random_textures(res, seed).to(DEV), synth.random_textures(res, seed + 1).to(DEV)
    with th.no_grad():
        target = r.eval(gt)
    return r, t0, target


@pytest.mark.parametrize("res,n,coloc", [(64, 9, True), (50, 4, False), (25, 9, True)])
def test_forward_is_eval_then_normalize_then_mse(res, n, coloc):
    r, t0, target = _setup(res, n, coloc)
    with th.no_grad():
        img = r.eval(t0)
        norm, l2 = r.eval_normalized(t0, MEAN, STD, target)
        norm_only, zero = r.eval_normalized(t0, MEAN, STD)
    mean = th.tensor(MEAN, device=DEV).view(1, 3, 1, 1)
    std = th.tensor(STD, device=DEV).view(1, 3, 1, 1)
    assert th.equal(norm, (img - mean) / std)                      # sub then IEEE div, as torchvision Normalize does
    assert th.equal(norm_only, norm) and float(zero) == 0.0
    ref = th.nn.functional.mse_loss(img.double(), target.double())
    assert float(l2) == pytest.approx(float(ref), rel=2e-6)
    # uint8 targets are decoded in-kernel
    t8 = (target * 255).round().to(th.uint8)
    with th.no_grad():
        _, l8 = r.eval_normalized(t0, MEAN, STD, t8)
    assert float(l8) == pytest.approx(float(th.nn.functional.mse_loss(img.double(), t8.double() / 255)), rel=2e-6)


@pytest.mark.parametrize("res,n,coloc,u8,chunk", [(256, 9, True, False, 0), (128, 16, True, True, 0), (96, 9, False, False, 4),
                                                  (500, 4, True, False, 0)])
def test_forward_on_the_ring_matches_the_ldg_kernel(res, n, coloc, u8, chunk, monkeypatch):
    """svbrdf_render_norm_l2_fwd streams its targets through the TMA ring (tile_kernel<NormFwd>) and divides by std with the
    fixed-divisor FMA sequence; the one-thread-per-texel kernel (general IEEE division) must give the same image bit for
    bit — and both the image torch's own sub/div gives.  Several tiles per CTA, full and partial light chunks (9 lights in
    4-light slots: 4 + 4 + 1), 3- and 4-light slots, a non-power-of-two resolution."""
    if chunk:
        monkeypatch.setenv("SVBRDF_B200_CHUNK", str(chunk))
    r, t0, target = _setup(res, n, coloc, seed=5)
    tgt = (target * 255).round().to(th.uint8) if u8 else target
    with th.no_grad():
        img = r.eval(t0)
        norm, l2 = r.eval_normalized(t0, MEAN, STD, tgt)
        monkeypatch.setenv("SVBRDF_B200_FORCE_LDG", "1")
        norm_ldg, l2_ldg = r.eval_normalized(t0, MEAN, STD, tgt)
        monkeypatch.delenv("SVBRDF_B200_FORCE_LDG")
    mean = th.tensor(MEAN, device=DEV).view(1, 3, 1, 1)
    std = th.tensor(STD, device=DEV).view(1, 3, 1, 1)
    assert th.equal(norm, norm_ldg)
    assert th.equal(norm, (img - mean) / std)
    assert float(l2) == pytest.approx(float(l2_ldg), rel=2e-6)
    # awkward divisors: all-ones significand, a power of two, a large and a small one
    odd_std = [float(np.float32(2.0) - np.float32(2.0 ** -23)), 0.5, 37.25]
    with th.no_grad():
        norm2, _ = r.eval_normalized(t0, MEAN, odd_std, tgt)
    assert th.equal(norm2, (img - mean) / th.tensor(odd_std, device=DEV).view(1, 3, 1, 1))


@pytest.mark.parametrize("coloc", [True, False])
def test_backward_matches_unfused_autograd(coloc):
    """Fused backward vs eval -> normalize -> conv stand-in + mse through torch autograd (the unfused route ends in the
    same native VJP, so this checks the fusion: upstream assembly from both gradients, std division, L2 weight)."""
    res, n = 48, 4
    r, t0, target = _setup(res, n, coloc, seed=3)
    th.manual_seed(0)
    net = th.nn.Sequential(th.nn.Conv2d(3, 8, 3, padding=1), th.nn.ReLU(), th.nn.Conv2d(8, 4, 3, padding=1)).to(DEV)
    mean = th.tensor(MEAN, device=DEV).view(1, 3, 1, 1)
    std = th.tensor(STD, device=DEV).view(1, 3, 1, 1)
    pw = r._pow.clone().requires_grad_(True)
    r.update_light(pw)
    ta = t0.clone().requires_grad_(True)
    img = r.eval(ta)
    la = th.nn.functional.mse_loss(img, target) * 0.7 + net((img - mean) / std).square().mean() * 3.0
    la.backward()
    ga, gpa = ta.grad.clone(), pw.grad.clone()
    pw.grad = None
    tb = t0.clone().requires_grad_(True)
    norm, l2 = r.eval_normalized(tb, MEAN, STD, target)
    lb = l2 * 0.7 + net(norm).square().mean() * 3.0
    lb.backward()
    assert float(lb.detach()) == pytest.approx(float(la.detach()), rel=1e-5)
    scale = float(ga.abs().max())
    assert float((tb.grad - ga).abs().max()) < 2e-5 * scale
    np.testing.assert_allclose(pw.grad.cpu().numpy(), gpa.cpu().numpy(), rtol=2e-4)
    # feature term only (no targets) and L2 term only (feature output unused)
    tc = t0.clone().requires_grad_(True)
    nrm, _ = r.eval_normalized(tc, MEAN, STD)
    net(nrm).square().mean().backward()
    td = t0.clone().requires_grad_(True)
    net((r.eval(td) - mean) / std).square().mean().backward()
    assert float((tc.grad - td.grad).abs().max()) < 2e-5 * float(td.grad.abs().max())
    te = t0.clone().requires_grad_(True)
    _, l2e = r.eval_normalized(te, MEAN, STD, target)
    l2e.backward()
    tf = t0.clone().requires_grad_(True)
    th.nn.functional.mse_loss(r.eval(tf), target).backward()
    assert float((te.grad - tf.grad).abs().max()) < 2e-5 * float(tf.grad.abs().max())
    r.update_light(pw.detach())


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_combined_loss_against_reference_golden(path):
    """materialgan.py:141-147 on the reference's own classes (CPU, seeded VGG19) vs the fused path + cuDNN VGG19."""
    g = np.load(path)
    res, n = g["tex"].shape[-1], g["cam"].shape[0]
    cl = [th.from_numpy(g[k]).to(DEV) for k in ("cam", "light", "power")]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, DEV)
    from torchvision.models import vgg19
    th.manual_seed(int(g["vgg_seed"]))
    vgg = pkg.VGGLoss(DEV, net=vgg19(weights=None).features)
    targets = th.from_numpy(g["targets"]).to(DEV)
    vgg.load(targets)
    tex = th.from_numpy(g["tex"]).to(DEV).requires_grad_(True)
    norm, l2 = r.eval_normalized(tex, vgg.mean, vgg.std, targets)
    # single texels at a GGX peak are ill-conditioned in fp32 (the reference's own fp32 render is off by >10 % there,
    # SURVEY.md Appendix C), so: pass fraction + median, as for the render goldens; /std multiplies errors by ~4.4
    err = (norm.detach().cpu() - th.from_numpy(g["normalized"])).abs()
    assert float((err <= 3e-5).float().mean()) > 0.999 and float(err.median()) < 1e-6
    lf = vgg.forward_normalized(norm) * 0.1
    assert float(l2) == pytest.approx(float(g["loss_image"]), rel=1e-5)
    assert float(lf) == pytest.approx(float(g["loss_feature"]), rel=1e-4)
    gfeat, = th.autograd.grad(lf, tex, retain_graph=True)
    gnorm, = th.autograd.grad(lf, norm, retain_graph=True)      # what the feature network hands the render backward
    (l2 + lf).backward()
    # (1) the native backward against the oracle's autograd for the SAME upstream gradient (fp64 arbiter, fp32 oracle's own
    # noise beside it): this is the kernel check proper and uses the tolerances of tests/parity.py
    from oracle import torch_port as tp
    from tests import parity
    cl_cpu = [th.from_numpy(g[k]) for k in ("cam", "light", "power")]
    up = gnorm.detach().cpu()
    oracle = {}
    for dt in (th.float32, th.float64):
        sc = tp.Scene(res, cl_cpu[0], cl_cpu[1], cl_cpu[2], synth.IM_SIZE_CM, dt)
        t = th.from_numpy(g["tex"]).to(dt).requires_grad_(True)
        img = tp.shade(sc, t)
        nrm = (img - th.tensor(MEAN, dtype=dt)[None, :, None, None]) / th.tensor(STD, dtype=dt)[None, :, None, None]
        total = (nrm * up.to(dt)).sum() + tp.l2_loss(img, th.from_numpy(g["targets"]).to(dt))
        oracle[dt], = th.autograd.grad(total, t)
    parity.check_against_arbiter(tex.grad.cpu().numpy(), oracle[th.float32].numpy(), oracle[th.float64].numpy(), parity.RTOL_GRAD,
                                 f"{os.path.basename(path)[:-4]} consumer backward, same upstream gradient")
    # (2) end to end against the reference's own run (CPU VGG19).  The seeded random-weight VGG19 sits between the two
    # fp32 renders and the render backward and amplifies their ~1e-7 differences: two builds of this library whose
    # gradients both sit inside the fp32 oracle's own error for a FIXED upstream (gpurun visit r02 f3: no element outside
    # the mixed tolerance for either) score 0.9923 and > 0.999 here.  Hence a looser pass fraction than (1); recorded.
    for got, ref in ((tex.grad, g["grad"]), (gfeat, g["grad_feature"])):
        ref = th.from_numpy(ref).to(DEV)
        ok = (got - ref).abs() <= 1e-4 * ref.abs() + 2e-4 * ref.abs().max()
        parity.record_margin(f"{os.path.basename(path)[:-4]} end-to-end grad vs reference golden", frac=float(ok.float().mean()),
                             mean_err=float((got - ref).abs().mean() / ref.abs().max()))
        assert float(ok.float().mean()) > 0.98
        assert float((got - ref).abs().mean()) < 2e-5 * float(ref.abs().max())
    # drop-in class semantics: forward(x) == forward_normalized(normalize(x)) and equals the oracle's restatement
    with th.no_grad():
        img = r.eval(tex.detach())
        assert float(vgg(img)) * 0.1 == pytest.approx(float(lf), rel=1e-5)
    net = dp.seeded_vgg_features(int(g["vgg_seed"]))
    ref_l = float(dp.vgg_loss(net, img.cpu(), dp.feature_vector(net, dp.normalize(targets.cpu())))) * 0.1
    assert float(lf) == pytest.approx(ref_l, rel=1e-4)


def test_optim_with_features_descends():
    """BASELINE configs[1] wording: per-pixel optimisation on L2 + descriptor loss (64^2 here; seeded VGG19)."""
    from torchvision.models import vgg19
    r, t0, target = _setup(64, 9)
    th.manual_seed(7)
    vgg = pkg.VGGLoss(DEV, net=vgg19(weights=None).features)
    vgg.load(target)
    o = pkg.SvbrdfOptim(DEV, r)
    o.load_targets(target)
    o.init_from_tex(t0.clone())
    li, lf = o.optim_with_features(15, 0.01, vgg, 0.1)
    assert len(li) == 15 and li[-1] < 0.7 * li[0] and lf[-1] < lf[0] and all(np.isfinite(li)) and all(np.isfinite(lf))
    # with a zero feature weight the trajectory is the plain mode-B L2 optimisation
    o2 = pkg.SvbrdfOptim(DEV, r)
    o2.load_targets(target)
    o2.init_from_tex(t0.clone())
    l0, _ = o2.optim_with_features(5, 0.01, vgg, 0.0)
    o3 = pkg.SvbrdfOptim(DEV, r)
    o3.load_targets(target)
    o3.init_from_tex(t0.clone())
    l1 = o3.optim(5, 0.01, None, False, fused=False, progress=False)
    np.testing.assert_allclose(l0, l1, rtol=1e-5)
