"""CPU: pin the oracle port against the UNMODIFIED reference imported from /root/reference.

Only runs where that tree exists (the build container).  On the GPU box the same pin is carried
by tests/golden/*.npz, which oracle/make_golden.py generated from the reference.
"""
import numpy as np
import pytest
import torch as th

from oracle import ref_loader, torch_port as tp
from svbrdf_diff_renderer_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")

CPU = th.device("cpu")


@pytest.mark.parametrize("colocated", [True, False])
@pytest.mark.parametrize("res,n", [(40, 9), (16, 16)])
def test_forward_backward_adam_bit_exact(colocated, res, n):
    Microfacet, SvbrdfOptim, _ = ref_loader.load()
    cl = synth.calibration(n, colocated)
    gt, t0 = synth.random_textures(res, 31), synth.edge_case_textures(res, 32)
    with ref_loader.quiet():
        ref = Microfacet(res, n, synth.IM_SIZE_CM, [c.clone() for c in cl], CPU)
    sc = tp.Scene(res, cl[0], cl[1], cl[2], synth.IM_SIZE_CM)
    with th.no_grad():
        target = ref.eval(gt)
        assert th.equal(target, tp.shade(sc, gt))

    opt_ref = SvbrdfOptim(CPU, ref)
    opt_ref.load_targets(target)
    opt_ref.init_from_tex(t0.clone())
    adam = th.optim.Adam([opt_ref.textures], lr=0.01, betas=(0.9, 0.999))   # svbrdf.py:52
    ref_losses = []
    for _ in range(4):                                                      # svbrdf.py:60-71
        loss = opt_ref.compute_image_loss(ref.eval(opt_ref.textures.clamp(-1, 1)))
        ref_losses.append(loss.item())
        adam.zero_grad()
        loss.backward()
        adam.step()
    maps, losses, _ = tp.optimise(sc, t0, target, 4, 0.01)
    assert losses == ref_losses
    assert th.equal(maps, opt_ref.textures.detach())


def test_fp64_arbiter_matches_reference_in_double():
    Microfacet, _, _ = ref_loader.load()
    cl = synth.calibration(9, False)
    tex = synth.random_textures(24, 5)
    with ref_loader.quiet():
        ref = ref_loader.to_double(Microfacet(24, 9, synth.IM_SIZE_CM, [c.double() for c in cl], CPU))
    sc = tp.Scene(24, cl[0], cl[1], cl[2], synth.IM_SIZE_CM, th.float64)
    a, b = ref.eval(tex.double()), tp.shade(sc, tex.double())
    assert th.equal(a, b)


def test_synthetic_textures_follow_reference_initialiser():
    """synth.random_textures(seed) == torch.manual_seed(seed); SvbrdfOptim.init_from_randn() (svbrdf.py:32-39)."""
    Microfacet, SvbrdfOptim, _ = ref_loader.load()
    with ref_loader.quiet():
        ref = Microfacet(32, 9, synth.IM_SIZE_CM, synth.calibration(9), CPU)
    th.manual_seed(77)
    o = SvbrdfOptim(CPU, ref)
    o.init_from_randn()
    assert th.equal(o.textures.detach(), synth.random_textures(32, 77))


def test_render_json_geometry():
    import json
    with open(ref_loader.REFERENCE_ROOT + "/data/random/render.json") as f:
        cfg = json.load(f)
    cl = synth.calibration(9)
    assert np.array_equal(np.array(cfg["camera_pos"], "float32"), cl[0].numpy())
    assert np.array_equal(np.array(cfg["light_pos"], "float32"), cl[1].numpy())
    assert cfg["light_pow"] == list(synth.LIGHT_POW) and cfg["im_size"] == synth.IM_SIZE_CM
