"""CPU: the combined-loss oracle (oracle/descriptor_port.py: L2 + 0.1 * VGG descriptor loss, materialgan.py:141-147)
against the reference-generated golden vectors, and against the live reference where /root/reference exists."""
import glob
import os

import numpy as np
import pytest
import torch as th

from oracle import descriptor_port as dp
from oracle import torch_port as tp
from svbrdf_diff_renderer_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "features_*.npz")))


def _scene(g):
    res = g["tex"].shape[-1]
    return tp.Scene(res, th.from_numpy(g["cam"]), th.from_numpy(g["light"]), th.from_numpy(g["power"]), synth.IM_SIZE_CM, th.float32)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_reference_combined_loss(path):
    pytest.importorskip("torchvision")
    g = np.load(path)
    net = dp.seeded_vgg_features(int(g["vgg_seed"]))
    sc = _scene(g)
    tex, targets = th.from_numpy(g["tex"]), th.from_numpy(g["targets"])
    assert np.array_equal(dp.normalize(tp.shade(sc, tex)).numpy(), g["normalized"])          # torch port is bit-identical to the reference
    l2, lf, grad = dp.combined_loss_and_grad(sc, tex, targets, net)
    assert l2 == pytest.approx(float(g["loss_image"]), rel=1e-6)
    assert lf == pytest.approx(float(g["loss_feature"]), rel=1e-5)
    np.testing.assert_allclose(grad.numpy(), g["grad"], rtol=0, atol=1e-6 * float(np.abs(g["grad"]).max()))


def test_two_feature_goldens_present():
    assert len(GOLD) == 2


def test_pinned_against_live_reference_vggloss():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("needs /root/reference (build container only)")
    pytest.importorskip("torchvision")
    from oracle.make_golden_features import reference_vggloss
    vgg = reference_vggloss(3)
    net = dp.seeded_vgg_features(3)
    x = th.rand(3, 3, 32, 32, generator=th.Generator().manual_seed(1))
    y = th.rand(3, 3, 32, 32, generator=th.Generator().manual_seed(2))
    assert th.equal(vgg.normalize(x), dp.normalize(x))
    vgg.load(y)
    ref = float(vgg(x))
    mine = float(dp.vgg_loss(net, x, dp.feature_vector(net, dp.normalize(y))))
    assert mine == pytest.approx(ref, rel=1e-6)
