"""ORACLE tooling: ``tests/golden/features_*.npz`` from the UNMODIFIED reference classes ``Microfacet``
(/root/reference/src/microfacet.py) and ``VGGLoss`` (/root/reference/src/descriptor.py) driven as
materialgan.py:136-147 drives them (image loss + 0.1 * feature loss, backward).  Only the constructor call
``vgg19(weights='DEFAULT')`` is patched to ``vgg19(weights=None)`` after ``torch.manual_seed`` (the pretrained weights
are a download).  Run in the build container:  python -m oracle.make_golden_features
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch as th

from oracle import ref_loader
from svbrdf_diff_renderer_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"features_32x4_coloc": (32, 4, True, 5), "features_32x4_offaxis": (32, 4, False, 6)}


def reference_vggloss(seed):
    ref_loader.load()
    import torchvision.models
    import src.descriptor as rd
    orig = rd.vgg19
    th.manual_seed(seed)
    rd.vgg19 = lambda weights=None: torchvision.models.vgg19(weights=None)
    try:
        with ref_loader.quiet():
            return rd.VGGLoss(th.device("cpu"))
    finally:
        rd.vgg19 = orig


def main():
    Microfacet, _, _ = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for name, (res, n, coloc, seed) in CASES.items():
        cl = synth.calibration(n, coloc)
        with ref_loader.quiet():
            r = Microfacet(res, n, synth.IM_SIZE_CM, [c.clone() for c in cl], th.device("cpu"))
        gt, t0 = synth.random_textures(res, seed), synth.random_textures(res, seed + 100)
        with th.no_grad():
            targets = r.eval(gt)
        vgg = reference_vggloss(seed)
        vgg.load(targets)
        tex = t0.clone().requires_grad_(True)
        img = r.eval(tex)
        l2 = th.nn.functional.mse_loss(img, targets)             # optimization.py:28-29
        lf = vgg(img) * 0.1                                      # materialgan.py:144
        g_feat, = th.autograd.grad(lf, tex, retain_graph=True)  # the feature path alone (small next to the L2 term)
        (l2 + lf).backward()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), tex=t0.numpy(), targets=targets.numpy(), cam=cl[0].numpy(), light=cl[1].numpy(),
                            power=cl[2].numpy(), vgg_seed=seed, loss_image=float(l2.detach()), loss_feature=float(lf.detach()), grad=tex.grad.numpy(),
                            grad_feature=g_feat.numpy(), normalized=vgg.normalize(img.detach()).numpy())
        print(name, float(l2), float(lf), float(tex.grad.abs().max()))


if __name__ == "__main__":
    main()
