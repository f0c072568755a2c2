#!/usr/bin/env python
"""Is the rendered ground truth an exact fixed point of the fused step?  Prints the epoch-0 loss at the ground truth and
the number of texels with a non-zero gradient, for the library selected by SVBRDF_B200_LIB / SVBRDF_B200_* switches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th  # noqa: E402

import svbrdf_diff_renderer_b200 as pkg  # noqa: E402
from svbrdf_diff_renderer_b200 import synth  # noqa: E402

res, n = int(os.environ.get("RES", 1024)), 9
dev = th.device("cuda:0")
cl = [c.to(dev) for c in synth.calibration(n)]
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
gt = synth.random_textures(res, 1).to(dev)
with th.no_grad():
    target = r.eval(gt)
o = pkg.SvbrdfOptim(dev, r)
o.load_targets(target)
o.init_from_tex(gt.clone())
losses = o.optim(3, 0.01, None, False, progress=False)
moved = (o.textures.detach() - gt).abs()
print("losses", losses, "max move", float(moved.max()), "texels moved > 1e-6:", int((moved.amax(1) > 1e-6).sum()))
