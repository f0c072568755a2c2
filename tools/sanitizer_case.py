#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitizer_case.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th  # noqa: E402

import svbrdf_diff_renderer_b200 as pkg  # noqa: E402
from svbrdf_diff_renderer_b200 import maps, synth  # noqa: E402
from svbrdf_diff_renderer_b200.descriptor import MEAN, STD  # noqa: E402

dev = th.device("cuda:0")
res, n = int(os.environ.get("RES", 300)), 9           # 90000 texels: 188 tiles (partial last tile), > 148 CTAs
cl = [c.to(dev) for c in synth.calibration(n)]
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
gt, t0 = synth.random_textures(res, 1).to(dev), synth.random_textures(res, 2).to(dev)
with th.no_grad():
    target = r.eval(gt)
for env in ({}, {"SVBRDF_B200_TSTORE": "1"}, {"SVBRDF_B200_PACKED": "1"}, {"SVBRDF_B200_FORCE_LDG": "1"}):
    os.environ.update(env)
    for tgt in (target, (target * 255).to(th.uint8)):
        o = pkg.SvbrdfOptim(dev, r)
        o.load_targets(tgt)
        o.init_from_tex(t0.clone())
        losses = o.optim(3, 0.01, None, False, progress=False)
    t = t0.clone().requires_grad_(True)
    th.nn.functional.mse_loss(r.eval(t), target).backward()
    for k in env:
        os.environ.pop(k)
    print("fused/vjp ok", env, losses[-1], float(t.grad.abs().max()))
t = t0.clone().requires_grad_(True)
norm, l2 = r.eval_normalized(t, MEAN, STD, target)
(l2 + norm.square().mean()).backward()
print("norm_l2 ok", float(l2))
up = maps.handoff(t0, 2 * res)
dn = maps.resize_lanczos4_u8(maps.encode_u8(up), 77, 131)
print("maps ok", tuple(up.shape), tuple(dn.shape))
th.cuda.synchronize()
