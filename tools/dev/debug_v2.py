#!/usr/bin/env python
"""Development: gradient error of the library named by SVBRDF_B200_LIB against the fp64 oracle on the features golden
(random upstream gradient through the normalised render + the L2 term), next to the fp32 oracle's own error."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th
import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import synth
from oracle import torch_port as tp

DEV = th.device("cuda:0")
for name in ("features_32x4_coloc", "features_32x4_offaxis"):
    g = np.load(f"tests/golden/{name}.npz")
    cl_cpu = [th.from_numpy(g[k]) for k in ("cam", "light", "power")]
    r = pkg.Microfacet(32, 4, synth.IM_SIZE_CM, [c.to(DEV) for c in cl_cpu], DEV)
    targets = th.from_numpy(g["targets"])
    mean = th.tensor([0.485, 0.456, 0.406]); std = th.tensor([0.229, 0.224, 0.225])
    w = th.from_numpy(np.random.default_rng(0).standard_normal((4, 3, 32, 32)).astype(np.float32))
    tex = th.from_numpy(g["tex"]).to(DEV).requires_grad_(True)
    norm, l2 = r.eval_normalized(tex, mean.to(DEV), std.to(DEV), targets.to(DEV))
    gf, = th.autograd.grad((norm * w.to(DEV)).sum() + 1000 * l2, tex)
    ref = {}
    for dt in (th.float32, th.float64):
        sc = tp.Scene(32, cl_cpu[0], cl_cpu[1], cl_cpu[2], synth.IM_SIZE_CM, dt)
        t = th.from_numpy(g["tex"]).to(dt).requires_grad_(True)
        img = tp.shade(sc, t)
        loss = (((img - mean.to(dt)[None, :, None, None]) / std.to(dt)[None, :, None, None]) * w.to(dt)).sum() + 1000 * ((img - targets.to(dt)) ** 2).mean()
        ref[dt], = th.autograd.grad(loss, t)
    x, r32, r64 = gf.cpu().double().numpy().reshape(9, -1), ref[th.float32].double().numpy().reshape(9, -1), ref[th.float64].numpy().reshape(9, -1)
    m = np.abs(r64).max()
    print(f"== {name} lib={os.environ.get('SVBRDF_B200_LIB', 'default')}  max|ref| {m:.3e}")
    for c in range(9):
        ex, er = np.abs(x[c] - r64[c]), np.abs(r32[c] - r64[c])
        tol = 1e-4 * np.abs(r64[c]) + 2e-4 * m
        tol2 = 1e-4 * np.abs(r64[c]) + 1e-4 * m
        print(f"  ch{c}: kernel max {ex.max():.3e} mean {ex.mean():.3e} bad {int((ex > tol).sum())}/{int((ex > tol2).sum())} | ref32 max {er.max():.3e} mean {er.mean():.3e} bad {int((er > tol).sum())}/{int((er > tol2).sum())}"
              f" | worst texel {int(ex.argmax())} t5={g['tex'][0, 5].ravel()[ex.argmax()]:.4f}")
