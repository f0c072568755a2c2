#!/usr/bin/env python
"""Times SvbrdfOptim.optim(E epochs) on device-resident inputs (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th
import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import synth
dev = th.device("cuda:0")
res, n = 1024, 9
cl = [c.to(dev) for c in synth.calibration(n)]
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
with th.no_grad():
    tgt = r.eval(synth.random_textures(res, 1).to(dev)).contiguous()
t0 = synth.random_textures(res, 2).to(dev)
o = pkg.SvbrdfOptim(dev, r)
o.load_targets(tgt)
for E in (1, 2, 5, 19, 20, 40):
    for rep in range(3):
        o.init_from_tex(t0.clone())
        th.cuda.synchronize()
        t = time.perf_counter()
        losses = o.optim(E, 0.01, None, False, progress=False)
        th.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e3
    print(f"optim({E}): {dt:.3f} ms = {dt / E * 1e3:.1f} us/epoch, last loss {losses[-1]:.6f}", flush=True)
