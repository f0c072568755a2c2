#!/usr/bin/env python
"""Times the fused consumer path (render + normalise + L2, one pass each way) against the unfused torch route on the same
GPU (native render fwd/bwd + torch normalise + torch MSE), without the feature network.  Measurement tool."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th  # noqa: E402

import svbrdf_diff_renderer_b200 as pkg  # noqa: E402
from svbrdf_diff_renderer_b200 import synth  # noqa: E402
from svbrdf_diff_renderer_b200.descriptor import MEAN, STD  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--lights", type=int, default=9)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--peak", type=float, default=6540.8)
    a = ap.parse_args()
    dev = th.device("cuda:0")
    res, n = a.res, a.lights
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    with th.no_grad():
        target = r.eval(synth.random_textures(res, 1).to(dev))
    t0 = synth.random_textures(res, 2).to(dev)
    mean = th.tensor(MEAN, device=dev).view(1, 3, 1, 1)
    std = th.tensor(STD, device=dev).view(1, 3, 1, 1)
    w = th.randn(n, 3, res, res, device=dev) * 1e-6          # stands for the feature network's gradient w.r.t. its input

    def fused():
        t = t0.clone().requires_grad_(True)
        norm, l2 = r.eval_normalized(t, MEAN, STD, target)
        (l2 + (norm * w).sum()).backward()

    def unfused():
        t = t0.clone().requires_grad_(True)
        img = r.eval(t)
        (th.nn.functional.mse_loss(img, target) + (((img - mean) / std) * w).sum()).backward()

    px = res * res
    us_f, us_u = timed(fused, a.reps), timed(unfused, a.reps)
    # algorithmic bytes of the two native passes: fwd reads tex 36 + targets 12N, writes image 12N; bwd reads tex 36 +
    # grad 12N + targets 12N, writes grad_tex 36  (per texel)
    nbytes = px * (36 + 24 * n + 36 + 24 * n + 36)
    out = {"res": res, "lights": n, "fused_us": round(us_f, 1), "unfused_us": round(us_u, 1), "speedup": round(us_u / us_f, 2),
           "native_algorithmic_MB": round(nbytes / 1e6, 1), "note": "both figures include the stand-in consumer (norm*w).sum() and its backward"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
