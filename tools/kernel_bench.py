#!/usr/bin/env python
"""Times kernel variants of the fused step (env-selected) on one GPU: a development tool, not the bench.

    python tools/kernel_bench.py [--res 1024 --lights 9 --steps 30]
"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th  # noqa: E402

import svbrdf_diff_renderer_b200 as pkg  # noqa: E402
from svbrdf_diff_renderer_b200 import _native as nv, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--lights", type=int, default=9)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--mats", type=int, default=4)
    ap.add_argument("--u8", action="store_true")
    ap.add_argument("--fused-epochs", action="store_true", help="time ONE svbrdf_l2_adam_run call of `steps` epochs per material")
    ap.add_argument("--variants", default="ldg;tma2;tma2s6;tma2s8;tma1;tma3")
    a = ap.parse_args()
    dev = th.device("cuda:0")
    res, n = a.res, a.lights
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    mats = []
    for i in range(a.mats):
        with th.no_grad():
            tgt = r.eval(synth.random_textures(res, 100 + 2 * i).to(dev)).contiguous()
        if a.u8:
            tgt = (tgt * 255).to(th.uint8)
        mats.append((tgt, synth.random_textures(res, 101 + 2 * i)[0].to(dev)))
    L = nv.lib()
    ws = r._workspace()
    geom = r._geom(r._pow)
    loss = th.zeros(1, device=dev)
    presets = {
        "ldg": {"SVBRDF_B200_FORCE_LDG": "1"},
        "tma1": {"SVBRDF_B200_CTAS_PER_SM": "1"},
        "tma2": {"SVBRDF_B200_CTAS_PER_SM": "2"},
        "tma3": {"SVBRDF_B200_CTAS_PER_SM": "3"},
        "tma2s6": {"SVBRDF_B200_CTAS_PER_SM": "2", "SVBRDF_B200_SLOTS": "6"},
        "tma2s8": {"SVBRDF_B200_CTAS_PER_SM": "2", "SVBRDF_B200_SLOTS": "8"},
        "tma2s4": {"SVBRDF_B200_CTAS_PER_SM": "2", "SVBRDF_B200_SLOTS": "4"},
    }
    keys = ("SVBRDF_B200_FORCE_LDG", "SVBRDF_B200_CTAS_PER_SM", "SVBRDF_B200_SLOTS", "SVBRDF_B200_PACKED", "SVBRDF_B200_TSTORE")
    out = {}

    def preset(name):
        """`ldg`, or tma<ctas>[s<slots>][p][n] (p = packed FP32x2 shape, n = per-thread STG instead of TMA store-back)."""
        if name in presets:
            return presets[name]
        import re
        mo = re.fullmatch(r"tma(\d)(?:s(\d+))?(p)?(n)?", name)
        env = {"SVBRDF_B200_CTAS_PER_SM": mo.group(1)}
        if mo.group(2):
            env["SVBRDF_B200_SLOTS"] = mo.group(2)
        if mo.group(3):
            env["SVBRDF_B200_PACKED"] = "1"
        if mo.group(4):
            env["SVBRDF_B200_TSTORE"] = "0"
        return env

    for name in a.variants.split(";"):
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(preset(name))
        state = [(t, x.clone(), th.zeros_like(x), th.zeros_like(x)) for t, x in mats]

        def step(i, k=[0]):
            k[0] += 1
            tgt, tex, m, v = state[i % len(state)]
            ad = nv.Adam(0.01, 0.9, 0.999, 1e-8, k[0])
            nv.check(L.svbrdf_l2_adam_step(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(tgt), nv.target_dtype_code(tgt),
                                           ctypes.byref(ad), nv.ptr(loss), None, nv.ptr(ws), nv.stream_ptr(dev)), "step")
        curve = th.zeros(max(a.steps, 8), device=dev)

        def run(i, epochs, first):
            tgt, tex, m, v = state[i % len(state)]
            ad = nv.Adam(0.01, 0.9, 0.999, 1e-8, first)
            nv.check(L.svbrdf_l2_adam_run(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(tgt), nv.target_dtype_code(tgt),
                                          ctypes.byref(ad), epochs, nv.ptr(curve), None, nv.ptr(ws), nv.stream_ptr(dev)), "run")
        for i in range(5):
            step(i)
        if a.fused_epochs:
            run(0, 3, 6)
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        if a.fused_epochs:
            for i in range(len(state)):
                run(i, a.steps, 10)
            total = a.steps * len(state)
        else:
            for i in range(a.steps):
                step(i)
            total = a.steps
        e1.record()
        th.cuda.synchronize()
        us = e0.elapsed_time(e1) / total * 1e3
        if a.fused_epochs:
            loss.copy_(curve[a.steps - 1:a.steps])
        tb = 3 if a.u8 else 12
        gbs = (216 + tb * n) * res * res / (us * 1e-6) / 1e9
        out[name] = {"us_per_step": round(us, 2), "Gsamples_s": round(res * res * n / us * 1e-3, 2), "algo_GBs": round(gbs, 1),
                     "loss": float(loss.item())}
        print(name, out[name], flush=True)
    print(json.dumps({"res": res, "lights": n, "u8": a.u8, "results": out}))


if __name__ == "__main__":
    main()
