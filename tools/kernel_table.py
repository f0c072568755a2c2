#!/usr/bin/env python
"""Time every kernel of the path at the BASELINE.json configurations that fit one GPU; writes a markdown table.

    python tools/kernel_table.py > profiles/r01_kernel_table.md
"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th
import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import _native as nv, synth

dev = th.device("cuda:0")
L = nv.lib()
PEAK = 6540.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


rows = []
for res, n, reps in ((256, 9, 50), (1024, 9, 30), (2048, 64, 8), (4096, 64, 4)):
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    P = res * res
    gt = synth.random_textures(res, 1).to(dev)
    tex = synth.random_textures(res, 2)[0].to(dev).contiguous()
    out = th.empty(n, 3, res, res, device=dev)
    geom = r._geom(r._pow)
    ws = r._workspace()
    st = nv.stream_ptr(dev)
    nv.check(L.svbrdf_render_fwd(ctypes.byref(geom), nv.ptr(gt[0].contiguous()), nv.ptr(out), st), "fwd")
    target = out.clone()
    grad = th.empty(9, res, res, device=dev)
    loss = th.zeros(1, device=dev)
    m, v = th.zeros_like(tex), th.zeros_like(tex)
    curve = th.zeros(64, device=dev)
    gout = th.randn(n, 3, res, res, device=dev)

    def fwd():
        nv.check(L.svbrdf_render_fwd(ctypes.byref(geom), nv.ptr(tex), nv.ptr(out), st), "fwd")

    def bwd():
        nv.check(L.svbrdf_render_bwd(ctypes.byref(geom), nv.ptr(tex), nv.ptr(gout), nv.ptr(grad), None, nv.ptr(ws), st), "bwd")

    def l2g():
        nv.check(L.svbrdf_l2_grad(ctypes.byref(geom), nv.ptr(tex), nv.ptr(target), 0, n, nv.ptr(grad), nv.ptr(loss), None, nv.ptr(ws), st), "l2g")

    def adam():
        a = nv.Adam(0.01, 0.9, 0.999, 1e-8, 1)
        nv.check(L.svbrdf_l2_adam_run(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(target), 0, ctypes.byref(a), 8, nv.ptr(curve), None,
                                      nv.ptr(ws), st), "adam")

    for name, fn, bpt, per_call in (("render_fwd (texel_kernel)", fwd, 36 + 12 * n, 1), ("render_bwd / VJP (tile_kernel)", bwd, 72 + 12 * n, 1),
                                    ("l2_grad (tile_kernel)", l2g, 72 + 12 * n, 1), ("l2_adam, 8 epochs per launch (tile_kernel)", adam, 216 + 12 * n, 8)):
        us = timeit(fn, reps) / per_call
        gbs = bpt * P / us * 1e-3
        rows.append((f"{res}^2 x {n}", name, us, P * n / us * 1e-3, bpt, gbs, gbs / PEAK))
    del out, target, gout, grad
    th.cuda.empty_cache()

print(f"| config | kernel | us per pass | G samples/s | algorithmic B/texel | GB/s | of measured HBM peak ({PEAK:.0f} GB/s) |")
print("|---|---|---|---|---|---|---|")
for c, k, us, gs, bpt, gbs, fr in rows:
    print(f"| {c} | {k} | {us:.1f} | {gs:.1f} | {bpt} | {gbs:.0f} | {fr * 100:.1f} % |")
