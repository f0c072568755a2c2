#!/usr/bin/env python
"""Launches every mode-B / sharded-mode kernel twice at 1024^2 x 9 (for `ncu --set full -k regex:...`): render_fwd
(texel_kernel<Render>), render_bwd (tile_kernel<Vjp>), l2_grad (tile_kernel<L2Grad>), norm_l2 forward and backward."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th
import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import _native as nv, synth
from svbrdf_diff_renderer_b200.descriptor import MEAN, STD

res = int(os.environ.get("RES", "1024")); n = int(os.environ.get("LIGHTS", "9"))
dev = th.device("cuda:0")
L = nv.lib()
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in synth.calibration(n)], dev)
tex = synth.random_textures(res, 2)[0].to(dev).contiguous()
out = th.empty(n, 3, res, res, device=dev)
geom, ws, st = r._geom(r._pow), r._workspace(), nv.stream_ptr(dev)
grad = th.empty(9, res, res, device=dev); loss = th.zeros(1, device=dev)
gout = th.randn(n, 3, res, res, device=dev)
with th.no_grad():
    target = r.eval(synth.random_textures(res, 1).to(dev)).contiguous()
for _ in range(2):
    nv.check(L.svbrdf_render_fwd(ctypes.byref(geom), nv.ptr(tex), nv.ptr(out), st), "fwd")
    nv.check(L.svbrdf_render_bwd(ctypes.byref(geom), nv.ptr(tex), nv.ptr(gout), nv.ptr(grad), None, nv.ptr(ws), st), "bwd")
    nv.check(L.svbrdf_l2_grad(ctypes.byref(geom), nv.ptr(tex), nv.ptr(target), 0, n, nv.ptr(grad), nv.ptr(loss), None, nv.ptr(ws), st), "l2g")
    t = tex[None].clone().requires_grad_(True)
    norm, l2 = r.eval_normalized(t, MEAN, STD, target)
    th.autograd.grad([norm, l2], [t], [gout, th.ones_like(l2)])
th.cuda.synchronize()
print("done")
