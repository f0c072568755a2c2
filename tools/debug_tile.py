import ctypes, faulthandler, os, sys
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th
import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import _native as nv, synth
dev = th.device("cuda:0")
res, n = int(os.environ.get("RES", "64")), 9
cl = [c.to(dev) for c in synth.calibration(n)]
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
print("renderer ok", flush=True)
with th.no_grad():
    tgt = r.eval(synth.random_textures(res, 1).to(dev)).contiguous()
th.cuda.synchronize(); print("render ok", float(tgt.mean()), flush=True)
tex = synth.random_textures(res, 2)[0].to(dev)
m, v = th.zeros_like(tex), th.zeros_like(tex)
ws = r._workspace(); loss = th.zeros(1, device=dev)
geom = r._geom(r._pow)
L = nv.lib()
for it in range(3):
    a = nv.Adam(0.01, 0.9, 0.999, 1e-8, it + 1)
    print("launching step", it, flush=True)
    code = L.svbrdf_l2_adam_step(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(tgt), 0, ctypes.byref(a), nv.ptr(loss), None, nv.ptr(ws), nv.stream_ptr(dev))
    print("returned", code, L.svbrdf_error_string(code), flush=True)
    th.cuda.synchronize()
    print("synced, loss", float(loss.item()), flush=True)
print("done", flush=True)
