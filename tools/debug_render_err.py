import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import numpy as np, torch as th
import svbrdf_diff_renderer_b200 as pkg
from tests import parity
dev = th.device("cuda:0")
for name in ("coloc_32x9", "offaxis_32x9", "wellcond_32x9"):
    g = parity.golden(name)
    T = lambda a: th.from_numpy(np.ascontiguousarray(a)).to(dev)
    r = pkg.Microfacet(int(g["res"]), int(g["n"]), float(g["size"]), [T(g["cam"]), T(g["light"]), T(g["power_render"])], dev)
    with th.no_grad():
        img = r.eval(T(g["tex_gt"])).cpu().numpy().astype(np.float64)
    ref = g["render_gt_f64"]
    rel = (img - ref) / ref
    print(name, "signed rel err: mean %.3e median %.3e std %.3e | abs mean %.3e p99.9 %.3e" % (rel.mean(), np.median(rel), rel.std(), np.abs(rel).mean(), np.quantile(np.abs(rel), 0.999)))
    lg = np.log2(ref)
    for lo, hi in ((-10, -4), (-4, -2), (-2, -1), (-1, -0.3), (-0.3, 0.01)):
        msk = (lg >= lo) & (lg < hi)
        if msk.sum():
            print("   lg2(out) in [%5.1f,%5.1f): n=%6d mean signed rel %.3e" % (lo, hi, msk.sum(), rel[msk].mean()))
    ref32 = g["target"].astype(np.float64)
    rr = (ref32 - ref) / ref
    print("   reference fp32: signed mean %.3e abs mean %.3e" % (rr.mean(), np.abs(rr).mean()))
