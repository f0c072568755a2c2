"""ORACLE tooling: generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden

Every fixture stores its inputs next to the reference's outputs, so tests never
depend on RNG reproducibility.  Outputs come from the reference classes
``Microfacet`` (``/root/reference/src/microfacet.py``) and ``SvbrdfOptim``
(``/root/reference/src/svbrdf.py``) driven exactly as ``SvbrdfOptim.optim`` drives
them (svbrdf.py:48-71: Adam(lr, betas=(0.9,0.999)); clamp -> eval -> MSE ->
zero_grad/backward/step), minus tqdm and file dumps.  ``*_f64`` arrays come from
the same classes with their geometry attributes widened to double (the fp64
arbiter of SURVEY.md Appendix C).
"""

from __future__ import annotations

import os

import numpy as np
import torch as th

from oracle import ref_loader
from svbrdf_diff_renderer_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CPU = th.device("cpu")


def _renderer(res, n, cl, double=False):
    Microfacet, _, _ = ref_loader.load()
    cl = [c.clone() for c in cl]
    if double:
        cl = [c.double() for c in cl]
    with ref_loader.quiet():
        r = Microfacet(res, n, synth.IM_SIZE_CM, cl, CPU)
    if double:
        ref_loader.to_double(r)
    return r, cl


def _run_optim(res, n, cl, tex0, target, epochs, lr, optim_light, double):
    """The reference loop body, svbrdf.py:48-71."""
    _, SvbrdfOptim, _ = ref_loader.load()
    r, cl = _renderer(res, n, cl, double)
    o = SvbrdfOptim(CPU, r)
    o.load_targets(target.double() if double else target)
    o.init_from_tex((tex0.double() if double else tex0).clone())
    params = [o.textures]
    if optim_light:
        cl[2] = o.gradient(cl[2])
        params = params + [cl[2]]
    opt = th.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
    losses, grad0 = [], None
    for epoch in range(epochs):
        if optim_light:
            r.update_light(cl[2])
        img = r.eval(o.textures.clamp(-1, 1))
        loss = o.compute_image_loss(img)
        losses.append(loss.item())
        opt.zero_grad()
        loss.backward()
        if epoch == 0:
            grad0 = o.textures.grad.detach().clone()
            gpow0 = cl[2].grad.detach().clone() if optim_light else None
        opt.step()
    return o.textures.detach(), np.array(losses), grad0, gpow0, cl[2].detach()


def make_case(name, res, n, colocated, tex_gt, tex0, epochs=20, lr=0.01, optim_light=False, power0=None):
    cl = synth.calibration(n, colocated)
    r32, _ = _renderer(res, n, cl)
    r64, _ = _renderer(res, n, cl, double=True)
    with th.no_grad():
        target = r32.eval(tex_gt)                      # fp32 targets, as the workflow makes them (scripts.py:31-41)
        render_f64 = r64.eval(tex_gt.double())
        start_f32 = r32.eval(tex0.clamp(-1, 1))
        start_f64 = r64.eval(tex0.double().clamp(-1, 1))
    cl_run = [cl[0], cl[1], cl[2] if power0 is None else power0]
    maps32, loss32, g32, gp32, pw32 = _run_optim(res, n, cl_run, tex0, target, epochs, lr, optim_light, False)
    maps64, loss64, g64, gp64, pw64 = _run_optim(res, n, cl_run, tex0, target, epochs, lr, optim_light, True)

    # mode B: VJP of eval() alone for an arbitrary upstream gradient
    gen = th.Generator().manual_seed(11)
    gimg = th.randn(n, 3, res, res, generator=gen)
    vjp = {}
    for tag, r, dt in (("f32", r32, th.float32), ("f64", r64, th.float64)):
        t = tex0.clamp(-1, 1).to(dt).requires_grad_(True)
        p = cl[2].to(dt).clone().requires_grad_(True)
        r.update_light(p)
        r.eval(t).backward(gimg.to(dt))
        vjp[tag] = (t.grad.clone(), p.grad.clone())

    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(
        path,
        res=res, n=n, size=synth.IM_SIZE_CM, epochs=epochs, lr=lr, optim_light=optim_light,
        cam=cl[0].numpy(), light=cl[1].numpy(), power=cl_run[2].numpy(), power_render=cl[2].numpy(),
        tex_gt=tex_gt.numpy(), tex0=tex0.numpy(),
        target=target.numpy(), render_gt_f64=render_f64.numpy(),
        render_start_f32=start_f32.numpy(), render_start_f64=start_f64.numpy(),
        loss_f32=loss32, loss_f64=loss64,
        grad0_f32=g32.numpy(), grad0_f64=g64.numpy(),
        gpow0_f32=(gp32.numpy() if gp32 is not None else np.zeros(0)),
        gpow0_f64=(gp64.numpy() if gp64 is not None else np.zeros(0)),
        maps_f32=maps32.numpy(), maps_f64=maps64.numpy(),
        power_final_f32=pw32.numpy(), power_final_f64=pw64.numpy(),
        grad_img=gimg.numpy(),
        vjp_tex_f32=vjp["f32"][0].numpy(), vjp_tex_f64=vjp["f64"][0].numpy(),
        vjp_pow_f32=vjp["f32"][1].numpy(), vjp_pow_f64=vjp["f64"][1].numpy(),
    )
    print(f"{name}: loss {loss32[0]:.6g} -> {loss32[-1]:.6g}   ({os.path.getsize(path) / 1024:.0f} KiB)")


def make_stats_256():
    """Config 1 of BASELINE.json (256^2 x 9, 20 epochs): scalars and per-channel sums only."""
    res, n = 256, 9
    cl = synth.calibration(n, True)
    tex_gt, tex0 = synth.random_textures(res, 1), synth.random_textures(res, 2)
    r32, _ = _renderer(res, n, cl)
    with th.no_grad():
        target = r32.eval(tex_gt)
    out = {}
    for tag, dbl in (("f32", False), ("f64", True)):
        maps, loss, g0, _, _ = _run_optim(res, n, cl, tex0, target, 20, 0.01, False, dbl)
        out[f"loss_{tag}"] = loss
        out[f"grad0_chan_sum_{tag}"] = g0.double().sum((0, 2, 3)).numpy()
        out[f"grad0_chan_abs_{tag}"] = g0.double().abs().sum((0, 2, 3)).numpy()
        out[f"maps_chan_sum_{tag}"] = maps.double().sum((0, 2, 3)).numpy()
        out[f"maps_chan_sq_{tag}"] = (maps.double() ** 2).sum((0, 2, 3)).numpy()
        # a 16x256 strip of the final maps, enough for an element-wise spot check
        out[f"maps_strip_{tag}"] = maps[:, :, 120:136, :].numpy()
    out["target_chan_sum"] = target.double().sum((0, 2, 3)).numpy()
    out["target_sq_sum"] = float((target.double() ** 2).sum())
    path = os.path.join(OUT, "config1_256x9_stats.npz")
    np.savez_compressed(path, res=res, n=n, seed_gt=1, seed_start=2, **out)
    print(f"config1 stats: loss {out['loss_f32'][0]:.6g} -> {out['loss_f32'][-1]:.6g}   ({os.path.getsize(path) / 1024:.0f} KiB)")


def make_stats_1024():
    """Config 2 of BASELINE.json (1024^2 x 9, 20 epochs, run.py:56): loss curves, per-channel sums and three 8-row
    strips of the final maps (top, middle, bottom), fp32 and fp64, from the unmodified reference."""
    res, n = 1024, 9
    cl = synth.calibration(n, True)
    tex_gt, tex0 = synth.random_textures(res, 1), synth.random_textures(res, 2)
    r32, _ = _renderer(res, n, cl)
    with th.no_grad():
        target = r32.eval(tex_gt)
    out = {}
    strips = ((0, 8), (508, 516), (1016, 1024))
    for tag, dbl in (("f32", False), ("f64", True)):
        maps, loss, g0, _, _ = _run_optim(res, n, cl, tex0, target, 20, 0.01, False, dbl)
        out[f"loss_{tag}"] = loss
        out[f"grad0_chan_sum_{tag}"] = g0.double().sum((0, 2, 3)).numpy()
        out[f"grad0_chan_abs_{tag}"] = g0.double().abs().sum((0, 2, 3)).numpy()
        out[f"maps_chan_sum_{tag}"] = maps.double().sum((0, 2, 3)).numpy()
        out[f"maps_chan_sq_{tag}"] = (maps.double() ** 2).sum((0, 2, 3)).numpy()
        out[f"maps_strips_{tag}"] = np.concatenate([maps[:, :, a:b, :].numpy() for a, b in strips], 2)
        out[f"grad0_strips_{tag}"] = np.concatenate([g0[:, :, a:b, :].numpy() for a, b in strips], 2)
    out["target_chan_sum"] = target.double().sum((0, 2, 3)).numpy()
    out["target_sq_sum"] = float((target.double() ** 2).sum())
    path = os.path.join(OUT, "config2_1024x9_stats.npz")
    np.savez_compressed(path, res=res, n=n, seed_gt=1, seed_start=2, strips=np.array(strips), **out)
    print(f"config2 stats: loss {out['loss_f32'][0]:.6g} -> {out['loss_f32'][-1]:.6g}   ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    import sys
    only = set(sys.argv[1:])
    os.makedirs(OUT, exist_ok=True)
    th.set_num_threads(os.cpu_count() or 1)
    if only:
        # python -m oracle.make_golden config2 nonpow2   (regenerate a subset; the default regenerates the round-1 set)
        if "nonpow2" in only:
            # non-power-of-two resolutions through the TMA kernel (40^2, 48^2 texels are multiples of 4): (j + 0.5)/R is not
            # exactly (j + 0.5) * (1/R) there (microfacet.py:16-19)
            make_case("coloc_40x9", 40, 9, True, synth.random_textures(40, 31), synth.random_textures(40, 32))
            make_case("offaxis_48x9", 48, 9, False, synth.random_textures(48, 33), synth.random_textures(48, 34), epochs=10)
        if "config2" in only:
            make_stats_1024()
        return
    make_case("coloc_32x9", 32, 9, True, synth.random_textures(32, 1), synth.random_textures(32, 2))
    make_case("offaxis_32x9", 32, 9, False, synth.random_textures(32, 1), synth.random_textures(32, 2))
    make_case("edges_32x9", 32, 9, True, synth.random_textures(32, 1), synth.edge_case_textures(32))
    make_case("edges_offaxis_32x9", 32, 9, False, synth.random_textures(32, 1), synth.edge_case_textures(32))
    make_case("wellcond_32x9", 32, 9, True, synth.well_conditioned_textures(32, 21), synth.well_conditioned_textures(32, 22))
    make_case("light_32x9", 32, 9, True, synth.random_textures(32, 1), synth.random_textures(32, 2),
              epochs=10, optim_light=True, power0=th.tensor([1200.0, 1500.0, 1800.0]))
    make_case("coloc_24x16", 24, 16, True, synth.random_textures(24, 5), synth.random_textures(24, 6), epochs=5)
    make_stats_256()


if __name__ == "__main__":
    main()
