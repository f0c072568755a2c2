"""ORACLE (test infrastructure, not product code).

CPU restatement, in stock torch tensor ops, of the reference's per-pixel SVBRDF
optimisation path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker or as the timed CPU baseline — never as a
product path.

What is restated (reference file:line):

* planar-sample geometry            ``/root/reference/src/microfacet.py:16-24``
* texel -> material maps            ``/root/reference/src/microfacet.py:64-79``
* GGX / Schlick-SG / Smith terms    ``/root/reference/src/microfacet.py:28-32,43-52``
* the shading integral + gamma      ``/root/reference/src/microfacet.py:84-120``
* clamp -> render -> MSE -> Adam    ``/root/reference/src/svbrdf.py:44-71``,
                                    ``/root/reference/src/optimization.py:14,28-29``

The arithmetic keeps the reference's tensor shapes ([N,3,R,R] everywhere, scalar
quantities replicated over the 3 channels) and its operation order, so in fp32
it reproduces the reference bit for bit on the same torch build, and its CPU
cost (one full-size temporary per op, autograd graph of the same depth) is the
reference's.  The backward pass is torch autograd and the optimiser is
``torch.optim.Adam`` — third-party arithmetic the reference also delegates to
(torch 2.11.0+cu128 in this image; the reference README pins 2.7.0).

Pinning: ``tests/test_oracle_pin.py`` checks this port against the unmodified
reference imported from ``/root/reference`` (when that tree is present, i.e. in
the build container) and against ``tests/golden/*.npz`` generated from the
reference by ``oracle/make_golden.py``.  The reference ships no tests or golden
vectors of its own, so reference-generated fixtures are the pin.
"""

from __future__ import annotations

import math

import torch as th

EPS = 1e-6          # /root/reference/src/microfacet.py:14
GAMMA = 2.2


class Scene:
    """Texture-independent geometry of one capture: positions broadcast to [N,3,R,R].

    Follows ``/root/reference/src/microfacet.py:16-24``.  ``dtype`` selects the
    fp32 oracle or the fp64 arbiter (same code, every attribute in double).
    """

    def __init__(self, res, cam, light, power, size, dtype=th.float32, rows=None, device=None):
        n = cam.shape[0]
        self.res, self.n, self.dtype = res, n, dtype
        ticks = th.arange(res, dtype=th.float32)
        ticks = ((ticks + 0.5) / res - 0.5) * size          # fp32 like the reference, then widened
        gx, gy = th.meshgrid(ticks, ticks, indexing="xy")
        plane = th.stack((gx, -gy, th.zeros_like(gx)), 2).permute(2, 0, 1).to(dtype)
        if device is not None:                               # the reference on `cuda:0` (scripts.py:68): bench.py's eager side line
            plane, cam, light, power = plane.to(device), cam.to(device), light.to(device), power.to(device)
        if rows is not None:                                 # row band [r0, r1) of the full image
            plane = plane[:, rows[0]:rows[1], :]
        h, w = plane.shape[1], plane.shape[2]
        self.h, self.w = h, w
        self.plane = plane.unsqueeze(0).expand(n, -1, -1, -1)
        self.cam = cam.to(dtype)[:, :, None, None].expand(-1, -1, h, w)
        self.light = light.to(dtype)[:, :, None, None].expand(-1, -1, h, w)
        self.set_power(power)

    def set_power(self, power):
        # /root/reference/src/microfacet.py:24,81-82 (keeps the autograd link to `power`)
        self.power = power.to(self.dtype)[None, :, None, None].expand(self.n, -1, self.h, self.w)


def _unit(vec):
    return vec / vec.norm(2.0, 1, keepdim=True)


def _dot3(a, b):
    return (a * b).sum(1, keepdim=True).expand(-1, 3, -1, -1)


def _direction(scene, point):
    delta = point - scene.plane
    return _unit(delta), _dot3(delta, delta)


def texel_maps(scene, tex):
    """[1,9,H,W] -> normal, diffuse, specular, roughness, each [N,3,H,W] (microfacet.py:64-79)."""
    n = scene.n
    nx = tex[:, 3, :, :].clamp(-1, 1)
    ny = tex[:, 4, :, :].clamp(-1, 1)
    planar = (nx ** 2 + ny ** 2).clamp(0, 1 - EPS)
    nz = (1 - planar).sqrt()
    normal = _unit(th.stack((nx, ny, nz), 1)).expand(n, -1, -1, -1)
    diffuse = (((tex[:, 0:3, :, :] + 1) / 2) ** GAMMA).expand(n, -1, -1, -1)
    rough = (((tex[:, 5, :, :] + 1) / 2) ** GAMMA).expand(n, 3, -1, -1)
    specular = (((tex[:, 6:9, :, :] + 1) / 2) ** GAMMA).expand(n, -1, -1, -1)
    return normal, diffuse, specular, rough


def shade(scene, tex):
    """The reference's ``Microfacet.eval`` (microfacet.py:84-120): [1,9,H,W] -> [N,3,H,W]."""
    normal, diffuse, specular, rough = texel_maps(scene, tex)

    v, _ = _direction(scene, scene.cam)
    l, dist2 = _direction(scene, scene.light)
    h = _unit(l + v)

    ndv = _dot3(normal, v).clamp(min=0)
    ndl = _dot3(normal, l).clamp(min=0)
    ndh = _dot3(normal, h).clamp(min=0)
    vdh = _dot3(v, h).clamp(min=0)

    lambert = diffuse / math.pi
    lambert = lambert * (1 - specular)

    # GGX, microfacet.py:28-32 (alpha = rough^2 is formed once per use, as at microfacet.py:106,108)
    c2 = ndh ** 2
    a2 = (rough ** 2) ** 2
    den = c2 * a2 + (1 - c2)
    ndf = a2 / (math.pi * den ** 2 + EPS)
    # spherical-gaussian Schlick, microfacet.py:43-45
    sphg = th.pow(2.0, ((-5.55473 * vdh) - 6.98316) * vdh)
    fresnel = specular + (1.0 - specular) * sphg
    # Smith-Schlick, microfacet.py:47-52
    k = (rough ** 2) * 0.5 + EPS
    geo = (ndv / (ndv * (1.0 - k) + k)) * (ndl / (ndl * (1.0 - k) + k))
    spec = ndf * fresnel * geo / (4 * ndv * ndl + EPS)

    brdf = 1 * lambert + 1 * spec
    radiance = scene.power * brdf * ndl / dist2
    return radiance.clamp(EPS, 1) ** (1 / GAMMA)


def l2_loss(pred, target):
    """``torch.nn.MSELoss()`` with mean reduction (optimization.py:14,28-29)."""
    return th.nn.functional.mse_loss(pred, target)


def loss_and_grad(scene, tex, target, power=None):
    """One forward + backward of clamp -> shade -> MSE (svbrdf.py:60-70).

    Returns (loss, dL/dtex, dL/dpower or None, image).  ``tex`` is the *unclamped*
    parameter; the clamp of svbrdf.py:60 is part of the differentiated graph.
    """
    tex = tex.detach().clone().to(scene.dtype).requires_grad_(True)
    pw = None
    if power is not None:
        pw = power.detach().clone().to(scene.dtype).requires_grad_(True)
        scene.set_power(pw)
    img = shade(scene, tex.clamp(-1, 1))
    loss = l2_loss(img, target.to(scene.dtype))
    loss.backward()
    return loss.detach(), tex.grad.detach(), (pw.grad.detach() if pw is not None else None), img.detach()


def image_grad(scene, tex, grad_img, power=None):
    """Vector-Jacobian product of ``shade`` for an arbitrary upstream ``grad_img`` (mode B)."""
    tex = tex.detach().clone().to(scene.dtype).requires_grad_(True)
    pw = None
    if power is not None:
        pw = power.detach().clone().to(scene.dtype).requires_grad_(True)
        scene.set_power(pw)
    img = shade(scene, tex)
    img.backward(grad_img.to(scene.dtype))
    return tex.grad.detach(), (pw.grad.detach() if pw is not None else None), img.detach()


def optimise(scene, tex0, target, epochs, lr, power=None, optim_light=False, on_epoch=None, loss_scale=None):
    """The loop body of ``SvbrdfOptim.optim`` (svbrdf.py:48-71) without tqdm and dumps.

    Returns (final unclamped textures, list of per-epoch losses, final power).

    ``loss_scale`` (big-config oracle only): when ``scene`` is a row band of a larger image, the
    band's MSE mean is rescaled by rows_in_band/rows_in_image so the gradient — and with it Adam's eps-dependent
    step — is the full image's (texels are independent; the loss returned is then the band's SHARE of the full loss).
    """
    tex = tex0.detach().clone().to(scene.dtype).requires_grad_(True)
    target = target.to(scene.dtype)
    params = [tex]
    pw = None
    if optim_light:
        pw = power.detach().clone().to(scene.dtype).requires_grad_(True)
        params.append(pw)
    opt = th.optim.Adam(params, lr=lr, betas=(0.9, 0.999))
    losses = []
    for epoch in range(epochs):
        if optim_light:
            scene.set_power(pw)
        img = shade(scene, tex.clamp(-1, 1))
        loss = l2_loss(img, target)
        if loss_scale is not None:
            loss = loss * loss_scale
        losses.append(loss.item())
        opt.zero_grad()
        loss.backward()
        opt.step()
        if on_epoch is not None:
            on_epoch(epoch, tex, loss)
    return tex.detach(), losses, (pw.detach() if pw is not None else None)


def shade_chunked(res, cam, light, power, size, tex, chunk, dtype=th.float32, rows=None):
    """Forward render evaluated ``chunk`` lights at a time (big-config oracle)."""
    outs = []
    with th.no_grad():
        for s in range(0, cam.shape[0], chunk):
            sc = Scene(res, cam[s:s + chunk], light[s:s + chunk], power, size, dtype, rows)
            outs.append(shade(sc, tex.to(dtype)))
    return th.cat(outs, 0)


def loss_and_grad_chunked(res, cam, light, power, size, tex, target, chunk, dtype=th.float32, rows=None):
    """Loss/gradient accumulated over light chunks; lights are independent so
    sum_chunks(n_chunk/N * chunk result) equals the one-piece result."""
    n = cam.shape[0]
    loss = th.zeros((), dtype=dtype)
    grad = None
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        sc = Scene(res, cam[s:e], light[s:e], power, size, dtype, rows)
        l_c, g_c, _, _ = loss_and_grad(sc, tex, target[s:e])
        wgt = (e - s) / n
        loss = loss + l_c * wgt
        grad = g_c * wgt if grad is None else grad + g_c * wgt
    return loss, grad
