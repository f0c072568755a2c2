"""Drop-in for the reference's ``src/microfacet.py`` (class ``Microfacet``).

Same constructor, attributes and methods as ``/root/reference/src/microfacet.py:9-120``;
``eval`` runs the hand-written sm_100a kernels through the C ABI
(``include/svbrdf_b200.h``) instead of ~270 torch element-wise ops:

* forward  -> ``svbrdf_render_fwd``  (one pass, writes only the [N,3,R,R] image)
* backward -> ``svbrdf_render_bwd``  (recomputes the forward per light; saves only
  the [1,9,R,R] textures and the 3-float light power between passes, instead of the
  reference's 48 image-sized autograd buffers — SURVEY.md Appendix C)

It is differentiable w.r.t. ``textures`` and, through ``update_light`` (microfacet.py:81-82),
w.r.t. ``light_pow``, so any torch loss can sit on top (the MaterialGAN latent optimiser and
the VGG descriptor loss call it this way: materialgan.py:136-146).

There is no CPU path: constructing on a non-CUDA device raises.
"""

from __future__ import annotations

import ctypes
import math
import os

import torch as th

from . import _native as nv


def _log(msg):
    if not os.environ.get("SVBRDF_B200_QUIET"):
        print(msg)


class _RenderFn(th.autograd.Function):
    """``[1,9,R,R] textures, [3] light_pow -> [N,3,R,R]`` through the native kernels."""

    @staticmethod
    def forward(ctx, textures, light_pow, renderer):
        tex = nv.dev_f32(textures.detach().contiguous(), "textures")
        pw = nv.dev_f32(light_pow.detach().contiguous(), "light_pow")
        out = th.empty(renderer.n_of_imgs, 3, renderer.res, renderer.res, dtype=th.float32, device=tex.device)
        geom = renderer._geom(pw)
        nv.check(nv.lib().svbrdf_render_fwd(ctypes.byref(geom), nv.ptr(tex), nv.ptr(out), nv.stream_ptr(tex.device)),
                 "svbrdf_render_fwd")
        ctx.renderer = renderer
        ctx.save_for_backward(tex, pw)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        tex, pw = ctx.saved_tensors
        renderer = ctx.renderer
        want_tex, want_pow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_tex or want_pow):
            return None, None, None
        gout = nv.dev_f32(grad_out.contiguous(), "grad_out")
        grad_tex = th.empty_like(tex)
        grad_pow = th.empty(3, dtype=th.float32, device=tex.device) if want_pow else None
        ws = renderer._workspace()
        geom = renderer._geom(pw)
        nv.check(nv.lib().svbrdf_render_bwd(ctypes.byref(geom), nv.ptr(tex), nv.ptr(gout), nv.ptr(grad_tex), nv.ptr(grad_pow),
                                            nv.ptr(ws), nv.stream_ptr(tex.device)), "svbrdf_render_bwd")
        return (grad_tex if want_tex else None), grad_pow, None


class _RenderNormL2Fn(th.autograd.Function):
    """``textures, light_pow -> (normalised image [N,3,R,R], L2 loss vs targets)`` in one native pass each way: the
    render as the mode-B consumers use it (materialgan.py:136-147: ``compute_image_loss(rendereds)`` plus
    ``VGGLoss(rendereds)``, whose first step is the per-channel normalisation of descriptor.py:65-75)."""

    @staticmethod
    def forward(ctx, textures, light_pow, renderer, targets, mean, std):
        tex = nv.dev_f32(textures.detach().contiguous(), "textures")
        pw = nv.dev_f32(light_pow.detach().contiguous(), "light_pow")
        n, res = renderer.n_of_imgs, renderer.res
        out = th.empty(n, 3, res, res, dtype=th.float32, device=tex.device)
        loss = th.zeros((), dtype=th.float32, device=tex.device)
        tcode = 0
        if targets is not None:
            if not targets.is_cuda or not targets.is_contiguous() or tuple(targets.shape) != (n, 3, res, res):
                raise RuntimeError(f"targets must be a contiguous CUDA tensor [{n},3,{res},{res}]")
            tcode = nv.target_dtype_code(targets)
        f3 = ctypes.c_float * 3
        ctx.mean, ctx.std = f3(*[float(x) for x in mean]), f3(*[float(x) for x in std])
        geom = renderer._geom(pw)
        nv.check(nv.lib().svbrdf_render_norm_l2_fwd(ctypes.byref(geom), nv.ptr(tex), ctx.mean, ctx.std, nv.ptr(targets), tcode, nv.ptr(out),
                                                    nv.ptr(loss) if targets is not None else None, nv.ptr(renderer._workspace()),
                                                    nv.stream_ptr(tex.device)), "svbrdf_render_norm_l2_fwd")
        ctx.renderer, ctx.targets, ctx.tcode = renderer, targets, tcode
        ctx.save_for_backward(tex, pw)
        return out, loss

    @staticmethod
    def backward(ctx, grad_out, grad_loss):
        tex, pw = ctx.saved_tensors
        renderer = ctx.renderer
        want_tex, want_pow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_tex or want_pow):
            return None, None, None, None, None, None
        gout = th.zeros_like(tex.new_empty(renderer.n_of_imgs, 3, renderer.res, renderer.res)) if grad_out is None else \
            nv.dev_f32(grad_out.contiguous(), "grad_out")
        l2g = None
        if ctx.targets is not None and grad_loss is not None:
            l2g = nv.dev_f32(grad_loss.reshape(1).contiguous(), "grad_loss")
        grad_tex = th.empty_like(tex)
        grad_pow = th.empty(3, dtype=th.float32, device=tex.device) if want_pow else None
        geom = renderer._geom(pw)
        nv.check(nv.lib().svbrdf_render_norm_l2_bwd(ctypes.byref(geom), nv.ptr(tex), ctx.std, nv.ptr(gout), nv.ptr(ctx.targets), ctx.tcode,
                                                    nv.ptr(l2g), nv.ptr(grad_tex), nv.ptr(grad_pow), nv.ptr(renderer._workspace()),
                                                    nv.stream_ptr(tex.device)), "svbrdf_render_norm_l2_bwd")
        return (grad_tex if want_tex else None), grad_pow, None, None, None, None


class Microfacet:
    """Cook-Torrance renderer of a planar sample under N point lights (microfacet.py:9-120)."""

    def __init__(self, res, n, size, cl, device):
        device = th.device(device)
        if device.type != "cuda":
            raise RuntimeError("svbrdf_diff_renderer_b200.Microfacet needs a CUDA device: the B200 path has no CPU fallback")
        nv.lib()                                   # fail now, loudly, if the extension cannot be loaded
        self.res = res
        self.n_of_imgs = n
        self.f0 = 0.04
        self.eps = 1e-6
        self.size = float(size)
        self.device = device

        self._cam = nv.dev_f32(cl[0].to(device=device, dtype=th.float32).contiguous(), "camera_pos")
        self._light = nv.dev_f32(cl[1].to(device=device, dtype=th.float32).contiguous(), "light_pos")
        if self._cam.shape != (n, 3) or self._light.shape != (n, 3):
            raise RuntimeError(f"camera_pos/light_pos must be [{n},3]")
        self._ws = None
        self.update_light(cl[2])

        # Attributes the reference exposes (microfacet.py:16-24); views, no image-sized memory.
        ticks = th.arange(res, dtype=th.float32, device=device)
        ticks = ((ticks + 0.5) / res - 0.5) * size
        gx, gy = th.meshgrid(ticks, ticks, indexing="xy")
        self.pos = th.stack((gx, -gy, th.zeros_like(gx)), 0).unsqueeze(0).expand(n, -1, -1, -1)
        self.camera_pos = self._cam[:, :, None, None].expand(-1, -1, res, res)
        self.light_pos = self._light[:, :, None, None].expand(-1, -1, res, res)

        _log("[DONE:Microfacet] Initial object")

    # ---- native plumbing -------------------------------------------------------------------
    def _geom(self, pw, rows=None, row_offset=0, n_lights=None, light_offset=0, plane_stride=0):
        n = self.n_of_imgs if n_lights is None else n_lights
        return nv.Geom(self._cam.data_ptr() + 12 * light_offset, self._light.data_ptr() + 12 * light_offset, pw.data_ptr(),
                       self.size, self.res, self.res if rows is None else rows, row_offset, n, plane_stride)

    def _workspace(self):
        if self._ws is None:
            self._ws = nv.workspace(self.res, self.res, self.device)
        return self._ws

    # ---- reference API ---------------------------------------------------------------------
    def update_light(self, light_pow):
        """microfacet.py:81-82 — rebinding keeps the autograd link to ``light_pow``."""
        self._pow = light_pow if (light_pow.is_cuda and light_pow.dtype == th.float32) else \
            light_pow.to(device=self.device, dtype=th.float32)
        self.light_pow = self._pow[None, :, None, None].expand(self.n_of_imgs, -1, self.res, self.res)

    def eval(self, textures):
        assert (textures.shape[2] == textures.shape[3])
        assert (self.res == textures.shape[2])
        if textures.shape[0] != 1 or textures.shape[1] != 9:
            raise RuntimeError(f"textures must be [1,9,{self.res},{self.res}], got {tuple(textures.shape)}")
        return _RenderFn.apply(textures, self._pow, self)

    def eval_normalized(self, textures, mean, std, targets=None):
        """``eval`` fused with what the mode-B consumers do next (not in the reference API):
        returns ``((eval(textures) - mean[c]) / std[c], MSELoss(eval(textures), targets))`` — the input of the feature
        network (descriptor.py:65-79) and the image loss (optimization.py:28-29) — from ONE native pass, differentiable
        through one native backward pass that takes both upstream gradients.  ``targets`` may be float32 or uint8."""
        assert (textures.shape[2] == textures.shape[3])
        assert (self.res == textures.shape[2])
        if textures.shape[0] != 1 or textures.shape[1] != 9:
            raise RuntimeError(f"textures must be [1,9,{self.res},{self.res}], got {tuple(textures.shape)}")
        return _RenderNormL2Fn.apply(textures, self._pow, self, targets, tuple(mean), tuple(std))

    # ---- helpers kept for API compatibility (no caller outside eval in the reference) ----------
    def GGX(self, cos_h, alpha):
        c2, a2 = cos_h * cos_h, alpha * alpha
        d = c2 * a2 + (1 - c2)
        return a2 / (math.pi * d * d + self.eps)

    def Beckmann(self, cos_h, alpha):
        c2, a2 = cos_h * cos_h, alpha * alpha
        return th.exp(-((1 - c2) / c2) / a2) / (math.pi * a2 * c2 * c2)

    def Fresnel_f0(self, cos, f0):
        return f0 + (1 - f0) * (1 - cos) ** 5

    def Fresnel(self, cos, specular):
        return specular + (1.0 - specular) * th.exp2((-5.55473 * cos - 6.98316) * cos)

    def Smith(self, n_dot_v, n_dot_l, alpha):
        k = alpha * 0.5 + self.eps
        return (n_dot_v / (n_dot_v * (1.0 - k) + k)) * (n_dot_l / (n_dot_l * (1.0 - k) + k))

    def dot(self, a, b):
        return (a * b).sum(1, keepdim=True).expand(-1, 3, -1, -1)

    def normalize(self, vec):
        return vec / vec.norm(2.0, 1, keepdim=True)

    def get_dir(self, pos):
        vec = pos - self.pos
        return self.normalize(vec), self.dot(vec, vec)

    def reconstruct_normal(self, texture):
        xy = texture[:, 0:2, :, :].clamp(-1, 1)
        z = (1 - (xy * xy).sum(1, keepdim=True).clamp(0, 1 - self.eps)).sqrt()
        return self.normalize(th.cat((xy, z), 1))

    def tex2map(self, textures):
        n = self.n_of_imgs
        gamma = lambda t: ((t + 1) / 2) ** 2.2  # noqa: E731
        normal = self.reconstruct_normal(textures[:, 3:5, :, :]).expand(n, -1, -1, -1)
        diffuse = gamma(textures[:, 0:3, :, :]).expand(n, -1, -1, -1)
        roughness = gamma(textures[:, 5:6, :, :]).expand(n, 3, -1, -1)
        specular = gamma(textures[:, 6:9, :, :]).expand(n, -1, -1, -1)
        return normal, diffuse, specular, roughness
