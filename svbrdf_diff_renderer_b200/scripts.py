"""Drop-in for the two workflows of the reference's ``src/scripts.py`` that sit on the hot path:
``optim_perpixel`` (scripts.py:67-97) and the forward-only ``render`` (scripts.py:31-41).

Same signatures and files written; the device is ``cuda:0`` — unlike the reference
(scripts.py:68) there is no CPU branch: without a GPU the call raises.
"""

from __future__ import annotations

import torch as th

from . import maps
from .microfacet import Microfacet
from .svbrdf import SvbrdfIO, SvbrdfOptim


def _device():
    if not th.cuda.is_available():
        raise RuntimeError("svbrdf_diff_renderer_b200 needs a CUDA device (B200); there is no CPU fallback")
    return th.device("cuda:0")


def render(json_dir, res):
    device = _device()
    svbrdf_obj = SvbrdfIO(json_dir, device)
    textures = svbrdf_obj.load_textures_th(svbrdf_obj.reference_dir, res)
    render_obj = Microfacet(res, svbrdf_obj.n_of_imgs, svbrdf_obj.im_size, svbrdf_obj.cl, device)
    with th.no_grad():
        rendereds = render_obj.eval(textures)
    svbrdf_obj.save_images_th(rendereds, svbrdf_obj.target_dir)


def optim_perpixel(json_dir, res, lr, epochs, tex_init, optim_light=False, uint8_targets="auto"):
    """scripts.py:67-97.  ``uint8_targets``: ``"auto"`` (default) keeps the target PNGs as their bytes whenever they are 8-bit
    RGB files (imageio.py:18-19 divides them by 255; the fused kernel does the same division, correctly rounded, so the
    optimisation is bit-identical to the float32 route at a quarter of the upload and target traffic); ``False`` forces the
    reference's float32 stack, ``True`` insists on bytes."""
    device = _device()

    svbrdf_obj = SvbrdfIO(json_dir, device)
    targets = svbrdf_obj.load_images_th(svbrdf_obj.target_dir, res, as_uint8=uint8_targets)

    renderer_obj = Microfacet(res, svbrdf_obj.n_of_imgs, svbrdf_obj.im_size, svbrdf_obj.cl, device)

    optim_obj = SvbrdfOptim(device, renderer_obj)
    optim_obj.load_targets(targets)

    if tex_init == "random":
        optim_obj.init_from_randn()
    elif tex_init == "const":
        optim_obj.init_from_const()
    elif tex_init == "textures":
        optim_obj.init_from_tex(svbrdf_obj.load_textures_th(svbrdf_obj.reference_dir, res))
    elif isinstance(tex_init, th.Tensor):
        # device-side coarse-to-fine hand-off: the previous resolution's maps, quantised, Lanczos-resized and decoded on
        # the GPU exactly as the reference's save_textures_th -> load_textures_th(res) round trip does through PNG files
        optim_obj.init_from_tex(maps.handoff(tex_init.to(device), res))
    else:
        raise ValueError(f"tex_init must be 'random', 'const', 'textures' or a [1,9,r,r] tensor, got {tex_init!r}")

    optim_obj.optim(epochs, lr, svbrdf_obj, optim_light)

    with th.no_grad():
        final = optim_obj.textures.detach().clamp(-1, 1)
        svbrdf_obj.save_textures_th(final, svbrdf_obj.optimize_dir)
        if optim_light:
            print("Optimized light: ", svbrdf_obj.cl[2])
            renderer_obj.update_light(svbrdf_obj.cl[2])
        svbrdf_obj.save_images_th(renderer_obj.eval(final), svbrdf_obj.rerender_dir)
    return optim_obj


def optim_perpixel_pyramid(stages, lr, epochs, tex_init="const", optim_light=False, uint8_targets="auto"):
    """The reference's coarse-to-fine recipe (run.py:55-56: 256 -> 512 -> 1024, each stage initialised with the previous
    stage's maps) with the hand-off kept on the device.  ``stages`` is a list of ``(json_dir, res)``; every stage is
    one ``optim_perpixel`` call (same files written), stage k+1 starts from ``maps.handoff(stage k maps, res)`` — bit
    for bit what ``tex_init="textures"`` loads when ``reference_dir`` of stage k+1 is ``optimize_dir`` of stage k."""
    prev, out = tex_init, []
    for json_dir, res in stages:
        o = optim_perpixel(json_dir, res, lr, epochs, prev, optim_light, uint8_targets)
        prev = o.textures.detach()
        out.append(o)
    return out
