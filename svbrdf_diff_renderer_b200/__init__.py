"""svbrdf_diff_renderer_b200 — B200-native (sm_100a) per-pixel SVBRDF optimisation path.

Drop-in for the hot path of tflsguoyu/svbrdf-diff-renderer: ``Microfacet``
(src/microfacet.py), ``Optim`` (src/optimization.py), ``SvbrdfOptim`` / ``SvbrdfIO``
(src/svbrdf.py) and ``optim_perpixel`` / ``render`` (src/scripts.py), backed by hand-written
CUDA kernels behind the C ABI of ``include/svbrdf_b200.h``.  No CPU fallback.
"""

from ._native import build_native, lib  # noqa: F401
from .microfacet import Microfacet  # noqa: F401
from .optimization import Optim  # noqa: F401
from .svbrdf import SvbrdfIO, SvbrdfOptim  # noqa: F401
from .scripts import optim_perpixel, optim_perpixel_pyramid, render  # noqa: F401
from . import maps  # noqa: F401
from .descriptor import VGGLoss  # noqa: F401

__all__ = ["Microfacet", "Optim", "SvbrdfOptim", "SvbrdfIO", "optim_perpixel", "optim_perpixel_pyramid", "render", "maps", "VGGLoss",
           "build_native", "lib"]
