"""ORACLE helper (test infrastructure): import the UNMODIFIED reference classes.

Works only where ``/root/reference`` exists (the build container; never on the
GPU box).  ``src.optimization`` imports ``matplotlib.pyplot`` for its loss plot
(``/root/reference/src/optimization.py:7``), which is not installed here, so a
stub module is injected first.  ``src.scripts`` is never imported (it pulls in
``pupil_apriltags`` and ``mitsuba``, both absent — SURVEY.md §8(c)).
"""

from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SVBRDF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "microfacet.py"))


def load():
    """Return (Microfacet, SvbrdfOptim, SvbrdfIO) from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            for name in ("figure", "plot", "xlim", "legend", "title", "savefig", "close"):
                setattr(plt, name, lambda *a, **k: None)
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from src.microfacet import Microfacet
    from src.svbrdf import SvbrdfOptim, SvbrdfIO
    return Microfacet, SvbrdfOptim, SvbrdfIO


@contextlib.contextmanager
def quiet():
    """The reference prints a ``[DONE:...]`` line per call; keep test logs clean."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def to_double(renderer):
    """fp64 arbiter: widen the renderer's plain attributes."""
    for name in ("pos", "camera_pos", "light_pos", "light_pow"):
        setattr(renderer, name, getattr(renderer, name).double())
    return renderer
