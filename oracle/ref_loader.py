"""ORACLE helper (test infrastructure): import the UNMODIFIED reference classes.

Uses ``/root/reference`` where it exists (the build container) and otherwise the byte-for-byte copy of the path's
four files that ``oracle/stage_ref.py`` put into the git-ignored ``oracle/_ref/`` (that copy travels to the GPU box).  ``src.optimization`` imports ``matplotlib.pyplot`` for its loss plot
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
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/stage_ref.py (travels to the GPU box)
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "microfacet.py")) and os.path.isfile(os.path.join(STAGED_ROOT, "src", "microfacet.py")):
    REFERENCE_ROOT = STAGED_ROOT


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "microfacet.py"))


def staged() -> bool:
    """True when the classes come from the staged copy (``oracle/_ref``), i.e. not from the live reference tree."""
    return REFERENCE_ROOT == STAGED_ROOT


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
