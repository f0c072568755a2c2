#!/usr/bin/env python
"""ORACLE recipe (test infrastructure): stage the reference's own files for the hot path into ``oracle/_ref/``.

The reference is pure Python, so there is nothing to compile; what ``oracle/_ref`` carries for it is a byte-for-byte copy
of the four files the path consists of (``src/microfacet.py``, ``src/optimization.py``, ``src/svbrdf.py`` and the
``src/imageio.py`` they import), taken from where they lie under ``/root/reference``.  ``oracle/_ref/`` is git-ignored
(reference sources never enter the history) but travels to the GPU box with the snapshot, so ``bench.py --impl reference``
and the ``cpu_baseline`` leg can time the UNMODIFIED ``Microfacet.eval`` there (``cpu_baseline.kind = "reference"``)
instead of the port.  Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present.

    python oracle/stage_ref.py            # copies + writes oracle/_ref/MANIFEST.json (sha256 of every file)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("SVBRDF_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ("src/microfacet.py", "src/optimization.py", "src/svbrdf.py", "src/imageio.py")


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        if verbose:
            print(f"[stage_ref] {SRC_ROOT} not present: nothing staged")
        return False
    manifest = {"source_root": SRC_ROOT, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(f"[stage_ref] staged {len(FILES)} reference files into {DST_ROOT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
