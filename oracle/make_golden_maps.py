"""ORACLE tooling: generate ``tests/golden/maps_handoff_*.npz`` by running the UNMODIFIED reference's
``SvbrdfIO.save_textures_th`` / ``load_textures_th`` (``/root/reference/src/svbrdf.py:150-189``, which call
``/root/reference/src/imageio.py`` and ``cv2.resize(INTER_LANCZOS4)``) through real PNG files.

Run in the build container (needs ``/root/reference`` and cv2)::

    python -m oracle.make_golden_maps

Each fixture stores the input maps, the bytes the reference wrote (decoded back with cv2.imread, planar RGB order of
``oracle/maps_port.py``) and the maps the reference loaded at the next resolution.
"""

from __future__ import annotations

import os
import pathlib
import tempfile

import cv2
import numpy as np
import torch as th

from oracle import ref_loader
from svbrdf_diff_renderer_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {              # name: (res_in, res_out, seed, perturbation)
    "maps_handoff_24_to_48": (24, 48, 11, 0.3),        # the reference's 2x schedule (run.py:55-56)
    "maps_handoff_20_to_50": (20, 50, 12, 0.6),        # non-integer ratio, many clamped texels
    "maps_handoff_40_to_16": (40, 16, 13, 0.3),        # downscale
    "maps_handoff_32_to_32": (32, 32, 14, 0.3),        # same size: cv2.resize copies
}


def reference_roundtrip(tex, res_out):
    _, _, SvbrdfIO = ref_loader.load()
    io = SvbrdfIO.__new__(SvbrdfIO)
    io.device = th.device("cpu")
    with tempfile.TemporaryDirectory() as d:
        d = pathlib.Path(d)
        with ref_loader.quiet():
            io.save_textures_th(tex, d)
            files = {k: cv2.imread(str(d / f"{k}.png"), cv2.IMREAD_UNCHANGED) for k in ("nom", "dif", "spe", "rgh")}
            loaded = io.load_textures_th(d, res_out)
    rgb = lambda a: a[:, :, ::-1].transpose(2, 0, 1)  # noqa: E731
    planes = np.concatenate([rgb(files["dif"]), rgb(files["nom"]), files["rgh"][None], rgb(files["spe"])], 0)
    return np.ascontiguousarray(planes), loaded[0].numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (r_in, r_out, seed, noise) in CASES.items():
        tex = synth.random_textures(r_in, seed)
        g = th.Generator().manual_seed(seed)
        tex = (tex + th.randn(tex.shape, generator=g) * noise).clamp(-1, 1)      # what scripts.py:91 hands to save_textures_th
        planes, loaded = reference_roundtrip(tex, r_out)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), tex=tex[0].numpy(), planes_u8=planes, loaded=loaded, res_out=r_out)
        print(name, planes.shape, loaded.shape)


if __name__ == "__main__":
    main()
