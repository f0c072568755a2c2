"""PNG <-> array helpers used by ``SvbrdfIO`` — the data formats either side of the hot path.

Behavioural mirror of ``/root/reference/src/imageio.py:11-111`` for the pieces
``optim_perpixel`` touches (cv2 decode, optional Lanczos-4 resize, 8/16-bit -> [0,1] float,
BGR->RGB, the three interpretations "srgb" / "rough" / "normal"; contact sheets).  Host-side,
runs once per job; not on the timed path.
"""

from __future__ import annotations

import cv2
import numpy as np


def imresize(im, dim):
    return cv2.resize(im, dim, interpolation=cv2.INTER_LANCZOS4)


def imread_raw(filename, dim=None):
    """Decode without normalisation (uint8/uint16 as stored), optionally resized."""
    im = cv2.imread(str(filename), flags=cv2.IMREAD_ANYDEPTH | cv2.IMREAD_UNCHANGED)
    if im is None:
        raise FileNotFoundError(f"[ERROR:imageio:imread] cannot read {filename}")
    if dim is not None:
        im = imresize(im, dim)
    return im


def imread(filename, flag=None, dim=None):
    im = imread_raw(filename, dim)
    if im.dtype == np.uint8:
        im = im.astype("float32") / 255
    elif im.dtype == np.uint16:
        im = im.astype("float32") / 65535
    else:
        im = im.astype("float32")

    three = im.ndim == 3 and im.shape[2] == 3
    if flag == "srgb":
        if not three:
            raise ValueError(f"[ERROR:imageio:imread:srgb] {filename} should be a 3 channel image")
        im = im[:, :, ::-1]
    elif flag == "rough":
        if three:
            im = im.mean(axis=2)
        elif im.ndim != 2:
            raise ValueError(f"[ERROR:imageio:imread:rough] {filename} should be a 3 or 1 channel image")
    elif flag == "normal":
        if not three:
            raise ValueError(f"[ERROR:imageio:imread:normal] {filename} should be a 3 channel image")
        im = im[:, :, ::-1] * 2 - 1
        im = im / np.linalg.norm(im, axis=2, keepdims=True)
    return np.ascontiguousarray(im)


def imwrite(im, filename, flag=None, dim=None):
    if dim is not None:
        im = imresize(im, dim)
    if flag == "srgb":
        im = im.clip(0, 1)[:, :, ::-1]
    elif flag == "rough":
        im = im.clip(0, 1)
    elif flag == "normal":
        im = ((im.clip(-1, 1) + 1) / 2)[:, :, ::-1]
    cv2.imwrite(str(filename), (im * 255).astype("uint8"))


def imwrite_u8(im, filename):
    """Write already-quantised bytes (BGR interleaved or single channel), e.g. from ``maps.encode_u8``."""
    if im.dtype != np.uint8:
        raise ValueError("imwrite_u8 takes uint8 arrays")
    cv2.imwrite(str(filename), im)


def imconcat(im_list, size=(2, 2)):
    w, h = size
    return cv2.vconcat([cv2.hconcat([im_list[r * w + c] for c in range(w)]) for r in range(h)])


def img9to1(folder):
    sheet = imconcat([imread(folder / f"{i:02d}.png") for i in range(9)], (3, 3))
    imwrite(sheet, folder / "all.png")


def tex4to1(folder):
    maps = [imread(folder / name) for name in ("nom.png", "dif.png", "spe.png", "rgh.png")]
    if maps[3].ndim == 2:
        maps[3] = np.dstack((maps[3],) * 3)
    imwrite(imconcat(maps), folder / "tex.png")
