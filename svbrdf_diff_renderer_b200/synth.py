"""Seeded synthetic inputs for the per-pixel SVBRDF optimisation path.

Everything here is generated on the CPU with ``torch.manual_seed`` and copied to
the device by the caller, because CUDA and CPU RNG streams differ (SURVEY.md
§8(d)).  The distributions follow the reference's own random initialiser
(``/root/reference/src/svbrdf.py:32-39``) and its shipped 9-light geometry
(``/root/reference/data/random/render.json:5-24``).

Channel order of a texture stack is the reference's: diffuse 0:3, normal-xy
3:5, roughness 5, specular 6:9 (``/root/reference/src/svbrdf.py:38``).
"""

from __future__ import annotations

import math

import torch as th

# /root/reference/data/random/render.json:4,24
IM_SIZE_CM = 6.848
LIGHT_POW = (1500.0, 1500.0, 1500.0)
LIGHT_Z_CM = 16.0
LIGHT_SPAN_CM = 3.0


def grid_lights(n: int, dtype=th.float32) -> th.Tensor:
    """k x k grid of co-located light/camera positions at z = 16 cm.

    n = 9 reproduces render.json (rows y = +3, 0, -3; columns x = -3, 0, +3).
    n = 64 / 256 are the 8x8 / 16x16 grids of SURVEY.md §8(d).
    """
    k = int(round(math.sqrt(n)))
    if k * k != n:
        raise ValueError(f"grid_lights needs a square count, got {n}")
    xs = th.linspace(-LIGHT_SPAN_CM, LIGHT_SPAN_CM, k, dtype=th.float64)
    ys = th.linspace(LIGHT_SPAN_CM, -LIGHT_SPAN_CM, k, dtype=th.float64)
    pos = th.empty(n, 3, dtype=th.float64)
    for r in range(k):
        for c in range(k):
            pos[r * k + c, 0] = xs[c]
            pos[r * k + c, 1] = ys[r]
            pos[r * k + c, 2] = LIGHT_Z_CM
    return pos.to(dtype)


def calibration(n: int, colocated: bool = True, seed: int = 3):
    """Return ``cl = [camera_pos[N,3], light_pos[N,3], light_pow[3]]`` (fp32, CPU).

    ``colocated=False`` displaces every light from its camera by ``randn * 1.5``
    (seeded) — the parity-only geometry of SURVEY.md §8(d).
    """
    cam = grid_lights(n)
    if colocated:
        light = cam.clone()
    else:
        g = th.Generator().manual_seed(seed)
        light = cam + th.randn(n, 3, generator=g) * 1.5
        light[:, 2] = light[:, 2].abs().clamp(min=4.0)
    pw = th.tensor(LIGHT_POW, dtype=th.float32)
    return [cam, light, pw]


def random_textures(res: int, seed: int, dif=0.5, spe=0.04, rgh=0.2, height: int | None = None) -> th.Tensor:
    """One draw of the reference's ``init_from_randn`` distribution, ``[1,9,H,W]`` fp32 on CPU.

    Draw order (normal, diffuse, specular, roughness) and the clamps are those of
    ``/root/reference/src/svbrdf.py:33-36`` so that ``torch.manual_seed(seed)``
    followed by the reference initialiser gives the same tensor.
    """
    h = res if height is None else height
    g = th.Generator().manual_seed(seed)
    normal = (th.randn(1, 2, h, res, generator=g) / 4).clamp(-1, 1)
    diffuse = (th.randn(1, 3, h, res, generator=g) / 8 + dif).clamp(0, 1) * 2 - 1
    specular = (th.randn(1, 3, h, res, generator=g) / 32 + spe).clamp(0, 1) * 2 - 1
    roughness = (th.randn(1, 1, h, res, generator=g) / 16 + rgh).clamp(0, 1) * 2 - 1
    return th.cat((diffuse, normal, roughness, specular), 1).contiguous()


def edge_case_textures(res: int, seed: int = 7) -> th.Tensor:
    """Texture stack that exercises every clamp edge of the path (parity tests).

    Rows are split into bands: in-range random, values at exactly +-1, values
    outside [-1, 1], normals with nx^2+ny^2 >= 1, roughness -> 0, bright
    (saturating) and dark (floor) materials.
    """
    t = random_textures(res, seed)
    band = max(res // 8, 1)
    g = th.Generator().manual_seed(seed + 1)
    # exactly +-1
    t[:, :, 1 * band:2 * band, :] = th.where(th.rand(1, 9, band, res, generator=g) > 0.5, 1.0, -1.0)
    # outside [-1, 1]
    t[:, :, 2 * band:3 * band, :] = th.randn(1, 9, band, res, generator=g) * 1.5
    # grazing / over-unit normals
    t[:, 3:5, 3 * band:4 * band, :] = th.randn(1, 2, band, res, generator=g).clamp(-1, 1)
    # roughness -> 0 and exactly 0
    t[:, 5, 4 * band:5 * band, :] = -1.0 + th.rand(1, band, res, generator=g) * 0.05
    t[:, 5, 4 * band, :] = -1.0
    # very bright: white diffuse, high specular, smooth
    t[:, 0:3, 5 * band:6 * band, :] = 0.95
    t[:, 6:9, 5 * band:6 * band, :] = 0.5
    t[:, 5, 5 * band:6 * band, :] = -0.6
    # very dark: black diffuse, zero specular
    t[:, 0:3, 6 * band:7 * band, :] = -1.0
    t[:, 6:9, 6 * band:7 * band, :] = -1.0
    return t.contiguous()


def well_conditioned_textures(res: int, seed: int) -> th.Tensor:
    """|n_xy| <= 0.35, mid-range albedo, rough (alpha^2 >= 0.08) surfaces: no clamp edge is
    active and the GGX denominator c2*a2 + (1 - c2) never cancels, so a strict element-wise
    tolerance is expected to hold."""
    g = th.Generator().manual_seed(seed)
    normal = (th.randn(1, 2, res, res, generator=g) / 6).clamp(-0.35, 0.35)
    diffuse = (th.rand(1, 3, res, res, generator=g) * 0.3 + 0.2) * 2 - 1
    specular = (th.rand(1, 3, res, res, generator=g) * 0.04 + 0.02) * 2 - 1
    roughness = (th.rand(1, 1, res, res, generator=g) * 0.2 + 0.75) * 2 - 1
    return th.cat((diffuse, normal, roughness, specular), 1).contiguous()
