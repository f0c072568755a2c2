#!/usr/bin/env python
"""Times the map hand-off kernels (encode / Lanczos-4 resize / decode) against their HBM floor and against the host
path they replace (the PNG round trip through cv2 + numpy).  Development/measurement tool:

    python tools/maps_bench.py [--res 1024]      # hand-off res -> 2*res
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from svbrdf_diff_renderer_b200 import maps, synth  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--peak", type=float, default=6540.8, help="measured HBM GB/s (MEASURED_PEAKS.json)")
    a = ap.parse_args()
    dev = th.device("cuda:0")
    r, R = a.res, 2 * a.res
    pool = [synth.random_textures(r, 40 + i).to(dev) for i in range(8)]      # cycled: 8 x 37.7 MB at 1024^2 > L2
    planes = [maps.encode_u8(t) for t in pool]
    ups = [maps.resize_lanczos4_u8(p, R, R) for p in planes[:4]]
    k = [0]

    def nxt(lst):
        k[0] += 1
        return lst[k[0] % len(lst)]

    rows = {}
    us = timed(lambda: maps.encode_u8(nxt(pool)), a.reps)
    rows["encode"] = (us, (36 + 10) * r * r)
    us = timed(lambda: maps.resize_lanczos4_u8(nxt(planes), R, R), a.reps)
    rows["resize_lanczos4"] = (us, 10 * (r * r + R * R))
    us = timed(lambda: maps.decode_u8(nxt(ups)), a.reps)
    rows["decode"] = (us, (10 + 36) * R * R)
    us = timed(lambda: maps.handoff(nxt(pool), R), a.reps)
    rows["handoff (3 launches)"] = (us, (36 + 10) * r * r + 10 * (r * r + R * R) + (10 + 36) * R * R)
    # host path: what the reference does — SvbrdfIO.save_textures_th -> PNG files -> load_textures_th(dir, 2r) with cv2 and
    # numpy on the host (svbrdf.py:150-189), here through this package's own host route (on_device=False), files on /tmp
    import pathlib
    import tempfile

    import svbrdf_diff_renderer_b200 as pkg
    io = pkg.SvbrdfIO.__new__(pkg.SvbrdfIO)
    io.device = th.device("cpu")
    t_host = pool[0].clamp(-1, 1).cpu()
    with tempfile.TemporaryDirectory() as d:
        d = pathlib.Path(d)
        t0 = time.perf_counter()
        io.save_textures_th(t_host, d)
        out = io.load_textures_th(d, R, on_device=False)
        host_ms = (time.perf_counter() - t0) * 1e3
    assert th.equal(out, maps.handoff(pool[0], R).cpu()), "device hand-off differs from the host round trip"
    res = {"res_in": r, "res_out": R, "host_png_round_trip_ms": round(host_ms, 2), "peak_GBs": a.peak, "kernels": {}}
    for name, (us, nbytes) in rows.items():
        gbs = nbytes / (us * 1e-6) / 1e9
        res["kernels"][name] = {"us": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2), "GBs": round(gbs, 1), "frac_of_peak": round(gbs / a.peak, 3)}
        print(f"{name:24s} {us:9.2f} us  {nbytes / 1e6:8.2f} MB  {gbs:8.1f} GB/s  {gbs / a.peak * 100:5.1f} % of measured HBM")
    print(f"host PNG round trip (save_textures_th + load_textures_th, cv2/numpy): {host_ms:.1f} ms; device result identical")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
