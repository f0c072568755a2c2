#!/usr/bin/env python
"""Where the per-step end-to-end time of bench.py's `e2e` leg goes (1024^2 x 9): upload alone, kernel alone, read-back,
the full step — wall clock and CUDA events, uint8 and float32 targets, one or several copy chunks."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
import torch as th
import svbrdf_diff_renderer_b200 as pkg
from svbrdf_diff_renderer_b200 import _native as nv, synth

dev = th.device("cuda:0")
res, n = 1024, 9
L = nv.lib()
r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in synth.calibration(n)], dev)
with th.no_grad():
    tgt = r.eval(synth.random_textures(res, 1).to(dev)).contiguous()
tex = synth.random_textures(res, 2)[0].to(dev).contiguous()
m, v = th.zeros_like(tex), th.zeros_like(tex)
geom, ws, st = r._geom(r._pow), r._workspace(), nv.stream_ptr(dev)
loss_dev = th.zeros(1, device=dev)
host_loss = th.empty(1).pin_memory()
host_f32 = [tgt.cpu().pin_memory() for _ in range(2)]
host_u8 = [(tgt * 255).to(th.uint8).cpu().pin_memory() for _ in range(2)]
stage_f32 = th.empty_like(tgt)
stage_u8 = th.empty(tgt.shape, dtype=th.uint8, device=dev)


def wall(fn, k=20):
    for i in range(3):
        fn(i)
    th.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(k):
        fn(i)
    th.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


def kernel(i, stage, code):
    a = nv.Adam(0.01, 0.9, 0.999, 1e-8, i + 1)
    nv.check(L.svbrdf_l2_adam_step(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(stage), code, ctypes.byref(a),
                                   nv.ptr(loss_dev), None, nv.ptr(ws), st), "step")


for name, host, stage, code in (("u8", host_u8, stage_u8, 1), ("f32", host_f32, stage_f32, 0)):
    nbytes = stage.numel() * stage.element_size()
    t_copy = wall(lambda i: (stage.copy_(host[i % 2], non_blocking=True), th.cuda.current_stream().synchronize()))
    t_copy_async = wall(lambda i: stage.copy_(host[i % 2], non_blocking=True))
    t_kernel = wall(lambda i: (kernel(i, stage, code), th.cuda.current_stream().synchronize()))
    t_kernel_async = wall(lambda i: kernel(i, stage, code))

    def full(i):
        stage.copy_(host[i % 2], non_blocking=True)
        kernel(i, stage, code)
        host_loss.copy_(loss_dev, non_blocking=True)
        th.cuda.current_stream().synchronize()

    t_full = wall(full)
    # the same bytes in 4 chunks (different DMA descriptors)
    hv = [h.view(-1) for h in host]
    sv = stage.view(-1)
    q = sv.numel() // 4

    def chunked(i):
        for c in range(4):
            sv[c * q:(c + 1) * q].copy_(hv[i % 2][c * q:(c + 1) * q], non_blocking=True)
        th.cuda.current_stream().synchronize()

    t_chunk = wall(chunked)
    # cudaMemcpyAsync through the runtime directly
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def raw(i):
        rt.cudaMemcpyAsync(stage.data_ptr(), host[i % 2].data_ptr(), nbytes, 1, st)
        th.cuda.current_stream().synchronize()

    t_raw = wall(raw)
    print(f"{name}: {nbytes / 1e6:.1f} MB  copy+sync {t_copy:.3f} ms ({nbytes / t_copy / 1e6:.1f} GB/s)  copy enqueue {t_copy_async:.3f}  raw cudaMemcpyAsync+sync {t_raw:.3f}  "
          f"4 chunks {t_chunk:.3f}  kernel+sync {t_kernel:.3f}  kernel enqueue {t_kernel_async:.3f}  full step {t_full:.3f}")
