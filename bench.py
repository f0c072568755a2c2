#!/usr/bin/env python
"""Benchmark of the per-pixel SVBRDF optimisation hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's torch-CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...          # material-sharded, one material per rank

A step is one pass of the hot path over one material: clamp -> render (N lights) -> L2 -> backward
-> Adam (the loop body of SvbrdfOptim.optim, /root/reference/src/svbrdf.py:60-71) at the
configuration BASELINE.json quotes the metric on: 1024 x 1024 texels, 9 co-located flash lights,
fp32 targets, synthetic random SVBRDF maps (`optim_perpixel` uses the L2 loss only).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: `roofline`, `cpu_baseline`,
`e2e` (per-step host->device upload of the step's targets + device->host loss read, through the
public SvbrdfOptim API), `e2e_job` (one optim() call of 20 epochs incl. uploads/downloads),
`clocks`, `gpu_launches`.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line (NCCL_DEBUG=VERSION prints there otherwise)

import torch as th  # noqa: E402

RES, LIGHTS = 1024, 9
LR = 0.01
JOB_EPOCHS = 20               # run.py:55-56: the refinement recipe runs 20 epochs per optim_perpixel call
N_MATERIALS_CYCLED = 4        # inputs cycled so that every step streams cold data (4 x 340 MB >> 126 MB L2)
METRIC = "pixel_light_samples_per_s"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--res", type=int, default=RES)
    ap.add_argument("--lights", type=int, default=LIGHTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-view-sharded", action="store_true", help="skip the view-sharded (strong-scaling) leg")
    ap.add_argument("--vs-res", type=int, default=4096)
    ap.add_argument("--vs-lights", type=int, default=256)
    ap.add_argument("--no-config5", action="store_true", help="skip the full-size view-sharded leg (8192^2 x 256, uint8 targets)")
    ap.add_argument("--no-config3", action="store_true", help="skip the 4096^2 x 64 three-roof leg (BASELINE configs[2]; N=1 only)")
    ap.add_argument("--no-material-batch", action="store_true", help="skip the 38-material batch (BASELINE configs[3])")
    ap.add_argument("--materials", type=int, default=38)
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager-on-this-GPU side line (N=1 only)")
    ap.add_argument("--no-descriptor", action="store_true", help="skip the L2 + descriptor-loss step (BASELINE configs[1] wording; N=1 only)")
    return ap.parse_args()


def config_dict(args, extra=None):
    cfg = {
        "workload": f"optim_perpixel fused step (clamp+render+L2+backward+Adam), {args.res}x{args.res} texels, "
                    f"{args.lights} co-located flash lights, fp32 targets (BASELINE.json configs[1])",
        "res": args.res, "lights": args.lights, "lr": LR, "target_dtype": "f32",
        "samples_per_step_per_gpu": args.res * args.res * args.lights,
        "l2": f"inputs larger than L2: every step streams {(216 + 12 * args.lights) * args.res * args.res / 1e6:.0f} MB "
              f"(targets + textures + Adam state) through a 126 MB L2; measured DRAM bytes equal the algorithmic bytes "
              f"(profiles/roofline_traffic.json); the single-epoch-launch figure cycles {N_MATERIALS_CYCLED} materials",
        "sharding": "material-sharded: one independent material per GPU, no data-path collective",
    }
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, interval=0.005):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        # seconds between NVML polls.  Dense (5 ms) while a region is timed on the DEVICE; the regions timed through the
        # host API (`e2e`, `e2e_job`) are sampled every 100 ms: every poll takes the driver's global lock and the GIL, which
        # the submission path of an e2e step (upload, launch, read-back, synchronise) needs too.  (The e2e step itself varies
        # between boxes of the pool: 0.61-0.67 ms where the pinned upload runs at 53 GB/s, 1.0-1.1 ms where it reaches 28 GB/s;
        # tools/e2e_breakdown.py splits a step into upload / kernel / read-back.)
        self.interval = interval
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "poll_s": {"device_timed_legs": 0.005, "host_timed_legs": self.interval}}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(res, lights, steps, warmup, budget_s=150.0):
    """Times the reference's torch-CPU path on all host cores.  With the reference's files staged (oracle/_ref, see
    oracle/stage_ref.py — or the live tree in the build container) the step is the UNMODIFIED ``Microfacet.eval``
    (src/microfacet.py:84-120) inside the loop body of ``SvbrdfOptim.optim`` (src/svbrdf.py:60-71: clamp -> eval -> MSELoss ->
    backward -> torch.optim.Adam.step), kind "reference"; otherwise the op-for-op port (oracle/torch_port.py), kind "port".
    Each step is a bounded sample of the workload: a top-left square crop of the image (the reference asserts square
    textures), all lights, sized after a probe so the whole run fits ``budget_s``; per-pixel cost is uniform."""
    from oracle import ref_loader
    from oracle import torch_port as tp
    from svbrdf_diff_renderer_b200 import synth

    cores = os.cpu_count() or 1
    try:                                                     # the GPU arm binds the process to its GPU's NUMA node: give the CPU baseline every core back
        os.sched_setaffinity(0, set(range(cores)))
    except Exception:
        pass
    th.set_num_threads(cores)
    cl = synth.calibration(lights)
    gt, t0 = synth.random_textures(res, 1), synth.random_textures(res, 2)
    use_ref = ref_loader.available()
    if use_ref:
        with ref_loader.quiet():
            RefMicrofacet, _, _ = ref_loader.load()

    def make(side):
        """State for a side x side crop: renderer/scene, targets, parameter, optimiser."""
        if use_ref:
            with ref_loader.quiet():
                ren = RefMicrofacet(res, lights, synth.IM_SIZE_CM, cl, "cpu")
            ren.res = side                                   # plain attributes (microfacet.py:12-24): crop the pixel grid
            for name in ("pos", "camera_pos", "light_pos", "light_pow"):
                setattr(ren, name, getattr(ren, name)[:, :, :side, :side])
            shade = ren.eval
        else:
            sc = tp.Scene(res, cl[0], cl[1], cl[2], synth.IM_SIZE_CM, th.float32, (0, side))
            sc.plane, sc.cam, sc.light, sc.power = (a[:, :, :, :side] for a in (sc.plane, sc.cam, sc.light, sc.power))
            shade = lambda t: tp.shade(sc, t)                # noqa: E731
        with th.no_grad():
            tgt = shade(gt[:, :, :side, :side].contiguous())
        tex = t0[:, :, :side, :side].clone().requires_grad_(True)
        opt = th.optim.Adam([tex], lr=LR, betas=(0.9, 0.999))
        return shade, tgt, tex, opt, th.nn.MSELoss()

    def step(shade, tgt, tex, opt, mse):
        loss = mse(shade(tex.clamp(-1, 1)), tgt)             # svbrdf.py:60-63
        opt.zero_grad()
        loss.backward()
        opt.step()                                           # svbrdf.py:69-71
        return loss

    probe = min(res, 256)
    state = make(probe)
    step(*state)
    t = time.perf_counter()
    step(*state)
    per_px = (time.perf_counter() - t) / (probe * probe)
    side = int(min(res, max(64, (budget_s / max(steps + warmup, 1) / per_px) ** 0.5)))
    state = make(side)
    for _ in range(warmup):
        step(*state)
    t = time.perf_counter()
    for _ in range(steps):
        step(*state)
    dt = time.perf_counter() - t
    samples = side * side * lights
    src = ("the unmodified reference Microfacet.eval (" + ("oracle/_ref staged copy" if ref_loader.staged() else "live reference tree") + ")") if use_ref \
        else "oracle/torch_port.py (op-for-op port)"
    return {"value": samples * steps / dt, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{steps} optim steps (after {warmup} warm-up) on a {side}x{side} crop of the {res}x{res}x{lights} workload, {src}, "
                      f"torch {th.__version__} CPU fp32, {cores} threads",
            "ms_per_step_sample": dt / steps * 1e3, "iters_per_s_full_image": (samples * steps / dt) / (res * res * lights)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_run(args.res, args.lights, args.steps, max(args.warmup, 1))
    v = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": args.res * args.res * args.lights / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(args),
        "reference_path": base["sample"],
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import ctypes

    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import _native as nv
    from svbrdf_diff_renderer_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not th.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a GPU (the CUDA path has no CPU fallback)")
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    from svbrdf_diff_renderer_b200 import sharding as _sh
    numa = _sh.bind_to_gpu_cpus(local)                     # pinned host buffers end up on the GPU's own NUMA node
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    res, n, K, W = args.res, args.lights, args.steps, max(args.warmup, 3)
    P = res * res
    L = nv.lib()

    # ---- synthetic inputs (CPU-seeded); every rank gets its own materials ----
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    mats = []
    for i in range(N_MATERIALS_CYCLED):
        seed = 1000 + 2 * (rank * N_MATERIALS_CYCLED + i)
        with th.no_grad():
            tgt = r.eval(synth.random_textures(res, seed).to(dev)).contiguous()
        tex0 = synth.random_textures(res, seed + 1).to(dev)
        mats.append({"target": tgt, "tex0": tex0, "tex": tex0[0].clone(), "m": th.zeros(9, res, res, device=dev),
                     "v": th.zeros(9, res, res, device=dev)})
    ws = r._workspace()
    geom = r._geom(r._pow)
    loss_dev = th.zeros(1, device=dev)
    stream = nv.stream_ptr(dev)
    step_no = [0]

    def fused_step(mat, with_loss=True):
        step_no[0] += 1
        a = nv.Adam(LR, 0.9, 0.999, 1e-8, step_no[0])
        nv.check(L.svbrdf_l2_adam_step(ctypes.byref(geom), nv.ptr(mat["tex"]), nv.ptr(mat["m"]), nv.ptr(mat["v"]), nv.ptr(mat["target"]), 0,
                                       ctypes.byref(a), nv.ptr(loss_dev) if with_loss else None, None, nv.ptr(ws), stream), "l2_adam_step")

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = th.tensor([ms], device=dev, dtype=th.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- value: device-resident inputs, K steps (= epochs) of the optimisation loop.  As SvbrdfOptim.optim does,
    # the epochs of one material are enqueued by ONE svbrdf_l2_adam_run call (up to 64 epochs per persistent launch);
    # the single-epoch-launch comparison below cycles 4 materials. ----
    curve = th.zeros(max(K, 64), device=dev)
    launches = [0]

    def run_epochs(mat, epochs, first_step):
        a = nv.Adam(LR, 0.9, 0.999, 1e-8, first_step)
        nv.check(L.svbrdf_l2_adam_run(ctypes.byref(geom), nv.ptr(mat["tex"]), nv.ptr(mat["m"]), nv.ptr(mat["v"]), nv.ptr(mat["target"]), 0,
                                      ctypes.byref(a), epochs, nv.ptr(curve), None, nv.ptr(ws), stream), "l2_adam_run")
        launches[0] += (epochs + 63) // 64

    def k_steps(_):
        # K epochs of ONE material, exactly what SvbrdfOptim.optim(K) enqueues: ceil(K/64) persistent launches.  Every
        # epoch streams the material's 340 MB (>> 126 MB L2) again; ncu's DRAM byte count equals the algorithmic bytes.
        run_epochs(mats[0], K, 1 + W)

    for i in range(W):
        fused_step(mats[i % N_MATERIALS_CYCLED])
    launches[0] = 0
    clk = ClockSampler(local)
    clk.__enter__()                                      # sampled over ALL timed regions below (the K-step region alone is a few ms)
    ms = timed(k_steps, 1)
    gpu_launches = launches[0]
    samples_per_step = P * n
    value = samples_per_step * world * K / (ms * 1e-3)
    ms_k = ms / K                                          # mean time of one epoch inside the persistent launches

    # ---- the same K steps as K separate single-epoch launches (svbrdf_l2_adam_step), for comparison ----
    ms_single = timed(lambda i: fused_step(mats[i % N_MATERIALS_CYCLED]), K) / K
    algo_bytes = (216 + 12 * n) * P
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get(f"l2_adam_{res}x{n}")
    except Exception:
        pass
    achieved = algo_bytes / (ms_k * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "svbrdf::tile_kernel<kModeL2Adam> (persistent TMA-pipelined fused render+L2+backward+Adam)", "kernel_ms_per_epoch": ms_k,
                "ms_per_epoch_single_epoch_launches": ms_single,
                "algorithmic_bytes_per_epoch": algo_bytes, "bytes_per_texel": 216 + 12 * n, "peak_source": peak_src,
                "samples_per_s_kernel_only": samples_per_step / (ms_k * 1e-3)}

    # ---- e2e: per step, upload that step's targets from pinned host memory, run, read the loss back ----
    e2e = e2e_job = None
    clk.interval = 0.1                                   # host-timed regions: keep the poller out of the submission path
    if not args.no_e2e:
        o = pkg.SvbrdfOptim(dev, r)
        host_targets = [m["target"].cpu().pin_memory() for m in mats[:2]]
        stage = th.empty_like(mats[0]["target"])
        o.init_from_tex(mats[0]["tex0"].clone())
        o.load_targets(stage)
        host_loss = th.empty(1, dtype=th.float32).pin_memory()
        e2e_steps = max(3, min(K, 20))
        em, ev = th.zeros(9, res, res, device=dev), th.zeros(9, res, res, device=dev)

        def e2e_step(i):
            stage.copy_(host_targets[i % 2], non_blocking=True)                     # H2D of this step's inputs
            a = nv.Adam(LR, 0.9, 0.999, 1e-8, i + 1)
            nv.check(L.svbrdf_l2_adam_step(ctypes.byref(geom), nv.ptr(o.textures.data), nv.ptr(em), nv.ptr(ev), nv.ptr(stage), 0,
                                           ctypes.byref(a), nv.ptr(loss_dev), None, nv.ptr(ws), stream), "l2_adam_step")
            host_loss.copy_(loss_dev, non_blocking=True)                            # D2H of the step's result
            th.cuda.current_stream().synchronize()

        for i in range(max(W, 3)):
            e2e_step(i)
        ms_e = timed(e2e_step, e2e_steps)
        e2e_f32 = {"value": samples_per_step * world * e2e_steps / (ms_e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": stage.numel() * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps, "ms_per_step": ms_e / e2e_steps,
                   "what": "same protocol with the targets uploaded as the reference's float32 stack (x/255 done on the host)"}

        # the same per-step protocol with the targets kept as what they are on disk — uint8 PNG bytes (imageio.py:18-19);
        # the fused kernel divides by 255 itself (bit-identical to the host division), a quarter of the upload
        host_u8 = [(t * 255).to(th.uint8).pin_memory() for t in host_targets]
        stage_u8 = th.empty(stage.shape, dtype=th.uint8, device=dev)

        def e2e_step_u8(i):
            stage_u8.copy_(host_u8[i % 2], non_blocking=True)
            a = nv.Adam(LR, 0.9, 0.999, 1e-8, i + 1)
            nv.check(L.svbrdf_l2_adam_step(ctypes.byref(geom), nv.ptr(o.textures.data), nv.ptr(em), nv.ptr(ev), nv.ptr(stage_u8), 1,
                                           ctypes.byref(a), nv.ptr(loss_dev), None, nv.ptr(ws), stream), "l2_adam_step")
            host_loss.copy_(loss_dev, non_blocking=True)
            th.cuda.current_stream().synchronize()

        for i in range(max(W, 3)):
            e2e_step_u8(i)
        ms_u8 = timed(e2e_step_u8, e2e_steps)
        e2e = {"value": samples_per_step * world * e2e_steps / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": stage_u8.numel(),
               "d2h_bytes_per_step": 4, "steps": e2e_steps, "ms_per_step": ms_u8 / e2e_steps,
               "what": "every step re-uploads its [N,3,R,R] targets from pinned host memory as what they are on disk and what optim_perpixel ingests "
                       "by default — uint8 PNG bytes (imageio.py:18-19), divided by 255 in-kernel, bit-identical to the host division — runs the fused "
                       "step through the C ABI and reads the loss back",
               "fp32_targets": e2e_f32}

        # the call a user makes: SvbrdfOptim.optim(20 epochs) on host-resident inputs, result back on the host
        host_tex0 = mats[0]["tex0"].cpu().pin_memory()
        host_out = th.empty_like(host_tex0).pin_memory()

        def make_job(bufs):
            def job(i):
                o.load_targets(bufs[i % 2].to(dev, non_blocking=True))
                o.init_from_tex(host_tex0.to(dev, non_blocking=True))
                losses = o.optim(JOB_EPOCHS, LR, None, False, progress=False)          # includes the loss-curve readback
                host_out.copy_(o.textures.detach(), non_blocking=True)
                th.cuda.current_stream().synchronize()
                return losses
            return job

        jobs = 4
        res_job = {}
        for key, bufs in (("u8", host_u8), ("f32", host_targets)):
            job = make_job(bufs)
            for i in range(2):                   # both host buffers once: the device blocks they alternate between are cached afterwards
                job(i)
            ms_j = timed(job, jobs)
            res_job[key] = {"value": samples_per_step * world * JOB_EPOCHS * jobs / (ms_j * 1e-3), "unit": UNIT, "epochs_per_call": JOB_EPOCHS,
                            "ms_per_call": ms_j / jobs, "h2d_bytes_per_call": bufs[0].numel() * bufs[0].element_size() + host_tex0.numel() * 4,
                            "d2h_bytes_per_call": host_out.numel() * 4 + 4 * JOB_EPOCHS}
        e2e_job = res_job["u8"]
        e2e_job["what"] = ("SvbrdfOptim.optim(20 epochs): pinned-host uint8 targets (optim_perpixel's default ingest) + init maps uploaded, 20 fused "
                           "epochs in one persistent launch, maps + loss curve downloaded")
        e2e_job["fp32_targets"] = res_job["f32"]

    clk.__exit__(None, None, None)

    # ---- view-sharded mode: ONE material, lights split over the ranks, NCCL all-reduce of the gradient ----
    view = config5 = None
    del mats
    th.cuda.empty_cache()
    if not args.no_view_sharded:
        view = view_sharded_bench(args, dev, world, rank, barrier, args.vs_res, args.vs_lights, th.float32,
                                  f"one material, {args.vs_res}^2 texels x {args.vs_lights} lights, fp32 targets (round 1's strong-scaling leg)")
    if not args.no_config5:
        config5 = view_sharded_bench(args, dev, world, rank, barrier, 8192, 256, th.uint8,
                                     "BASELINE.json configs[4] at full size: one material, 8192^2 texels x 256 light views, uint8 targets "
                                     "(51.5 GB in total: fits one B200)")

    # ---- BASELINE configs[3]: the material batch, round-robin over the ranks (every N) ----
    batch = None
    if not args.no_material_batch:
        batch = material_batch_bench(args, dev, world, rank, barrier)

    # ---- single-GPU extra legs: configs[2] three-roof line, L2 + descriptor step, torch-eager side line ----
    config3 = descriptor = eager = None
    if world == 1:
        if not args.no_config3:
            config3 = config3_bench(dev, local)
        if not args.no_descriptor:
            descriptor = descriptor_bench(dev)
        if not args.no_eager:
            eager = eager_bench(dev)

    # ---- cpu baseline (rank 0 only, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_reference_run(res, n, 3, 1, budget_s=25.0)
        cpu = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args), "iters_per_s_per_gpu": K / (ms * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_job": e2e_job, "view_sharded": view, "clocks": clk.summary(),
            "config5_view_sharded": config5, "config3": config3, "material_batch": batch, "config2_l2_plus_descriptor": descriptor, "reference_cuda_eager": eager,
            "gpu_launches": gpu_launches, "cpu_affinity": numa,
            "gpu_launches_what": "tile_kernel<L2Adam> launches in the timed region: one persistent launch per material and <=64 epochs (loss reduction fused: last CTA finalises each epoch)",
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def render_targets(res, cl_cpu, lights, rows, dev, dtype, gt, chunk=4):
    """Synthetic targets of light range `lights` = (start, end), rows `rows` = (r0, r1): rendered `chunk` lights at a time by
    the native forward kernel (a full [256,3,8192,8192] fp32 stack is 206 GB), quantised to the PNG bytes when `dtype` is
    uint8 (imageio.py:18-19 read side: x/255).  Untimed set-up."""
    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import synth
    s0, s1 = lights
    r0, r1 = rows
    out = th.empty(s1 - s0, 3, r1 - r0, res, dtype=dtype, device=dev)
    for a in range(s0, s1, chunk):
        b = min(a + chunk, s1)
        r = pkg.Microfacet(res, b - a, synth.IM_SIZE_CM, [cl_cpu[0][a:b].to(dev), cl_cpu[1][a:b].to(dev), cl_cpu[2].to(dev)], dev)
        with th.no_grad():
            img = r.eval(gt)[:, :, r0:r1, :]
            out[a - s0:b - s0] = (img * 255).round().to(th.uint8) if dtype == th.uint8 else img
        del img, r
    return out


def view_sharded_bench(args, dev, world, rank, barrier, res, n, dtype, label):
    """Strong scaling of ONE material: `res`^2 texels x `n` lights in total.  Transports timed on the
    same problem (same seeds at every N, so `loss_first_last` must agree across N):
      * N = 1: the fused single-GPU kernel (svbrdf_l2_adam_run: the denominator of the speed-up) and the unfused
        svbrdf_l2_grad + svbrdf_adam_apply pair the NCCL mode is built from;
      * "nccl": lights sharded; per row band svbrdf_l2_grad -> async NCCL all_reduce -> replicated svbrdf_adam_apply;
      * "peer_push": lights sharded; the reduce-scatter and the all-gather are fused into the kernels over NVLink peer
        memory (svbrdf_l2_grad_push / svbrdf_reduce_adam_push);
      * "hybrid_BxS": B row bands x S light shards — peer_push inside each band's group of S ranks, nothing between bands."""
    import ctypes

    import torch.distributed as dist
    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import _native as nv
    from svbrdf_diff_renderer_b200 import sharding, synth
    cl = synth.calibration(n)
    gt = synth.random_textures(res, 1).to(dev)
    tex0 = synth.random_textures(res, 2)
    epochs = 5
    tb = 1 if dtype == th.uint8 else 4
    out = {"workload": label, "unit": UNIT, "scaling": "strong", "res": res, "lights_total": n, "epochs_timed": epochs,
           "target_dtype": "u8" if dtype == th.uint8 else "f32", "target_bytes_total": n * 3 * res * res * tb,
           "gradient_bytes_per_epoch": 9 * res * res * 4}

    def reduce_max(ms):
        if world > 1:
            t = th.tensor([ms], device=dev, dtype=th.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def time_optim(opt):
        opt.optim(2, LR)                                 # warm-up (NCCL channels / symmetric-memory barriers, kernels)
        barrier()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        losses = opt.optim(epochs, LR)
        e1.record()
        barrier()
        ms = reduce_max(e0.elapsed_time(e1))
        return {"value": res * res * n * epochs / (ms * 1e-3), "ms_per_epoch": ms / epochs, "loss_first_last": [losses[0], losses[-1]]}

    start, end = sharding.split_range(n, world, rank)
    tgt = render_targets(res, cl, (start, end), (0, res), dev, dtype, gt)
    out["lights_per_gpu"] = end - start

    if world == 1:
        # the fused single-GPU kernel: what one B200 does on this problem (the speed-up's denominator)
        r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev)
        L = nv.lib()
        tex = tex0[0].to(dev).clone()
        m, v = th.zeros_like(tex), th.zeros_like(tex)
        curve = th.zeros(64, device=dev)
        geom, ws = r._geom(r._pow), r._workspace()

        def run(k, first):
            a = nv.Adam(LR, 0.9, 0.999, 1e-8, first)
            nv.check(L.svbrdf_l2_adam_run(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(tgt), nv.target_dtype_code(tgt), ctypes.byref(a), k,
                                          nv.ptr(curve), None, nv.ptr(ws), nv.stream_ptr(dev)), "l2_adam_run")
        run(2, 1)
        m.zero_()                                            # same protocol as the sharded optimisers below: every optim() call
        v.zero_()                                            # starts Adam afresh on the maps the warm-up left behind
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        run(epochs, 1)
        e1.record()
        th.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ls = curve[:epochs].tolist()
        out["fused_single_gpu"] = {"value": res * res * n * epochs / (ms * 1e-3), "ms_per_epoch": ms / epochs, "loss_first_last": [ls[0], ls[-1]],
                                   "what": "svbrdf_l2_adam_run: the fused render+L2+backward+Adam kernel, one launch per epoch at this size"}
        del tex, m, v, r

    vs = sharding.ViewShardedOptim(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev, bands=4)
    vs.load_targets(tgt)
    if world == 1:
        del tgt                                          # ViewShardedOptim keeps band-major copies
    vs.init_from_tex(tex0)
    out["nccl"] = time_optim(vs)
    out["nccl"]["what"] = ("4 row bands: svbrdf_l2_grad -> async NCCL all_reduce(SUM) of the band gradient (overlaps the next band) -> "
                           "replicated svbrdf_adam_apply") if world > 1 else "1 GPU: svbrdf_l2_grad + svbrdf_adam_apply per band, no collective"
    del vs
    th.cuda.empty_cache()
    if world > 1:
        ps = sharding.PeerShardedOptim(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev)
        ps.load_targets(tgt)
        ps.init_from_tex(tex0)
        out["peer_push"] = time_optim(ps)
        out["peer_push"]["pull_textures_from_owner"] = bool(ps.pull)
        out["peer_push"]["what"] = ("svbrdf_l2_grad_push (TMA-loads each tile's textures from the owner's replica over NVLink, stores the partial "
                                    "gradient into the owner's slot: all-gather and reduce-scatter fused into the gradient kernel) -> barrier -> "
                                    "svbrdf_reduce_adam_push (owner-side reduce + sharded Adam, local) -> barrier")
        del ps, tgt
        th.cuda.empty_cache()
        for shards in (2, 4):
            if shards >= world or world % shards:
                continue
            hy = sharding.HybridShardedOptim(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev, light_shards=shards)
            htgt = render_targets(res, cl, (hy.start, hy.end), hy.band, dev, dtype, gt)
            hy.load_targets(htgt)
            del htgt
            hy.init_from_tex(tex0)
            key = f"hybrid_{hy.n_bands}x{shards}"
            out[key] = time_optim(hy)
            out[key]["what"] = (f"{hy.n_bands} row bands x {shards} light shards: peer_push inside each band's group of {shards} ranks "
                                f"({(hy.end - hy.start)} lights x {hy.band[1] - hy.band[0]} rows per GPU), no traffic between bands")
            del hy
            th.cuda.empty_cache()
    keys = [k for k in out if isinstance(out[k], dict) and "ms_per_epoch" in out[k]]
    best = max(keys, key=lambda k: out[k]["value"])
    out["best"] = best
    out["value"] = out[best]["value"]
    out["ms_per_epoch"] = out[best]["ms_per_epoch"]
    del gt
    th.cuda.empty_cache()
    return out


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return None


def config3_bench(dev, local):
    """BASELINE configs[2]: the fused step at 4096^2 texels x 64 lights (1.07e9 samples, 16.5 GB streamed per epoch),
    reported against the three roofs of SURVEY.md section 8(d): HBM (984 B per texel), MUFU (16 lanes per SM and clock) and
    instruction issue (4 warp instructions per SM and clock).  Instructions and MUFU operations per sample are EXECUTED
    counts from the committed ncu capture of this kernel (profiles/r02_inst_mix_4096x64_v2.json, tools/ncu_inst_mix.py);
    the clock is the one NVML reports while the leg runs."""
    import ctypes

    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import _native as nv
    from svbrdf_diff_renderer_b200 import synth
    res, n, epochs = 4096, 64, 6
    P = res * res
    L = nv.lib()
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    with th.no_grad():
        tgt = r.eval(synth.random_textures(res, 1).to(dev)).contiguous()
    tex = synth.random_textures(res, 2)[0].to(dev)
    m, v = th.zeros_like(tex), th.zeros_like(tex)
    curve = th.zeros(64, device=dev)
    ws, geom, stream = r._workspace(), r._geom(r._pow), nv.stream_ptr(dev)

    def run(k, first):
        a = nv.Adam(LR, 0.9, 0.999, 1e-8, first)
        nv.check(L.svbrdf_l2_adam_run(ctypes.byref(geom), nv.ptr(tex), nv.ptr(m), nv.ptr(v), nv.ptr(tgt), 0, ctypes.byref(a), k, nv.ptr(curve), None,
                                      nv.ptr(ws), stream), "l2_adam_run")
    run(3, 1)                                                # warm-up epochs
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        run(epochs, 4)
        e1.record()
        th.cuda.synchronize()
    ms = e0.elapsed_time(e1) / epochs
    losses = curve[:epochs].tolist()
    clocks = clk.summary()
    mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965
    peaks = load_json("MEASURED_PEAKS.json") or {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    mix = load_json("profiles", "r02_inst_mix_4096x64_v2.json") or {}
    instr, mufu = mix.get("issue_slots_per_sample"), mix.get("mufu_per_sample")
    samples = P * n
    sps = samples / (ms * 1e-3)
    sms = 148
    out = {"workload": "fused step (clamp+render+L2+backward+Adam), 4096x4096 texels, 64 co-located lights, fp32 targets (BASELINE.json configs[2])",
           "epochs_timed": epochs, "ms_per_epoch": ms, "samples_per_s": sps, "loss_first_last": [losses[0], losses[-1]],
           "launches": epochs, "l2": "every epoch streams 16.5 GB through a 126 MB L2", "clocks": clocks,
           "hbm": {"bytes_per_texel": 216 + 12 * n, "achieved_gbs": (216 + 12 * n) * P / (ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback"}}
    out["hbm"]["frac"] = out["hbm"]["achieved_gbs"] / hbm_peak
    fracs = {"hbm": out["hbm"]["frac"]}
    if instr and mufu:
        mufu_peak = sms * 16 * mhz * 1e6                    # MUFU results per second
        issue_peak = sms * 128 * mhz * 1e6                  # thread instructions per second (4 schedulers x 32 lanes per SM)
        out["mufu"] = {"per_sample": mufu, "achieved_per_s": mufu * sps, "peak_per_s": mufu_peak, "frac": mufu * sps / mufu_peak}
        out["issue"] = {"thread_instructions_per_sample": instr, "achieved_per_s": instr * sps, "peak_per_s": issue_peak,
                        "frac": instr * sps / issue_peak}
        out["counts_source"] = "profiles/r02_inst_mix_4096x64_v2.json (ncu capture of this kernel at this size: smsp__inst_executed.sum x 32 / samples; MUFU from the SASS source counters)"
        out["sm_mhz_used"] = mhz
        fracs.update(mufu=out["mufu"]["frac"], issue=out["issue"]["frac"])
    out["binding_roof"] = max(fracs, key=fracs.get)
    out["fractions"] = fracs
    del tgt, tex, m, v
    th.cuda.empty_cache()
    return out


def eager_bench(dev):
    """The reference's torch path run EAGERLY on this GPU (what a user of the reference gets on the same B200 with
    `device=cuda:0`, scripts.py:68): the pinned op-for-op port (oracle/torch_port.py) — eval + MSELoss + backward +
    torch.optim.Adam.step per iteration, 745 aten calls.  Side line, not the --impl reference arm."""
    from oracle import torch_port as tp
    from svbrdf_diff_renderer_b200 import synth
    out = {}
    for res, iters in ((256, 20), (1024, 10)):
        n = 9
        cl = synth.calibration(n)
        sc = tp.Scene(res, cl[0], cl[1], cl[2], synth.IM_SIZE_CM, th.float32, device=dev)
        with th.no_grad():
            tgt = tp.shade(sc, synth.random_textures(res, 1).to(dev))
        tex = synth.random_textures(res, 2).to(dev).requires_grad_(True)
        opt = th.optim.Adam([tex], lr=LR, betas=(0.9, 0.999))

        def step():
            loss = tp.l2_loss(tp.shade(sc, tex.clamp(-1, 1)), tgt)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        for _ in range(3):
            step()
        th.cuda.synchronize()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            loss = step()
        e1.record()
        th.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[f"{res}x{res}x{n}"] = {"ms_per_iter": ms, "samples_per_s": res * res * n / (ms * 1e-3), "iters": iters, "loss": float(loss)}
        del sc, tgt, tex, opt
        th.cuda.empty_cache()
    out["what"] = "oracle/torch_port.py (bit-identical to the reference on CPU) on cuda:0, torch eager fp32, loss left on the device"
    return out


def descriptor_bench(dev):
    """BASELINE configs[1] as worded — "L2 + descriptor loss" — at 1024^2 x 9: one SvbrdfOptim.optim_with_features step
    (render + normalise + L2 in one native kernel each way, VGG19 feature network in torch/cuDNN, torch Adam), split into
    the native kernels' time and the rest by a second timing of the native part alone.  The pretrained VGG19 weights are a
    download (absent here): seeded random weights of the same architecture, same arithmetic cost."""
    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import synth
    from torchvision.models import vgg19
    res, n, steps = 1024, 9, 3
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    with th.no_grad():
        tgt = r.eval(synth.random_textures(res, 1).to(dev)).contiguous()
    th.manual_seed(7)
    vgg = pkg.VGGLoss(dev, net=vgg19(weights=None).features)
    vgg.load(tgt)
    o = pkg.SvbrdfOptim(dev, r)
    o.load_targets(tgt)
    o.init_from_tex(synth.random_textures(res, 2).to(dev))
    o.optim_with_features(1, LR, vgg, 0.1)                   # warm-up (cuDNN algorithm selection)
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    li, lf = o.optim_with_features(steps, LR, vgg, 0.1)
    e1.record()
    th.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / steps
    # the native share: the same forward + backward kernels with a stand-in upstream gradient instead of the network
    tex = o.textures.detach().clone().requires_grad_(True)
    up = None
    for it in range(steps + 1):                              # first pass untimed: the 113 MB image buffers come from cudaMalloc once
        if it == 1:
            th.cuda.synchronize()
            e0.record()
        norm, l2 = r.eval_normalized(tex.clamp(-1, 1), vgg.mean, vgg.std, tgt)
        up = th.ones_like(norm) if up is None else up
        th.autograd.grad([norm, l2], [tex], [up, th.ones_like(l2)])
        del norm, l2
    e1.record()
    th.cuda.synchronize()
    ms_native = e0.elapsed_time(e1) / steps
    out = {"workload": "optim_perpixel step with L2 + 0.1 x descriptor (VGG19) loss, 1024x1024 x 9 lights", "steps": steps, "ms_per_step": ms_step,
           "ms_native_render_fwd_bwd": ms_native, "ms_vgg_and_torch": ms_step - ms_native, "loss_image": li[-1], "loss_feature": lf[-1],
           "tf32": bool(th.backends.cudnn.allow_tf32), "weights": "seeded random VGG19 (the pretrained file is a download)",
           "what": "ms_native = svbrdf_render_norm_l2_fwd + _bwd (+ clamp and a ones_like per step); the remainder is the cuDNN VGG19 forward/backward, "
                   "torch.optim.Adam and autograd bookkeeping"}
    del vgg, o, tgt
    th.cuda.empty_cache()
    return out


def material_batch_bench(args, dev, world, rank, barrier):
    """BASELINE configs[3]: a batch of independent materials at 1024^2 x 9, 20 epochs each, dealt round-robin over the ranks
    (38 on 8 GPUs -> 5,5,5,5,5,5,4,4: ideal speed-up 7.6x) through sharding.optimise_materials — uploads of every material's
    targets and start maps from pinned host memory and the download of its result included.  The targets travel as what
    they are on disk, uint8 PNG bytes (imageio.py:18-19; decoded in-kernel, bit-identical to the host's /255); the fp32
    upload is timed beside it."""
    import torch.distributed as dist
    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import sharding, synth
    res, n, M = 1024, 9, args.materials
    cl = [c.to(dev) for c in synth.calibration(n)]
    r = pkg.Microfacet(res, n, synth.IM_SIZE_CM, cl, dev)
    mine = sharding.round_robin(M, world, rank)
    host = {}
    for i in mine:                                           # untimed: synthesise this rank's inputs, park them in pinned host memory
        with th.no_grad():
            t = r.eval(synth.random_textures(res, 1000 + 2 * i).to(dev))
        u8 = (t * 255).round().to(th.uint8)
        host[i] = {"u8": u8.cpu().pin_memory(), "f32": (u8.float() / 255).cpu().pin_memory(),
                   "tex0": synth.random_textures(res, 1001 + 2 * i).pin_memory()}
        del t, u8
    out = {"materials": M, "res": res, "lights": n, "epochs_per_material": JOB_EPOCHS, "materials_per_rank": [len(sharding.round_robin(M, world, k)) for k in range(world)],
           "ideal_speedup_vs_1_gpu": M / max(len(sharding.round_robin(M, world, k)) for k in range(world))}
    results = {}
    for key in ("warm", "u8", "f32"):
        def make_problem(i, key=key):
            return r, host[i]["u8" if key == "warm" else key].to(dev, non_blocking=True), host[i]["tex0"].to(dev, non_blocking=True)
        if key == "warm":
            # one untimed pass: device blocks and the pinned result buffers (cudaHostAlloc: ~1 ms per 40 MB map set) enter
            # the caching allocators, as they would for the second batch of a long-running job
            warm = sharding.optimise_materials(M, make_problem, 2, LR, dev, to_host=True)
            del warm
            continue
        barrier()
        t0 = time.perf_counter()
        mine_out, finals = sharding.optimise_materials(M, make_problem, JOB_EPOCHS, LR, dev, to_host=True)
        th.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = th.tensor([dt], device=dev, dtype=th.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        bytes_up = sum(host[i][key].numel() * host[i][key].element_size() + host[i]["tex0"].numel() * 4 for i in mine)
        results[key] = {"wall_s": dt, "materials_per_s": M / dt, "samples_per_s": M * JOB_EPOCHS * res * res * n / dt,
                        "h2d_bytes_this_rank": bytes_up, "final_loss_mean": float(sum(finals) / len(finals))}
        del mine_out
    out["uint8_targets"], out["fp32_targets"] = results["u8"], results["f32"]
    out["what"] = ("wall clock (max over ranks) of sharding.optimise_materials: per material upload targets + start maps from pinned host memory "
                   "(prefetched on a copy stream while the previous material runs), 20 fused epochs in one persistent launch, maps back to the host")
    del host
    th.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    # stdout carries exactly one JSON line: everything else any library writes to fd 1 during the run (NCCL prints
    # "NCCL version ..." there on communicator creation, whatever NCCL_DEBUG_FILE says) is sent to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
