"""Per-phase timing of the peer-sharded epoch (development tool): torchrun --nproc-per-node N tools/peer_phase_timing.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch as th, torch.distributed as dist
from svbrdf_diff_renderer_b200 import sharding, synth, _native as nv

res = int(os.environ.get("RES", "4096")); n = int(os.environ.get("LIGHTS", "64"))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local); dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cl = synth.calibration(n)
for pull in (True, False):
    ps = sharding.PeerShardedOptim(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev, pull=pull)
    with th.no_grad():
        tgt = ps.renderer.eval(synth.random_textures(res, 1).to(dev))
    ps.load_targets(tgt); del tgt
    ps.init_from_tex(synth.random_textures(res, 2))
    ps.optim(2, 0.01)
    L = nv.lib(); geom = ps.renderer._geom(ps.renderer._pow); stream = nv.stream_ptr(dev)
    if os.environ.get("PEER_ROTATE"):
        # timing experiment: pretend every tile is owned by ANOTHER rank (100 % remote traffic) by rotating the pointer tables
        rot = int(os.environ["PEER_ROTATE"])
        recv = [ps.peers.recv[i] for i in range(world)]; tex = [ps.peers.tex[i] for i in range(world)]
        for o in range(world):
            ps.peers.recv[o] = recv[(o + rot) % world]; ps.peers.tex[o] = tex[(o + rot) % world]
    m = th.zeros(9 * ps.chunk, device=dev); v = th.zeros_like(m); curve = th.zeros(8, device=dev)
    ev = [th.cuda.Event(enable_timing=True) for _ in range(5)]
    acc = [0.0] * 4
    reps = 5
    for e in range(reps):
        dist.barrier(); th.cuda.synchronize()
        ev[0].record()
        nv.check(L.svbrdf_l2_grad_push(ctypes.byref(geom), nv.ptr(ps.tex_sym), nv.ptr(ps.targets), 0, n, ctypes.byref(ps.peers),
                                       nv.ptr(curve), nv.ptr(ps.ws), stream), "k1")
        ev[1].record()
        ps.h_recv.barrier(channel=0)
        ev[2].record()
        a = nv.Adam(0.01, 0.9, 0.999, 1e-8, e + 3)
        nv.check(L.svbrdf_reduce_adam_push(ctypes.byref(ps.peers), ps.texels, nv.ptr(m), nv.ptr(v), ctypes.byref(a), stream), "k2")
        ev[3].record()
        ps.h_tex.barrier(channel=0)
        ev[4].record()
        th.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) / reps
    print(f"rank {rank} pull={pull} res={res} lights/gpu={ps.n_local}: grad+push {acc[0]:.3f} ms | barrier {acc[1]:.3f} | reduce+adam {acc[2]:.3f} | barrier {acc[3]:.3f}", flush=True)
    # reference: the same gradient kernel without any peer traffic (local l2_grad)
    grad = th.zeros(9, res, res, device=dev); loss = th.zeros(1, device=dev)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    dist.barrier(); th.cuda.synchronize(); e0.record()
    for _ in range(reps):
        nv.check(L.svbrdf_l2_grad(ctypes.byref(geom), nv.ptr(ps.tex_sym), nv.ptr(ps.targets), 0, n, nv.ptr(grad), nv.ptr(loss), None, nv.ptr(ps.ws), stream), "g")
    e1.record(); th.cuda.synchronize()
    if rank == 0:
        print(f"   local svbrdf_l2_grad (no peer traffic): {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
    del ps, grad
dist.destroy_process_group()
