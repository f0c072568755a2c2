"""Per-phase timing of the peer-sharded epoch (development tool):
    RES=4096 LIGHTS=256 SHARDS=2 torchrun --nproc-per-node 8 tools/peer_phase_timing.py
LIGHTS = lights in total; SHARDS = light shards per row band (default: world = the 1-D view-sharded mode; 2 on 8 ranks =
4 row bands x 2 light shards).  Prints, per rank, the four phases of an epoch and the same gradient kernel without peers."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SVBRDF_B200_QUIET", "1")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch as th, torch.distributed as dist
from svbrdf_diff_renderer_b200 import sharding, synth, _native as nv
from bench import render_targets

res = int(os.environ.get("RES", "4096")); n = int(os.environ.get("LIGHTS", "256"))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
shards = int(os.environ.get("SHARDS", str(world)))
dtype = th.uint8 if os.environ.get("U8") else th.float32
th.cuda.set_device(local); dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cl = synth.calibration(n)
if shards == world:
    ps = sharding.PeerShardedOptim(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev)
else:
    ps = sharding.HybridShardedOptim(res, n, synth.IM_SIZE_CM, [c.to(dev) for c in cl], dev, light_shards=shards).inner
tgt = render_targets(res, cl, (ps.start, ps.end), ps.band, dev, dtype, synth.random_textures(res, 1).to(dev))
ps.load_targets(tgt); del tgt
ps.init_from_tex(synth.random_textures(res, 2))
ps.optim(2, 0.01)
L = nv.lib(); stream = nv.stream_ptr(dev)
geom = ps.renderer._geom(ps.renderer._pow, rows=ps.rows, row_offset=ps.band[0])
code = nv.target_dtype_code(ps.targets)
m = th.zeros(9 * ps.chunk, device=dev); v = th.zeros_like(m); curve = th.zeros(8, device=dev)
ev = [th.cuda.Event(enable_timing=True) for _ in range(5)]
acc = [0.0] * 4
reps = 5
for e in range(reps):
    dist.barrier(); th.cuda.synchronize()
    ev[0].record()
    nv.check(L.svbrdf_l2_grad_push(ctypes.byref(geom), nv.ptr(ps.tex_sym), nv.ptr(ps.targets), code, n, ctypes.byref(ps.peers),
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
print(f"rank {rank} bands x shards = {world // shards} x {shards}, {res}^2 x {n} lights ({ps.rows} rows x {ps.n_local} lights per GPU): "
      f"grad+push {acc[0]:.3f} ms | barrier {acc[1]:.3f} | reduce+adam {acc[2]:.3f} | barrier {acc[3]:.3f} | sum {sum(acc):.3f}", flush=True)
# reference: the same gradient kernel on the same band and lights without any peer traffic (local svbrdf_l2_grad)
grad = th.zeros(9, ps.rows, res, device=dev); loss = th.zeros(1, device=dev)
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
dist.barrier(); th.cuda.synchronize(); e0.record()
for _ in range(reps):
    nv.check(L.svbrdf_l2_grad(ctypes.byref(geom), nv.ptr(ps.tex_sym), nv.ptr(ps.targets), code, n, nv.ptr(grad), nv.ptr(loss), None, nv.ptr(ps.ws), stream), "g")
e1.record(); th.cuda.synchronize()
if rank == 0:
    print(f"   local svbrdf_l2_grad on the same band and lights (no peer traffic): {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
dist.destroy_process_group()
