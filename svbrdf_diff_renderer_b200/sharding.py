"""Multi-GPU partitioning of the per-pixel optimisation path (one process per GPU, torch.distributed).

The reference is single-device (``/root/reference/src/scripts.py:68``); both shardings are new
(SURVEY.md §8(e)):

* **material-sharded** — independent materials are dealt round-robin to the ranks; every rank runs
  the single-GPU fused loop on its own materials; there is NO data-path collective (only the final
  losses are gathered on the host).

* **view-sharded** — one material whose light/view stack is split into contiguous shards.  Textures
  and Adam state are replicated.  Because ``dL/dtex = sum over lights`` and the MSE mean divides by
  the GLOBAL count ``N_total*3*R*R``, every rank calls ``svbrdf_l2_grad`` with its own lights and
  ``n_total``; the partial gradients add.  Per epoch, for each row band:
  ``svbrdf_l2_grad(band)`` -> ``all_reduce(SUM)`` of the band's ``[9,rows,R]`` gradient (NCCL over
  NVLink/NVSwitch, asynchronous: it overlaps the next band's compute) -> ``svbrdf_adam_apply(band)``.
  The 4-float ``[loss, dpow]`` vector is reduced once per epoch.  All ranks apply the identical Adam
  update, so the replicas stay bit-identical without a broadcast.

The arithmetic engine is injectable so the partitioning logic can be exercised with ``gloo`` on CPU
by the tests (tests/test_sharding_gloo.py plugs in the oracle there); the product engine is
``NativeEngine`` — the CUDA kernels behind the C ABI, nothing else.
"""

from __future__ import annotations

import ctypes

import torch as th
import torch.distributed as dist


def split_range(n: int, world: int, rank: int):
    """Contiguous, balanced split of ``range(n)``: the first ``n % world`` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def round_robin(n: int, world: int, rank: int):
    """Material indices of ``rank``: 38 materials on 8 GPUs -> 5,5,5,5,5,5,4,4."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n, world))


def bind_to_gpu_cpus(device_index: int):
    """One process per GPU: restrict this process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA
    node), so the pinned staging buffers it allocates afterwards — first touch by the allocating thread — live in the host
    memory that GPU's PCIe root complex reaches without crossing the inter-socket link.  Returns the core list (``None`` when
    NVML or the affinity call is unavailable: nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"cores": len(allowed), "first": allowed[0], "last": allowed[-1]}
    except Exception:
        return None


def row_bands(res: int, bands: int):
    """Split the rows of a ``res x res`` image into ``bands`` contiguous bands (multiples of 8 rows when possible)."""
    bands = max(1, min(bands, res))
    edges = [((res * b // bands) // 8) * 8 if res >= 8 * bands else res * b // bands for b in range(bands)] + [res]
    return [(edges[b], edges[b + 1]) for b in range(bands) if edges[b + 1] > edges[b]]


class NativeEngine:
    """The CUDA kernels behind ``include/svbrdf_b200.h`` for one rank's light shard."""

    def __init__(self, renderer):
        from . import _native as nv
        self.nv, self.r = nv, renderer
        self.L = nv.lib()
        self._ws = {}

    def _workspace(self, rows):
        if rows not in self._ws:
            self._ws[rows] = self.nv.workspace(self.r.res, rows, self.r.device)
        return self._ws[rows]

    def l2_grad_band(self, tex, targets, n_total, band, grad, loss_out):
        """One row band, all tensors band-contiguous: tex/grad ``[9,rows,R]``, targets ``[n_local,3,rows,R]``.
        grad <- this shard's share of dL/dtex for the band; loss_out[0] <- its share of the loss."""
        nv, r = self.nv, self.r
        r0, r1 = band
        geom = r._geom(r._pow, rows=r1 - r0, row_offset=r0)
        code = self.L.svbrdf_l2_grad(ctypes.byref(geom), nv.ptr(tex), nv.ptr(targets), nv.target_dtype_code(targets), n_total,
                                     nv.ptr(grad), nv.ptr(loss_out), None, nv.ptr(self._workspace(r1 - r0)), nv.stream_ptr(r.device))
        nv.check(code, "svbrdf_l2_grad")

    def adam_apply(self, p, m, v, g, step, lr):
        nv = self.nv
        a = nv.Adam(float(lr), 0.9, 0.999, 1e-8, int(step))
        nv.check(self.L.svbrdf_adam_apply(nv.ptr(p), nv.ptr(m), nv.ptr(v), nv.ptr(g), p.numel(), ctypes.byref(a), nv.stream_ptr(p.device)),
                 "svbrdf_adam_apply")


class ViewShardedOptim:
    """One material, lights split across the ranks of ``group`` (default: the world group).

    ``cl`` is the FULL calibration ``[camera_pos[N,3], light_pos[N,3], light_pow[3]]``; each rank keeps
    lights ``split_range(N, world, rank)`` and loads only those targets.

    Storage is band-major: textures, Adam state, gradient and targets are kept as one contiguous block per
    row band, so a band is a self-contained launch (``rows``/``row_offset`` of the ABI) and its gradient is
    one contiguous NCCL buffer.
    """

    def __init__(self, res, n_total, size, cl, device, group=None, engine_factory=None, bands=4):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.res, self.n_total, self.size, self.device = res, n_total, float(size), th.device(device)
        self.start, self.end = split_range(n_total, self.world, self.rank)
        self.n_local = self.end - self.start
        if self.n_local < 1:
            raise RuntimeError(f"rank {self.rank}: no lights to own ({n_total} lights on {self.world} ranks)")
        self.local_cl = [cl[0][self.start:self.end].contiguous(), cl[1][self.start:self.end].contiguous(), cl[2]]
        if engine_factory is None:
            from .microfacet import Microfacet
            self.renderer = Microfacet(res, self.n_local, size, [c.to(self.device) for c in self.local_cl], self.device)
            self.engine = NativeEngine(self.renderer)
        else:
            self.renderer = None
            self.engine = engine_factory(self)
        self.bands = row_bands(res, bands)
        self.losses = []

    def load_targets(self, local_targets):
        """``[n_local,3,R,R]`` — this rank's shard of the target stack (float32 or uint8); re-laid out per band once."""
        if tuple(local_targets.shape) != (self.n_local, 3, self.res, self.res):
            raise RuntimeError(f"rank {self.rank}: targets must be [{self.n_local},3,{self.res},{self.res}]")
        self.targets = [local_targets[:, :, r0:r1, :].to(self.device).contiguous() for r0, r1 in self.bands]

    def init_from_tex(self, textures):
        """Replicated start maps ``[1,9,R,R]``; rank 0's copy wins so replicas start bit-identical."""
        full = textures.detach().to(device=self.device, dtype=th.float32).contiguous().clone()
        if self.world > 1:
            dist.broadcast(full, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        self.tex = [full[0, :, r0:r1, :].contiguous() for r0, r1 in self.bands]

    @property
    def textures(self):
        """The current maps as one ``[1,9,R,R]`` tensor (assembled from the bands)."""
        return th.cat(self.tex, 1).unsqueeze(0)

    def optim(self, epochs, lr):
        nb = len(self.bands)
        m = [th.zeros_like(t) for t in self.tex]
        v = [th.zeros_like(t) for t in self.tex]
        grad = [th.zeros_like(t) for t in self.tex]
        scal = th.zeros(nb, dtype=th.float32, device=self.device)
        curve = th.zeros(max(epochs, 1), dtype=th.float32, device=self.device)
        for epoch in range(epochs):
            works = []
            for b, band in enumerate(self.bands):
                self.engine.l2_grad_band(self.tex[b], self.targets[b], self.n_total, band, grad[b], scal[b:b + 1])
                if self.world > 1:       # asynchronous: overlaps the next band's gradient kernel
                    works.append(dist.all_reduce(grad[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for b in range(nb):
                if works:
                    works[b].wait()
                # identical update on every rank: the replicas stay bit-identical without a broadcast
                self.engine.adam_apply(self.tex[b], m[b], v[b], grad[b], epoch + 1, lr)
            curve[epoch] = scal.sum()
        if self.world > 1 and epochs > 0:
            dist.all_reduce(curve, op=dist.ReduceOp.SUM, group=self.group)      # loss shares add; once per run
        self.losses = curve[:epochs].tolist()
        return self.losses


class PeerShardedOptim:
    """View-sharded optimisation with the collective fused into the kernels over NVLink peer memory.

    Same partitioning as ``ViewShardedOptim`` (contiguous light shards, replicated textures, global MSE normaliser)
    but no NCCL call on the data path.  Per epoch:

    1. ``svbrdf_l2_grad_push`` — the persistent gradient kernel stores each tile's partial gradient straight into
       the receive slot of the rank that OWNS the tile (texel ``p`` belongs to rank ``p // chunk``), i.e. the
       reduce-scatter traffic leaves over NVLink while the next tiles are being shaded;
    2. a cross-rank barrier (symmetric-memory signal pads, on the stream);
    3. ``svbrdf_reduce_adam_push`` — every rank sums the ``world`` partials of ITS texels in rank order, applies Adam
       to them (``m``/``v`` exist only for owned texels: 1/world of the optimiser state and work) and stores the new
       parameters into every rank's replica (all-gather by peer stores);
    4. a second barrier.

    Every texel is reduced and updated on exactly one rank, so the replicas are bit-identical by construction.
    Buffers come from ``torch.distributed._symmetric_memory`` (CUDA VMM handles exchanged at rendezvous); PyTorch
    only allocates and exchanges pointers — all arithmetic and all data movement is in the two kernels.
    """

    TILE = 480      # ownership granularity = tile size of the gradient kernel (include/svbrdf_b200.h)

    def __init__(self, res, n_total, size, cl, device, group=None, multicast=True, pull=True, band=None, loss_group=None):
        """``band=(r0, r1)``: the peer group shares only rows [r0, r1) of the image (targets ``[n_local,3,r1-r0,R]``, maps
        ``[1,9,r1-r0,R]``) — the building block of the 2-D decomposition (``HybridShardedOptim``).  ``loss_group``: the group
        the per-epoch loss shares are summed over (default: ``group``; the hybrid mode passes the world group)."""
        import torch.distributed._symmetric_memory as symm
        from . import _native as nv
        from .microfacet import Microfacet
        if not dist.is_initialized():
            raise RuntimeError("PeerShardedOptim needs an initialised process group (NCCL, one process per GPU)")
        self.nv, self.symm = nv, symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("peer-push mode covers one NVSwitch box: at most 8 ranks")
        self.res, self.n_total, self.size, self.device = res, n_total, float(size), th.device(device)
        self.start, self.end = split_range(n_total, self.world, self.rank)
        self.n_local = self.end - self.start
        if self.n_local < 1:
            raise RuntimeError(f"rank {self.rank}: no lights to own")
        self.renderer = Microfacet(res, self.n_local, size,
                                   [cl[0][self.start:self.end].to(self.device), cl[1][self.start:self.end].to(self.device), cl[2].to(self.device)],
                                   self.device)
        self.band = (0, res) if band is None else (int(band[0]), int(band[1]))
        if not (0 <= self.band[0] < self.band[1] <= res):
            raise ValueError(f"bad row band {band} for a {res}-row image")
        self.rows = self.band[1] - self.band[0]
        self.loss_group = self.group if loss_group is None else loss_group
        self.texels = self.rows * res
        if self.texels % 4:
            raise RuntimeError("peer-push mode needs the band's texel count to be a multiple of 4")
        per = -(-self.texels // self.world)
        self.chunk = -(-per // self.TILE) * self.TILE
        name = self.group.group_name
        self.tex_sym = symm.empty(9 * self.texels, dtype=th.float32, device=self.device)
        self.recv_sym = symm.empty(self.world * 9 * self.chunk, dtype=th.float32, device=self.device)
        self.h_tex = symm.rendezvous(self.tex_sym, name)
        self.h_recv = symm.rendezvous(self.recv_sym, name)
        self.peers = nv.Peers(self.world, self.rank, self.chunk)
        for r in range(self.world):
            self.peers.recv[r] = int(self.h_recv.buffer_ptrs[r])
            self.peers.tex[r] = int(self.h_tex.buffer_ptrs[r])
        # NVSwitch multicast (NVLS): one multimem.st replaces `world` peer stores in the all-gather
        mc = 0
        try:
            if multicast and self.h_tex.has_multicast_support(self.device.type, self.device.index):
                mc = int(self.h_tex.multicast_ptr or 0)
        except Exception:
            mc = 0
        self.multicast = mc != 0
        self.peers.tex_multicast = mc if mc else None
        # pull mode (default): no all-gather — the gradient kernel reads each tile's textures from its owner over NVLink
        self.pull = bool(pull)
        self.peers.pull_tex = 1 if self.pull else 0
        self.ws = nv.workspace(res, self.rows, self.device)
        self.losses = []

    def load_targets(self, local_targets):
        """``[n_local,3,rows,R]``: this rank's light shard of the (band of the) target stack, float32 or uint8."""
        if tuple(local_targets.shape) != (self.n_local, 3, self.rows, self.res):
            raise RuntimeError(f"rank {self.rank}: targets must be [{self.n_local},3,{self.rows},{self.res}]")
        self.targets = local_targets.to(self.device).contiguous()

    def init_from_tex(self, textures):
        """Start maps ``[1,9,R,R]`` (the band is cut out here) or ``[1,9,rows,R]``; the group's rank 0 wins."""
        full = textures.detach().to(device=self.device, dtype=th.float32)
        if full.shape[2] == self.res and self.rows != self.res:
            full = full[:, :, self.band[0]:self.band[1], :]
        full = full.contiguous().clone()
        dist.broadcast(full, src=dist.get_global_rank(self.group, 0), group=self.group)
        self.tex_sym.copy_(full.view(-1))
        th.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    @property
    def textures(self):
        """Current maps ``[1,9,R,R]``.  In pull mode a replica is authoritative only for the texels its rank owns, so
        the owned chunks are gathered (one collective, outside the optimisation loop)."""
        full = self.tex_sym.view(9, self.texels)
        if not self.pull:
            return full.view(1, 9, self.rows, self.res)
        mine = th.zeros(9, self.chunk, dtype=th.float32, device=self.device)
        lo = self.rank * self.chunk
        hi = min(lo + self.chunk, self.texels)
        if hi > lo:
            mine[:, :hi - lo] = full[:, lo:hi]
        parts = [th.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return th.cat(parts, 1)[:, :self.texels].reshape(1, 9, self.rows, self.res).contiguous()

    def optim(self, epochs, lr):
        nv, L = self.nv, self.nv.lib()
        m = th.zeros(9 * self.chunk, dtype=th.float32, device=self.device)
        v = th.zeros_like(m)
        curve = th.zeros(max(epochs, 1), dtype=th.float32, device=self.device)
        geom = self.renderer._geom(self.renderer._pow, rows=self.rows, row_offset=self.band[0])
        stream = nv.stream_ptr(self.device)
        for epoch in range(epochs):
            nv.check(L.svbrdf_l2_grad_push(ctypes.byref(geom), nv.ptr(self.tex_sym), nv.ptr(self.targets), nv.target_dtype_code(self.targets),
                                           self.n_total, ctypes.byref(self.peers), ctypes.c_void_p(curve.data_ptr() + 4 * epoch), nv.ptr(self.ws),
                                           stream), "svbrdf_l2_grad_push")
            self.h_recv.barrier(channel=0)                  # every rank's partials have landed
            a = nv.Adam(float(lr), 0.9, 0.999, 1e-8, epoch + 1)
            nv.check(L.svbrdf_reduce_adam_push(ctypes.byref(self.peers), self.texels, nv.ptr(m), nv.ptr(v), ctypes.byref(a), stream),
                     "svbrdf_reduce_adam_push")
            self.h_tex.barrier(channel=0)                   # every replica holds the new parameters
        if epochs > 0:
            dist.all_reduce(curve, op=dist.ReduceOp.SUM, group=self.loss_group)      # loss shares add; once per run
        self.losses = curve[:epochs].tolist()
        return self.losses


def hybrid_layout(world: int, light_shards: int, rank: int):
    """2-D decomposition of one material over ``world`` ranks: ``bands = world // light_shards`` row bands x ``light_shards``
    light shards.  Rank r works on band ``r // light_shards`` with light shard ``r % light_shards``; the ranks of one band
    are consecutive (one peer group per band).  Returns ``(band index, shard index, [[ranks of band 0], [ranks of band 1], ...])``."""
    if light_shards < 1 or world % light_shards:
        raise ValueError(f"{world} ranks cannot be split into groups of {light_shards} light shards")
    if not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    bands = world // light_shards
    groups = [list(range(b * light_shards, (b + 1) * light_shards)) for b in range(bands)]
    return rank // light_shards, rank % light_shards, groups


class HybridShardedOptim:
    """One material on ``world = bands x light_shards`` GPUs: the image is cut into ``bands`` row bands, every band is
    view-sharded over a peer group of ``light_shards`` ranks (``PeerShardedOptim`` on the band: gradient exchange fused
    into the kernels over NVLink), and bands never talk to each other (texels are independent; only the loss is summed).

    Why: in the 1-D view-sharded mode every rank runs the per-texel prologue/epilogue, pulls 9 texture planes and pushes a
    36-byte partial gradient for ALL texels, amortised over only N/world lights — at 8 ranks that fixed work is what
    bounds the scaling (profiles/r01_peer_phase_timing_8gpu.txt).  With b bands it shrinks by b, and the exchange stays
    inside groups of world/b ranks.  ``light_shards = world`` is the 1-D mode; ``light_shards = 1`` would be pure pixel-band
    sharding (no exchange at all) and is refused here — that is not a view-sharded run."""

    def __init__(self, res, n_total, size, cl, device, light_shards, multicast=True, pull=True):
        if not dist.is_initialized():
            raise RuntimeError("HybridShardedOptim needs an initialised process group (NCCL, one process per GPU)")
        world, rank = dist.get_world_size(), dist.get_rank()
        if light_shards < 2:
            raise ValueError("light_shards must be >= 2 (1 = pixel-band sharding, not a view-sharded run)")
        self.band_index, self.shard_index, groups = hybrid_layout(world, light_shards, rank)
        self.n_bands = len(groups)
        pgs = [dist.new_group(ranks=g) for g in groups]       # every rank creates every group (collective)
        self.group = pgs[self.band_index]
        self.band = row_bands(res, self.n_bands)[self.band_index] if self.n_bands > 1 else (0, res)
        if len(row_bands(res, self.n_bands)) != self.n_bands:
            raise RuntimeError(f"a {res}-row image cannot be cut into {self.n_bands} bands")
        self.inner = PeerShardedOptim(res, n_total, size, cl, device, group=self.group, multicast=multicast, pull=pull, band=self.band,
                                      loss_group=dist.group.WORLD)
        self.res, self.n_total, self.device = res, n_total, th.device(device)
        self.start, self.end, self.n_local = self.inner.start, self.inner.end, self.inner.n_local
        self.renderer = self.inner.renderer

    def load_targets(self, local_targets):
        """``[n_local,3,rows,R]``: lights ``start:end`` of rows ``band[0]:band[1]``."""
        self.inner.load_targets(local_targets)

    def init_from_tex(self, textures):
        full = textures.detach().to(device=self.device, dtype=th.float32).contiguous().clone()
        dist.broadcast(full, src=0)                            # world rank 0 wins, then each band group takes its rows
        self.inner.init_from_tex(full)

    @property
    def textures(self):
        """``[1,9,R,R]`` assembled from the bands (one all-gather over the world, outside the optimisation loop)."""
        mine = self.inner.textures                              # [1,9,rows,R], identical within the band group
        world = dist.get_world_size()
        bands = row_bands(self.res, self.n_bands)
        rows_max = max(b[1] - b[0] for b in bands)
        pad = th.zeros(1, 9, rows_max, self.res, dtype=th.float32, device=self.device)
        pad[:, :, :mine.shape[2]] = mine
        parts = [th.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        shards = world // self.n_bands
        return th.cat([parts[b * shards][:, :, :bands[b][1] - bands[b][0]] for b in range(self.n_bands)], 2).contiguous()

    def optim(self, epochs, lr):
        self.losses = self.inner.optim(epochs, lr)
        return self.losses


def optimise_materials(n_materials, make_problem, epochs, lr, device, group=None, to_host=False, prefetch=True):
    """Material-sharded driver: ``make_problem(i)`` -> ``(renderer, targets, start_textures)`` for material i.

    Returns ``{material index: (final loss, optimised textures)}`` for this rank's materials and the list of
    final losses of ALL materials (gathered on the host — the only communication of this mode).

    ``prefetch``: material i+1's ``make_problem`` (typically host->device uploads of its targets and start maps) runs on a
    copy stream while material i is being optimised, so the PCIe transfer hides behind the kernel (or the other way round).
    ``to_host``: the optimised maps are copied to pinned host memory (asynchronously) instead of being kept on the device.
    """
    from .svbrdf import SvbrdfOptim
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    device = th.device(device)
    ids = round_robin(n_materials, world, rank)
    main = th.cuda.current_stream(device)
    copy = th.cuda.Stream(device) if prefetch else main

    def fetch(i):
        """make_problem under the copy stream; the tensors are handed to the main stream by an event."""
        with th.cuda.stream(copy):
            renderer, targets, tex0 = make_problem(i)
            ev = copy.record_event()
        for t in (targets, tex0):
            if prefetch and isinstance(t, th.Tensor) and t.is_cuda:
                t.record_stream(main)                          # allocated on the copy stream, consumed on the main one
        return renderer, targets, tex0, ev

    mine = {}
    nxt = fetch(ids[0]) if ids else None
    for k, i in enumerate(ids):
        renderer, targets, tex0, ev = nxt
        nxt = fetch(ids[k + 1]) if k + 1 < len(ids) else None   # enqueued before this material's kernels: overlaps them
        main.wait_event(ev)
        opt = SvbrdfOptim(device, renderer)
        opt.load_targets(targets)
        opt.init_from_tex(tex0)
        curve = opt.optim(epochs, lr, None, False, progress=False, read_back=False)
        maps = opt.textures.detach()
        if to_host:
            pinned = th.empty(maps.shape, dtype=maps.dtype, pin_memory=True)
            pinned.copy_(maps, non_blocking=True)
            maps = pinned
        mine[i] = (curve, maps)
    # one synchronisation for the whole batch: the loss curves are read after everything has been enqueued
    def last(curve):
        if isinstance(curve, th.Tensor):
            return float(curve[-1]) if curve.numel() else float("nan")
        return curve[-1] if curve else float("nan")
    mine = {i: (last(c), m) for i, (c, m) in mine.items()}
    th.cuda.synchronize(device)
    final = {i: l for i, (l, _) in mine.items()}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, final, group=group)
        final = {k: val for d in gathered for k, val in d.items()}
    return mine, [final[i] for i in sorted(final)]
