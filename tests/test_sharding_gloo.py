"""CPU, world_size 2 over gloo: the multi-GPU partitioning logic of svbrdf_diff_renderer_b200.sharding.

The arithmetic engine is replaced by the oracle (tests only) so that the decomposition — contiguous light
shards with the GLOBAL MSE normaliser, band-major storage, per-band all-reduce, replicated Adam — is checked
against the one-piece oracle run without a GPU.  The product engine (NativeEngine) is exercised by the GPU
tests (tests/test_gpu_parity.py::test_view_sharded_single_rank_equals_fused and, with 2 GPUs,
tests/test_gpu_multi.py).
"""
import os
import socket

import numpy as np
import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

from svbrdf_diff_renderer_b200 import sharding, synth


def test_split_helpers():
    for n, w in ((64, 8), (9, 2), (9, 4), (256, 8), (5, 8)):
        got = [sharding.split_range(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        sizes = [e - s for s, e in got]
        assert max(sizes) - min(sizes) <= 1
    # 2-D decomposition: consecutive ranks share a band, light shard = rank within the band's group
    assert sharding.hybrid_layout(8, 2, 5) == (2, 1, [[0, 1], [2, 3], [4, 5], [6, 7]])
    assert sharding.hybrid_layout(8, 4, 5) == (1, 1, [[0, 1, 2, 3], [4, 5, 6, 7]])
    assert sharding.hybrid_layout(4, 4, 3) == (0, 3, [[0, 1, 2, 3]])
    with pytest.raises(ValueError):
        sharding.hybrid_layout(8, 3, 0)
    cover = sorted((b, s_) for b, s_, _ in (sharding.hybrid_layout(8, 2, r) for r in range(8)))
    assert cover == [(b, s_) for b in range(4) for s_ in range(2)]
    assert [len(sharding.round_robin(38, 8, r)) for r in range(8)] == [5, 5, 5, 5, 5, 5, 4, 4]
    assert sorted(sum((sharding.round_robin(38, 8, r) for r in range(8)), [])) == list(range(38))
    for res, b in ((1024, 4), (24, 4), (5, 8), (32, 1)):
        bands = sharding.row_bands(res, b)
        assert bands[0][0] == 0 and bands[-1][1] == res and all(x[1] == y[0] for x, y in zip(bands, bands[1:]))
    with pytest.raises(ValueError):
        sharding.split_range(4, 2, 2)


class OracleEngine:
    """TEST-ONLY engine: the oracle restatement on the CPU in place of the CUDA kernels."""

    def __init__(self, owner):
        self.o = owner

    def l2_grad_band(self, tex, targets, n_total, band, grad, loss_out):
        from oracle import torch_port as tp
        o = self.o
        sc = tp.Scene(o.res, o.local_cl[0], o.local_cl[1], o.local_cl[2], o.size, th.float32, band)
        loss, g, _, _ = tp.loss_and_grad(sc, tex[None], targets)
        # the band-only oracle averages over n_local*3*rows*R elements; the kernel contract is the global mean
        w = (o.n_local * (band[1] - band[0])) / (n_total * o.res)
        grad.copy_(g[0] * w)
        loss_out[0] = float(loss) * w

    def adam_apply(self, p, m, v, g, step, lr, b1=0.9, b2=0.999, eps=1e-8):
        # torch/optim/adam.py:531-547, single tensor
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)


def _worker(rank, world, port, n, res, epochs, out):
    os.environ.setdefault("SVBRDF_B200_QUIET", "1")
    th.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from oracle import torch_port as tp
        cl = synth.calibration(n, colocated=False)
        full_scene = tp.Scene(res, cl[0], cl[1], cl[2], synth.IM_SIZE_CM)
        target = tp.shade(full_scene, synth.random_textures(res, 1))
        tex0 = synth.random_textures(res, 2 + rank)          # different per rank: init_from_tex must broadcast rank 0's
        vs = sharding.ViewShardedOptim(res, n, synth.IM_SIZE_CM, cl, "cpu", engine_factory=OracleEngine, bands=3)
        vs.load_targets(target[vs.start:vs.end])
        vs.init_from_tex(tex0)
        losses = vs.optim(epochs, 0.01)
        th.save({"losses": losses, "tex": vs.textures, "range": (vs.start, vs.end)}, f"{out}/rank{rank}.pt")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_view_sharded_equals_one_piece_oracle(tmp_path):
    from oracle import torch_port as tp
    n, res, epochs, world = 9, 16, 4, 2
    mp.spawn(_worker, args=(world, _free_port(), n, res, epochs, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (th.load(tmp_path / f"rank{r}.pt") for r in range(world))
    assert r0["range"] == (0, 5) and r1["range"] == (5, 9)
    # replicas are bit-identical and share one loss curve
    assert th.equal(r0["tex"], r1["tex"]) and r0["losses"] == r1["losses"]
    # ... and equal the one-piece oracle run (fp32 summation order differs: tolerance, not bit-exactness)
    cl = synth.calibration(n, colocated=False)
    sc = tp.Scene(res, cl[0], cl[1], cl[2], synth.IM_SIZE_CM)
    target = tp.shade(sc, synth.random_textures(res, 1))
    maps, losses, _ = tp.optimise(sc, synth.random_textures(res, 2), target, epochs, 0.01)
    np.testing.assert_allclose(np.array(r0["losses"]), np.array(losses), rtol=2e-6)
    diff = (r0["tex"] - maps).abs()
    assert float(diff.mean()) < 1e-6 and float((diff > 1e-4).float().mean()) < 2e-3


def _material_worker(rank, world, port, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mine = sharding.round_robin(7, world, rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, {i: float(i) * 0.5 for i in mine})
        merged = {k: v for d in gathered for k, v in d.items()}
        th.save({"mine": mine, "all": [merged[i] for i in sorted(merged)]}, f"{out}/m{rank}.pt")
    finally:
        dist.destroy_process_group()


def test_material_sharding_bookkeeping(tmp_path):
    mp.spawn(_material_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = th.load(tmp_path / "m0.pt"), th.load(tmp_path / "m1.pt")
    assert a["mine"] == [0, 2, 4, 6] and b["mine"] == [1, 3, 5]
    assert a["all"] == b["all"] == [i * 0.5 for i in range(7)]
