"""GPU tests of the sharded modes.  Single-rank checks always run on a GPU box; the 2-rank NCCL checks run
when the box has >= 2 GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch as th

from tests import parity

pytestmark = pytest.mark.gpu


def T(a, dev):
    return th.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_view_sharded_single_rank_equals_fused():
    """world = 1: band-major l2_grad + adam_apply reproduces the fused kernel's optimisation."""
    import svbrdf_diff_renderer_b200 as pkg
    from svbrdf_diff_renderer_b200 import sharding
    dev = th.device("cuda:0")
    g = parity.golden("offaxis_32x9")
    cl = [T(g["cam"], dev), T(g["light"], dev), T(g["power"], dev)]
    vs = sharding.ViewShardedOptim(32, 9, float(g["size"]), cl, dev, bands=3)
    vs.load_targets(T(g["target"], dev))
    vs.init_from_tex(T(g["tex0"], dev))
    losses = vs.optim(int(g["epochs"]), float(g["lr"]))
    np.testing.assert_allclose(np.array(losses), g["loss_f64"], rtol=5e-5)
    parity.check_against_arbiter(vs.textures.cpu().numpy(), g["maps_f32"], g["maps_f64"], parity.RTOL_GRAD, "view-sharded maps",
                                 floor=2e-4, min_fraction=0.999)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["SVBRDF_B200_QUIET"] = "1"
    th.cuda.set_device(rank)
    dev = th.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        import svbrdf_diff_renderer_b200 as pkg
        from svbrdf_diff_renderer_b200 import sharding, synth
        g = parity.golden("coloc_24x16")
        cl = [T(g["cam"], dev), T(g["light"], dev), T(g["power"], dev)]
        # view-sharded: 16 lights over `world` ranks
        vs = sharding.ViewShardedOptim(24, 16, float(g["size"]), cl, dev, bands=2)
        vs.load_targets(T(g["target"], dev)[vs.start:vs.end])
        vs.init_from_tex(T(g["tex0"], dev))
        losses = vs.optim(int(g["epochs"]), float(g["lr"]))
        # the same problem with the collective fused into the kernels (peer stores over NVLink, no NCCL on the data path)
        ps = sharding.PeerShardedOptim(24, 16, float(g["size"]), cl, dev)
        ps.load_targets(T(g["target"], dev)[ps.start:ps.end])
        ps.init_from_tex(T(g["tex0"], dev))
        p2p_losses = ps.optim(int(g["epochs"]), float(g["lr"]))
        p2p_tex = ps.textures.clone().cpu()
        mc_used = ps.multicast
        assert ps.pull
        # and once more with the push all-gather (unicast peer stores) instead of pulling tiles from their owner
        pu = sharding.PeerShardedOptim(24, 16, float(g["size"]), cl, dev, multicast=False, pull=False)
        pu.load_targets(T(g["target"], dev)[pu.start:pu.end])
        pu.init_from_tex(T(g["tex0"], dev))
        pu.optim(int(g["epochs"]), float(g["lr"]))
        assert not pu.multicast
        uni_tex = pu.textures.clone().cpu()
        # material-sharded: 5 small materials
        def make(i):
            r = pkg.Microfacet(32, 9, synth.IM_SIZE_CM, [c.to(dev) for c in synth.calibration(9)], dev)
            with th.no_grad():
                tgt = r.eval(synth.random_textures(32, 100 + i).to(dev))
            return r, tgt, synth.random_textures(32, 200 + i).to(dev)
        mine, all_losses = sharding.optimise_materials(5, make, 5, 0.01, dev)
        # row bands in the peer-push mode (the building block of the 2-D decomposition) and, from 4 ranks up, the 2-D
        # decomposition itself, against the fused single-GPU kernel on a 96^2 x 16 problem (19.2 tiles, uint8 targets)
        R, N, E = 96, 16, 3
        cl96 = synth.calibration(N)
        r96 = pkg.Microfacet(R, N, synth.IM_SIZE_CM, [c.to(dev) for c in cl96], dev)
        with th.no_grad():
            t96 = (r96.eval(synth.random_textures(R, 11).to(dev)) * 255).round().to(th.uint8)
        s96 = synth.random_textures(R, 12)
        fused = pkg.SvbrdfOptim(dev, r96)
        fused.load_targets(t96)
        fused.init_from_tex(s96.to(dev))
        fused_losses = fused.optim(E, 0.01, None, False, progress=False)
        band_tex, band_losses = [], []
        for band in ((0, 40), (40, 96)):
            pb = sharding.PeerShardedOptim(R, N, synth.IM_SIZE_CM, [c.to(dev) for c in cl96], dev, band=band)
            pb.load_targets(t96[pb.start:pb.end, :, band[0]:band[1], :])
            pb.init_from_tex(s96)
            band_losses.append(pb.optim(E, 0.01))
            band_tex.append(pb.textures.cpu())
            del pb
        hyb = None
        if world >= 4:
            hy = sharding.HybridShardedOptim(R, N, synth.IM_SIZE_CM, [c.to(dev) for c in cl96], dev, light_shards=2)
            hy.load_targets(t96[hy.start:hy.end, :, hy.band[0]:hy.band[1], :])
            hy.init_from_tex(s96)
            hl = hy.optim(E, 0.01)
            hyb = {"losses": hl, "tex": hy.textures.cpu(), "band": hy.band, "lights": (hy.start, hy.end)}
        th.save({"losses": losses, "tex": vs.textures.cpu(), "mine": sorted(mine), "all": all_losses, "p2p_losses": p2p_losses, "p2p_tex": p2p_tex,
                 "uni_tex": uni_tex, "mc": mc_used, "fused_tex": fused.textures.detach().cpu(), "fused_losses": fused_losses,
                 "band_tex": th.cat(band_tex, 2), "band_losses": [a + b for a, b in zip(*band_losses)], "hybrid": hyb},
                f"{out}/r{rank}.pt")
    finally:
        dist.destroy_process_group()


def _check_against_fused(tex, losses, ref, what):
    """Sharded runs sum the lights in a different order than the fused kernel: equal up to fp32 rounding of the gradient,
    which Adam's normalised step can amplify only where |g| ~ rounding noise."""
    np.testing.assert_allclose(np.array(losses), np.array(ref["fused_losses"]), rtol=2e-6)
    d = (tex - ref["fused_tex"]).abs()
    frac = float((d <= 1e-5).float().mean())
    parity.record_margin(what, frac=frac, max_abs=float(d.max()))
    assert frac >= 0.999, f"{what}: only {frac * 100:.3f}% of the map elements within 1e-5 of the fused single-GPU run"


@pytest.mark.skipif(th.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_four_rank_hybrid_bands_x_light_shards(tmp_path):
    """2 row bands x 2 light shards on 4 GPUs: peer-push inside each band's group, nothing between bands."""
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(4, _free_port(), str(tmp_path)), nprocs=4, join=True)
    outs = [th.load(tmp_path / f"r{k}.pt") for k in range(4)]
    assert [o["hybrid"]["band"] for o in outs] == [(0, 48), (0, 48), (48, 96), (48, 96)]
    assert [o["hybrid"]["lights"] for o in outs] == [(0, 8), (8, 16), (0, 8), (8, 16)]
    for o in outs[1:]:
        assert th.equal(o["hybrid"]["tex"], outs[0]["hybrid"]["tex"]) and o["hybrid"]["losses"] == outs[0]["hybrid"]["losses"]
    _check_against_fused(outs[0]["hybrid"]["tex"], outs[0]["hybrid"]["losses"], outs[0], "4-rank hybrid 2x2 maps vs fused")
    _check_against_fused(outs[0]["band_tex"], outs[0]["band_losses"], outs[0], "4-rank banded peer-push maps vs fused")


@pytest.mark.skipif(th.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_view_and_material_sharding(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = th.load(tmp_path / "r0.pt"), th.load(tmp_path / "r1.pt")
    assert th.equal(a["band_tex"], b["band_tex"])
    _check_against_fused(a["band_tex"], a["band_losses"], a, "2-rank banded peer-push maps vs fused")
    g = parity.golden("coloc_24x16")
    assert th.equal(a["tex"], b["tex"]) and a["losses"] == b["losses"]          # replicas stay bit-identical
    np.testing.assert_allclose(np.array(a["losses"]), g["loss_f64"], rtol=5e-5)
    parity.check_against_arbiter(a["tex"].numpy(), g["maps_f32"], g["maps_f64"], parity.RTOL_GRAD, "2-rank view-sharded maps",
                                 floor=2e-4, min_fraction=0.999)
    # peer-push mode: replicas bit-identical by construction, same optimisation as the NCCL mode
    assert th.equal(a["p2p_tex"], b["p2p_tex"]) and a["p2p_losses"] == b["p2p_losses"]
    assert th.equal(a["uni_tex"], a["p2p_tex"]) and th.equal(b["uni_tex"], a["p2p_tex"])     # multicast == unicast all-gather
    print("NVLS multicast used:", a["mc"], b["mc"])
    np.testing.assert_allclose(np.array(a["p2p_losses"]), g["loss_f64"], rtol=5e-5)
    parity.check_against_arbiter(a["p2p_tex"].numpy(), g["maps_f32"], g["maps_f64"], parity.RTOL_GRAD, "2-rank peer-push maps",
                                 floor=2e-4, min_fraction=0.999)
    assert a["mine"] == [0, 2, 4] and b["mine"] == [1, 3]
    assert a["all"] == b["all"] and len(a["all"]) == 5 and all(np.isfinite(a["all"]))
