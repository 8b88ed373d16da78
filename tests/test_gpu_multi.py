"""Multi-GPU tests (need >= 2 visible GPUs; skipped on a single-GPU box): row-band data-parallel training step and
linearity measurement over NCCL reproduce the single-GPU whole-image result; stack sharding of the merge needs no
collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import clair_torch_b200 as ct
    from clair_torch_b200 import distributed as cd
    h, w = 270, 480
    val, std, t = ct.synthetic.make_stack(6, 3, h, w, bits=8, seed=11)          # same stack on every rank (CPU generator)
    exposures = torch.from_numpy(t)
    r0, r1 = cd.row_band(h, rank, world)
    band_val, band_std = cd.take_band(val, r0, r1).to(dev), cd.take_band(std, r0, r1).to(dev)
    rb = cd.band_row_base(3, h, w, r0)
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=True, alpha=1.0, beta=1.0, gamma=1.0, delta=1.0,
              exposure_ratio_threshold=0.1)

    def fresh():
        m = ct.ICRFModelDirect(256, 3, initial_power=2.2).to(dev)
        return m, [torch.optim.Adam(m.channel_params(c), lr=1e-3) for c in range(3)]

    # (1) at a fixed table: loss, spatial means and table gradient of the banded + all-reduced pass == whole image
    from clair_torch_b200.training import linearity_loss_and_table_grad
    table0 = torch.stack([torch.linspace(0, 1, 256) ** (2.2 + 0.1 * c) for c in range(3)]).to(dev)
    full_val, full_std = val.to(dev), std.to(dev)
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(exposures, 0.1)
    red = lambda x: cd.all_reduce_sum_(x)
    lin_b, sp_b, g_b = linearity_loss_and_table_grad(band_val, band_std, i_idx, j_idx, ratio, table0, 1 / 255, 254 / 255, True,
                                                     True, row_base=rb, reduce_fn=red)
    lin_w, sp_w, g_w = linearity_loss_and_table_grad(full_val, full_std, i_idx, j_idx, ratio, table0, 1 / 255, 254 / 255, True,
                                                     True)
    checks = {"lin": bool(torch.allclose(lin_b, lin_w, rtol=1e-7, atol=0)),
              "spatial": bool(torch.allclose(sp_b, sp_w, rtol=1e-7, atol=0)),
              "grad": bool((g_b - g_w).abs().max() <= 1e-6 * g_w.abs().max())}
    # (2) data-parallel steps: the first two losses equal the single-GPU ones (later ones depend on how Adam's
    # sign-like first update treats bins whose gradient is ~1e-8, which the fp32 reduction order decides), and all
    # replicas hold bit-identical parameters after every step because they apply the same all-reduced gradient
    model_dp, opt_dp = fresh()
    losses_dp = [cd.train_icrf_step_data_parallel(model_dp, opt_dp, band_val, band_std, exposures, rb, **kw) for _ in range(4)]
    model_1, opt_1 = fresh()
    losses_1 = [ct.train_icrf_step(model_1, opt_1, full_val, full_std, exposures, **kw) for _ in range(4)]
    checks["loss01"] = all(bool(torch.allclose(a, b, rtol=1e-6, atol=0)) for a, b in zip(losses_dp[:2], losses_1[:2]))
    checks["loss_later"] = all(bool(torch.allclose(a, b, rtol=2e-2, atol=0)) for a, b in zip(losses_dp[2:], losses_1[2:]))
    mine_t = model_dp.icrf.detach().contiguous()
    gathered = [torch.empty_like(mine_t) for _ in range(world)]
    dist.all_gather(gathered, mine_t)
    checks["replicas_identical"] = all(bool(torch.equal(gathered[0], g)) for g in gathered)
    diff = (model_dp.icrf.detach() - model_1.icrf.detach()).abs().max().item()
    ok = all(checks.values())
    # linearity measurement on bands == whole image
    table = ct.synthetic.reference_curve(3).to(dev)
    _, m_b, s_b, e_b = cd.measure_linearity_band(band_val, band_std, exposures, table, rb)
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    from torch.utils.data import DataLoader
    loader = DataLoader(ExposureStackDataset(list(val), list(std), list(t)), batch_size=6, collate_fn=custom_collate)
    _, m_w, s_w, e_w = ct.measure_linearity(loader, dev, True, True, ct.ICRFModelDirect(icrf=table.cpu().clone()).to(dev))
    checks["linearity"] = all(bool(torch.allclose(a, b, rtol=1e-7, atol=0)) for a, b in ((m_b, m_w), (s_b, s_w), (e_b, e_w)))
    ok = ok and checks["linearity"]
    # merge sharded by stack: no collective, each rank's result equals what rank 0 would compute for that stack
    from clair_torch_b200 import kernels
    mine = cd.stacks_for_rank(4, rank, world)
    for sid in mine:
        v, s, tt = ct.synthetic.make_stack(4, 3, 64, 96, seed=100 + sid)
        rad, sig = kernels.hdr_merge_update(kernels.HdrMergeState(), v.to(dev), s.to(dev), tt, table, True, True)
        ok = ok and bool(torch.isfinite(rad).all()) and bool(torch.isfinite(sig).all())
    with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
        fh.write(f"{int(ok)} {diff} {checks}")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_row_band_data_parallel_matches_single_gpu(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        report = open(tmp_path / f"ok{rank}").read()
        assert report.split()[0] == "1", report
