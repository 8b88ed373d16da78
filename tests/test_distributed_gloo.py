"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: stack / row-band partitioning, the per-band table-row
offsets and the one collective on the path (sum-all-reduce of the spatial sums and of the table gradient).
The per-band arithmetic is done by the CPU oracle here (the kernels need a GPU); what is under test is that the
partition + offsets + all-reduce reproduce the whole-image result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clair_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _band_sums(val, std, t, theta, r0, r1, h, w, thr):
    """(P, C, 5) sums of rows [r0, r1) with the band's flat offset, via the oracle's per-pair terms."""
    i, j, r = orc.exposure_pairs(t, thr)
    c = val.shape[1]
    out = np.zeros((len(i), c, 5))
    for ch in range(c):     # the flat offset depends on the channel (band of a (C, H, W) frame)
        rows = ((np.arange((r1 - r0) * w) + ch * h * w + r0 * w) % c).reshape(1, 1, r1 - r0, w)
        rows = np.broadcast_to(rows, val[:, ch:ch + 1, r0:r1].shape)
        f32, fp32, _, _ = orc.icrf_linear(val[:, ch:ch + 1, r0:r1], theta, rows=rows)
        # pair terms with the band's true table rows
        tm = _terms_with_rows(val[:, ch:ch + 1, r0:r1], std[:, ch:ch + 1, r0:r1], i, j, r, f32, fp32)
        mw = tm["mask"] * tm["wt"]
        out[:, ch, 0] = mw.sum((1, 2, 3))
        out[:, ch, 1] = (mw * tm["ell"]).sum((1, 2, 3))
        out[:, ch, 2] = (mw * tm["ell"] ** 2).sum((1, 2, 3))
        out[:, ch, 3] = (tm["mask"] * tm["err"]).sum((1, 2, 3))
        out[:, ch, 4] = tm["mask"].sum((1, 2, 3))
    return out


def _terms_with_rows(val, std, i, j, r, f32, fp32):
    f = f32.astype(np.float64)
    sig32 = np.abs(fp32 * std)
    rr = np.asarray(r).reshape(-1, 1, 1, 1)
    a, b = f[i], f[j]
    es = b * rr + 1e-6
    ell = np.abs((a - b * rr) / es)
    bs = np.maximum(f32[j], np.float32(1e-6)).astype(np.float64)
    err = np.sqrt((sig32[i].astype(np.float64) / es) ** 2 + ((f32[i] * sig32[j]).astype(np.float64) / (es * bs)) ** 2 + 1e-6)
    gw = orc.gaussian_value_weights(val, 10.0)
    wt = (gw[i] + gw[j]).astype(np.float64) + 1.0 / (err + 1e-6)
    mask = orc.pair_valid_mask(val, i, j, 1 / 255, 254 / 255).astype(np.float64)
    return {"ell": ell, "err": err, "wt": wt, "mask": mask}


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import clair_torch_b200 as ct
    from clair_torch_b200 import distributed as cd
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    h, w = 11, 9                                  # W not divisible by C, uneven bands
    val, std, t = ct.synthetic.make_stack(5, 3, h, w, bits=8, seed=4)
    val, std = val.numpy(), std.numpy()
    theta = ct.synthetic.reference_curve(3).numpy()
    r0, r1 = cd.row_band(h, rank, world)
    # the band's row offsets handed to the kernels agree with the flat-index rule
    base = cd.band_row_base(3, h, w, r0)
    assert all(base[c] == (c * h * w + r0 * w) % 3 for c in range(3))
    sums = torch.from_numpy(_band_sums(val, std, t, theta, r0, r1, h, w, 0.2))
    cd.all_reduce_sum_(sums)
    mean, sd, err = spatial_statistics(sums, True)
    _, o_mean, o_sd, o_err = orc.linearity_stats(val, std, t, theta, 0.2)
    ok = (np.max(np.abs(mean.numpy() - o_mean) / o_mean) < 1e-9 and np.max(np.abs(sd.numpy() - o_sd) / o_sd) < 1e-6
          and np.max(np.abs(err.numpy() - o_err) / o_err) < 1e-9)
    # stack sharding covers every stack exactly once
    mine = torch.zeros(13)
    mine[cd.stacks_for_rank(13, rank, world)] = 1
    cd.all_reduce_sum_(mine)
    ok = ok and bool((mine == 1).all())
    # a gradient-sized buffer reduces to the sum over ranks
    g = torch.full((3, 256), float(rank + 1), dtype=torch.float64)
    cd.all_reduce_sum_(g)
    ok = ok and bool((g == sum(range(1, world + 1))).all())
    with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
        fh.write("1" if ok else "0")
    dist.destroy_process_group()


def test_row_band_sums_allreduce_to_whole_image(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert open(tmp_path / f"ok{rank}").read() == "1"


def test_partitions():
    from clair_torch_b200 import distributed as cd
    for n, world in ((64, 8), (13, 4), (3, 8)):
        got = sum((cd.stacks_for_rank(n, r, world) for r in range(world)), [])
        assert got == list(range(n))
    for h, world in ((1080, 8), (11, 2), (7, 4)):
        bands = [cd.row_band(h, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == h and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
