"""Edge cases on the GPU: channel counts 1..4, table sizes, frame counts from 1 to the 64-frame limit (register,
shared-memory-parked and float64-sum kernels), odd image sizes, degenerate pixels, limits and error paths, and the
train_icrf driver loop end to end."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from _helpers import max_abs_over_max, max_rel
from oracle import c_oracle as corc
from oracle import clair_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ct():
    import clair_torch_b200 as pkg
    pkg._native.load()
    return pkg


def _merge(ct, val, std, t, theta, gaussian=True, **kw):
    from clair_torch_b200 import kernels
    return kernels.hdr_merge_update(kernels.HdrMergeState(), val.to(DEV), None if std is None else std.to(DEV), t,
                                    None if theta is None else theta.to(DEV), gaussian, True, radiance_dtype=torch.float32, **kw)


@pytest.mark.parametrize("channels", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("lut", [16, 256, 1024])
def test_merge_channel_counts_and_table_sizes(ct, channels, lut):
    val, std, t = ct.synthetic.make_stack(4, channels, 33, 52, bits=16, seed=channels * 7 + lut)
    theta = ct.synthetic.reference_curve(channels, lut)
    rad, sig = _merge(ct, val, std, t, theta)
    o_rad, o_sig = corc.hdr_merge(val.numpy(), std.numpy(), t, theta.numpy(), True)
    assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6 and max_rel(sig.cpu().numpy(), o_sig) < 5e-6


@pytest.mark.parametrize("n_frames", [1, 2, 8, 9, 13, 27, 40, 41, 64])
def test_merge_frame_counts(ct, n_frames):
    """1..8 register kernel, 9..40 shared-memory-parked kernel, 41..64 float64-sum kernel."""
    val, std, _ = ct.synthetic.make_stack(n_frames, 3, 24, 40, bits=16, seed=n_frames)
    t = 1e-3 * 1.19 ** np.arange(n_frames)
    theta = ct.synthetic.reference_curve(3)
    rad, sig = _merge(ct, val, std, t, theta)
    o_rad, o_sig = corc.hdr_merge(val.numpy(), std.numpy(), t, theta.numpy(), True)
    assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6 and max_rel(sig.cpu().numpy(), o_sig) < 5e-6


def test_limits_and_bad_arguments(ct):
    from clair_torch_b200 import kernels
    val, std, t = ct.synthetic.make_stack(3, 3, 8, 8)
    theta = ct.synthetic.reference_curve(3)
    big = torch.zeros((65, 1, 4, 4))
    with pytest.raises(ValueError, match="limit"):
        _merge(ct, big, big, np.ones(65), None)
    with pytest.raises(ValueError, match="limit"):
        _merge(ct, torch.zeros((2, 9, 4, 4)), torch.zeros((2, 9, 4, 4)), np.ones(2), None)
    with pytest.raises(ValueError):
        _merge(ct, val, std, t[:2], theta)                               # exposure count mismatch
    with pytest.raises(ValueError):
        _merge(ct, val, std[:, :, :4], t, theta)                         # std shape mismatch
    with pytest.raises(ValueError):
        _merge(ct, val, std, t, ct.synthetic.reference_curve(2))          # table rows != channels
    with pytest.raises(TypeError):
        _merge(ct, val.double(), std, t, theta)
    with pytest.raises(ValueError):
        kernels.pair_stats(val.to(DEV), std.to(DEV), [0, 7], [1, 2], [0.5, 0.5], theta.to(DEV), 0.0, 1.0, True, True)


def test_degenerate_pixels(ct):
    """All-saturated, all-black and out-of-range pixels: finite results that match the oracle; a pixel whose Gaussian
    weights all underflow gives NaN like the reference's 0/0."""
    val = torch.tensor([0.0, 1.0, 0.5, -0.2, 1.3, 1e-9, 254.5 / 255]).view(1, 1, 1, 7).repeat(3, 3, 2, 1).contiguous()
    val[1] = val[1] * 0.5 + 0.25
    std = torch.full_like(val, 0.01)
    t = np.array([1e-3, 2e-3, 4e-3])
    theta = ct.synthetic.reference_curve(3)
    rad, sig = _merge(ct, val, std, t, theta)
    o_rad, o_sig = orc.hdr_merge(val.numpy(), std.numpy(), t, theta.numpy(), True)
    assert torch.isfinite(rad).all() and torch.isfinite(sig).all()
    assert max_rel(rad.cpu().numpy(), o_rad, 1e-12) < 2e-6 and max_rel(sig.cpu().numpy(), o_sig, 1e-12) < 5e-6
    far = torch.full((2, 1, 1, 4), 5.0)                                  # exp(-30 * 4.5^2) underflows to 0
    rad, sig = _merge(ct, far, torch.ones_like(far), np.array([1.0, 2.0]), None)
    assert torch.isnan(rad).all()


def test_pair_kernels_many_pairs_and_odd_sizes(ct):
    """More pairs than one launch carries (78 > 64), H*W not a multiple of 4 (scalar staging), C = 1 and 4."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    from clair_torch_b200.training import linearity_loss_and_table_grad
    for channels, n, h, w in ((1, 13, 9, 11), (4, 5, 10, 13), (3, 13, 16, 16)):
        val, std, _ = ct.synthetic.make_stack(n, channels, h, w, bits=16, seed=n + channels)
        t = 1e-3 * 1.3 ** np.arange(n)
        theta = ct.synthetic.reference_curve(channels)
        i, j, r = orc.exposure_pairs(t, 0.0)
        sums = kernels.pair_stats(val.to(DEV), std.to(DEV), i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, True)
        mean, sd, err = spatial_statistics(sums.cpu(), True)
        o_mean, o_sd, o_err = corc.pair_stats(val.numpy(), std.numpy(), i, j, r, theta.numpy())
        assert max_rel(mean.numpy(), o_mean) < 2e-6 and max_rel(err.numpy(), o_err) < 2e-6
        # pairs with a tiny exposure ratio have an almost constant loss (std << mean): the fp32 rounding of each loss
        # shows in the spread, so the bound is relative to the scale the rounding has (see test_pair_stats_degenerate_spread)
        assert np.all(np.abs(sd.numpy() - o_sd) <= 1e-5 * o_sd + 2e-8 * o_mean)
        lin, spatial, grad = linearity_loss_and_table_grad(val.to(DEV), std.to(DEV), i, j, r, theta.to(DEV), 1 / 255, 254 / 255,
                                                           True, True)
        o_lin, o_m, o_grad = corc.train_grad(val.numpy(), std.numpy(), i, j, r, theta.numpy())
        assert max_rel(lin.cpu().numpy(), o_lin) < 2e-6
        assert max_abs_over_max(grad.cpu().numpy(), o_grad) < 1e-5


@pytest.mark.parametrize("n,pairs,h,w", [(2, "all", 30, 36), (3, "first", 8, 100), (7, "sparse", 24, 40), (12, "sparse", 20, 52),
                                         (20, "first", 12, 44), (32, "sparse", 8, 36), (33, "sparse", 8, 36), (9, "all", 5, 28)])
def test_packed_pair_kernels_block_shapes(ct, n, pairs, h, w):
    """The packed (fp32x2) pair kernels over their block shapes: a single exposure pair (one-warp blocks), fewer warps than
    frames (two staging items per thread), 32 frames (the most the packed statistics kernel holds) and 33 (scalar
    fallback), planes that end inside a tile and inside a warp's 64-pixel group; with and without the uncertainty term,
    relative and absolute loss.  Checked against the C oracle."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    from clair_torch_b200.training import linearity_loss_and_table_grad
    val, std, _ = ct.synthetic.make_stack(n, 3, h, w, bits=16, seed=100 + n)
    t = 1e-3 * 1.25 ** np.arange(n)
    theta = ct.synthetic.reference_curve(3)
    i, j, r = orc.exposure_pairs(t, 0.0)
    if pairs == "first":
        i, j, r = i[:1], j[:1], r[:1]
    elif pairs == "sparse":
        i, j, r = i[::5], j[::5], r[::5]
    dv, ds, dt = val.to(DEV), std.to(DEV), theta.to(DEV)
    for relative, unc in ((True, True), (True, False), (False, True)):
        sums = kernels.pair_stats(dv, ds, i, j, r, dt, 1 / 255, 254 / 255, relative, unc)
        mean, sd, err = spatial_statistics(sums.cpu(), True)
        o_mean, o_sd, o_err = corc.pair_stats(val.numpy(), std.numpy(), i, j, r, theta.numpy(), relative=relative, unc_weighting=unc)
        assert max_rel(mean.numpy(), o_mean) < 2e-6 and max_rel(err.numpy(), o_err) < 2e-6
        assert np.all(np.abs(sd.numpy() - o_sd) <= 1e-5 * o_sd + 2e-8 * o_mean)
        lin, spatial, grad = linearity_loss_and_table_grad(dv, ds, i, j, r, dt, 1 / 255, 254 / 255, relative, unc)
        o_lin, o_m, o_grad = corc.train_grad(val.numpy(), std.numpy(), i, j, r, theta.numpy(), relative=relative, unc_weighting=unc)
        assert max_rel(lin.cpu().numpy(), o_lin) < 2e-6
        assert max_abs_over_max(grad.cpu().numpy(), o_grad) < 1e-5


def test_all_pairs_masked_gives_zero_loss_and_gradient(ct):
    from clair_torch_b200.training import linearity_loss_and_table_grad
    val = torch.zeros((3, 3, 8, 8))                                      # every pixel below the validity threshold
    std = torch.zeros_like(val)
    t = torch.tensor([1e-3, 2e-3, 4e-3], dtype=torch.float64)
    i, j, r = ct.common.get_valid_exposure_pairs(t, 0.1)
    lin, spatial, grad = linearity_loss_and_table_grad(val.to(DEV), std.to(DEV), i, j, r, ct.synthetic.reference_curve(3).to(DEV),
                                                       1 / 255, 254 / 255, True, True)
    assert torch.equal(lin, torch.zeros_like(lin)) and torch.equal(spatial, torch.zeros_like(spatial))
    assert torch.equal(grad, torch.zeros_like(grad))


def test_train_icrf_driver_loop(ct):
    """train_icrf end to end: default per-channel Adam, early stopping bookkeeping, a ReduceLROnPlateau scheduler; the
    loss goes down on a stack whose true response is a 2.2 gamma when starting from a 2.5 power curve."""
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    val, std, t = ct.synthetic.make_stack(6, 3, 64, 96, bits=8, seed=2)
    loader = DataLoader(ExposureStackDataset(list(val), list(std), list(t)), batch_size=6, collate_fn=custom_collate)
    from clair_torch_b200.training import linearity_loss_and_table_grad
    model = ct.ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3) for c in range(3)]
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.25)

    def objective():
        return linearity_loss_and_table_grad(val.to(DEV), std.to(DEV), i, j, r, model.icrf.detach(), 1 / 255, 254 / 255, True,
                                             False, want_grad=False)[0].sum().item()

    before = objective()
    out = ct.train_icrf(loader, 6, DEV, model, opts, None, use_uncertainty_weighting=False, epochs=40, patience=50,
                        alpha=10.0, exposure_ratio_threshold=0.25, verbose=False)
    assert out is model
    assert objective() < 0.8 * before
    # schedulers are stepped with the per-channel epoch loss (icrf_training.py:174-176); early stopping ends the run
    scheds = [torch.optim.lr_scheduler.ReduceLROnPlateau(o, patience=0) for o in opts]
    ct.train_icrf(loader, 6, DEV, model, opts, scheds, use_uncertainty_weighting=False, epochs=6, patience=2, verbose=False)
    assert all(o.param_groups[0]["lr"] <= 1e-3 for o in opts)
    # defaults: optimisers created inside, one epoch
    ct.train_icrf(loader, 6, DEV, ct.ICRFModelDirect().to(DEV), epochs=2, verbose=False)
    with pytest.raises(ValueError, match="Mismatched"):
        ct.train_icrf(loader, 6, DEV, model, opts, [None], epochs=1, verbose=False)


def test_kernels_do_not_write_outside_their_outputs(ct):
    """compute-sanitizer is closed on this pool, so out-of-bounds WRITES are looked for with guard bands: every output
    of the C ABI is carved out of the middle of a sentinel-filled buffer and the guards must survive, on image sizes
    that exercise the scalar, 2-wide and 4-wide kernels and partially filled tiles."""
    import ctypes
    lib = ct._native.load()
    guard = 4096
    sentinel = -12345.0

    def carve(numel, dtype=torch.float32):
        buf = torch.full((numel + 2 * guard,), sentinel, dtype=dtype, device=DEV)
        return buf, buf[guard:guard + numel]

    def intact(buf, numel):
        return bool((buf[:guard] == sentinel).all()) and bool((buf[guard + numel:] == sentinel).all())

    theta = ct.synthetic.reference_curve(3).to(DEV)
    stream = torch.cuda.current_stream().cuda_stream
    for n, h, w in ((5, 17, 23), (3, 16, 20), (9, 9, 14), (4, 33, 2)):
        val, std, _ = ct.synthetic.make_stack(n, 3, h, w, bits=16, seed=n)
        v, s = val.to(DEV), std.to(DEV)
        t = np.ascontiguousarray(1e-3 * 1.2 ** np.arange(n))
        plane, numel = h * w, 3 * h * w
        rb, rad = carve(numel)
        sb, sig = carve(numel)
        rc = lib.clair_hdr_merge_update(v.data_ptr(), s.data_ptr(), t.ctypes.data_as(ctypes.c_void_p), n, theta.data_ptr(), 3, 256,
                                        plane, None, 1, None, None, None, 1, 1, rad.data_ptr(), 0, sig.data_ptr(), stream)
        assert rc == 0
        lb, lin = carve(n * numel)
        gb, lsig = carve(n * numel)
        assert lib.clair_linearize(v.data_ptr(), s.data_ptr(), theta.data_ptr(), lin.data_ptr(), lsig.data_ptr(), n, 3, plane, 256,
                                   2, None, stream) == 0
        mb, mixed = carve(n * numel)
        eb, seff = carve(n * numel)
        dark = torch.rand_like(v) * 0.1
        assert lib.clair_dark_field_mix(v.data_ptr(), s.data_ptr(), dark.data_ptr(), dark.data_ptr(), n, 3, h, w, 0.05, 50.0,
                                        mixed.data_ptr(), seff.data_ptr(), stream) == 0
        states = [carve(numel) for _ in range(4)]
        assert lib.clair_frame_stats_update(v.data_ptr(), None, theta.data_ptr(), n, 3, plane, 256, 2, None, states[0][1].data_ptr(),
                                            states[1][1].data_ptr(), states[2][1].data_ptr(), states[3][1].data_ptr(), 1, stream) == 0
        i, j, r = orc.exposure_pairs(t, 0.0)
        i32, j32 = np.ascontiguousarray(i, dtype=np.int32), np.ascontiguousarray(j, dtype=np.int32)
        p = len(i)
        sums_b, sums = carve(p * 3 * 5, torch.float64)
        if n <= 16:
            assert lib.clair_pair_stats(v.data_ptr(), s.data_ptr(), n, 3, plane, i32.ctypes.data_as(ctypes.c_void_p),
                                        j32.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p), p, theta.data_ptr(), 256,
                                        2, None, 1 / 255, 254 / 255, 1, 1, sums.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        assert intact(rb, numel) and intact(sb, numel) and intact(lb, n * numel) and intact(gb, n * numel)
        assert intact(mb, n * numel) and intact(eb, n * numel) and all(intact(b, numel) for b, _ in states)
        assert intact(sums_b, p * 3 * 5)


def test_staged_host_pipelines_back_to_back_without_sync(ct):
    """The staged entry points recycle device staging buffers (torch's caching allocator hands the same blocks to the next
    call) while copies and kernels of the previous call may still be in flight: 12 merges and 12 linearisations of
    alternating inputs are queued without any synchronisation and every result is checked afterwards."""
    from clair_torch_b200 import kernels
    theta = ct.synthetic.reference_curve(3).to(DEV)
    stacks = [ct.synthetic.make_stack(5, 3, 120, 200, bits=16, seed=s) for s in (1, 2, 3)]
    want = [kernels.hdr_merge_update(kernels.HdrMergeState(), v.to(DEV), s.to(DEV), t, theta, True, True, radiance_dtype=torch.float32)
            for v, s, t in stacks]
    want_lin = [kernels.linearize(v.to(DEV), s.to(DEV), theta) for v, s, _ in stacks]
    pinned = [(v.pin_memory(), s.pin_memory(), t) for v, s, t in stacks]
    outs, lins = [], []
    for k in range(12):
        v, s, t = pinned[k % 3]
        rad = torch.empty((3, 120, 200), dtype=torch.float32).pin_memory()
        sig = torch.empty_like(rad).pin_memory()
        kernels.hdr_merge_update(kernels.HdrMergeState(), v, s, t, theta, True, True, radiance_dtype=torch.float32,
                                 device=torch.device(DEV), host_out=(rad, sig), staged=True, bands=5)
        outs.append((rad, sig))
        lins.append(kernels.linearize(v, s, theta, device=torch.device(DEV), pinned_out=True, staged=True, bands=7))
    torch.cuda.synchronize()
    for k in range(12):
        assert torch.equal(outs[k][0], want[k % 3][0].cpu()) and torch.equal(outs[k][1], want[k % 3][1].cpu()), k
        assert torch.equal(lins[k][0], want_lin[k % 3][0].cpu()) and torch.equal(lins[k][1], want_lin[k % 3][1].cpu()), k


class _RecyclingPinnedBatches(torch.utils.data.Dataset):
    """Whole pinned batches handed out from a pool that behaves like torch's caching host allocator: a buffer whose
    tensor nobody references any more is considered free and is immediately overwritten (poisoned) before it is reused."""

    def __init__(self, batches):
        import weakref
        self.batches, self.weakref = batches, weakref
        self.slots, self.poisoned = [], 0          # slots: [val base, std base, weakref of the handed-out view]

    def __len__(self):
        return len(self.batches)

    def __getitem__(self, i):
        idx, val, std, meta = self.batches[i]
        slot = None
        for s in self.slots:
            if s[2] is not None and s[2]() is None:            # released by the consumer: "freed" pinned memory
                s[0].fill_(float("nan"))
                s[1].fill_(float("nan"))
                s[2] = None
                self.poisoned += 1
            if s[2] is None and slot is None and s[0].shape == val.shape:
                slot = s
        if slot is None:
            slot = [torch.empty(val.shape, dtype=val.dtype).pin_memory(), torch.empty(std.shape, dtype=std.dtype).pin_memory(), None]
            self.slots.append(slot)
        slot[0].copy_(val)
        slot[1].copy_(std)
        view = slot[0].view(val.shape)                          # a fresh tensor object over the pooled pinned storage
        slot[2] = self.weakref.ref(view)
        return idx, view, slot[1].view(std.shape), meta


@pytest.mark.parametrize("staged", [True, False])
def test_pinned_batches_stay_alive_until_the_gpu_has_read_them(ct, staged):
    """compute_hdr_image hands pinned host pointers to the copy engine / the kernels; it must keep every batch referenced
    until an event behind its last reader has completed.  The pool above poisons a buffer the moment it is dropped, as a
    DataLoader pin thread re-using a freed block would."""
    n, per_batch = 12, 3
    val, std, t = ct.synthetic.make_stack(n, 3, 512, 1024, bits=16, seed=77)
    theta = ct.synthetic.reference_curve(3)
    batches = []
    for b in range(0, n, per_batch):
        sl = slice(b, b + per_batch)
        batches.append((torch.arange(b, b + per_batch), val[sl].contiguous(), std[sl].contiguous(),
                        {"exposure_time": torch.from_numpy(t[sl])}))
    ds = _RecyclingPinnedBatches(batches)
    loader = DataLoader(ds, batch_size=None, shuffle=False)
    model = ct.ICRFModelDirect(icrf=theta.clone()).to(DEV)
    rad, sig = ct.compute_hdr_image(loader, DEV, model, max, staged=staged)
    torch.cuda.synchronize()
    dev_loader = DataLoader(ct.datasets.ExposureStackDataset(list(val.to(DEV)), list(std.to(DEV)), list(t)), batch_size=per_batch,
                            shuffle=False, collate_fn=ct.datasets.custom_collate)
    want_rad, want_sig = ct.compute_hdr_image(dev_loader, DEV, model, max)
    assert not torch.isnan(rad).any() and not torch.isnan(sig).any()
    assert torch.equal(rad, want_rad) and torch.equal(sig, want_sig)


def test_host_out_results_are_ready_when_compute_hdr_image_returns(ct):
    """host_out buffers are written by the kernel (or a non_blocking copy): the call synchronises before returning them."""
    val, std, t = ct.synthetic.make_stack(6, 3, 1024, 1536, bits=16, seed=5)
    theta = ct.synthetic.reference_curve(3)
    model = ct.ICRFModelDirect(icrf=theta.clone()).to(DEV)
    loader = DataLoader(ct.datasets.ExposureStackDataset(list(val.to(DEV)), list(std.to(DEV)), list(t)), batch_size=6, shuffle=False,
                        collate_fn=ct.datasets.custom_collate)
    want_rad, want_sig = ct.compute_hdr_image(loader, DEV, model, max, radiance_dtype=torch.float32)
    torch.cuda.synchronize()
    for _ in range(3):
        rad_h = torch.full((3, 1024, 1536), float("nan")).pin_memory()
        sig_h = torch.full((3, 1024, 1536), float("nan")).pin_memory()
        rad, sig = ct.compute_hdr_image(loader, DEV, model, max, radiance_dtype=torch.float32, host_out=(rad_h, sig_h))
        got_rad, got_sig = rad.numpy().copy(), sig.numpy().copy()          # read immediately, no synchronize
        assert np.array_equal(got_rad, want_rad.cpu().numpy()) and np.array_equal(got_sig, want_sig.cpu().numpy())


def test_single_pair_step_falls_back_when_it_cannot_fuse(ct):
    """An odd plane (or std images with uncertainty weighting) keeps the two-pass kernels; both routes agree with the oracle."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.training import linearity_loss_and_table_grad
    val, std, _ = ct.synthetic.make_stack(2, 3, 7, 9, bits=16, seed=4)           # 63 pixels: odd
    t = np.array([0.01, 0.02])
    theta = ct.synthetic.reference_curve(3)
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.25)
    dv, ds = val.to(DEV), std.to(DEV)
    assert not kernels.can_fuse_pair(dv, ds, 1, False)
    for unc in (False, True):
        lin, spatial, grad = linearity_loss_and_table_grad(dv, ds, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, unc)
        o = orc.train_loss_and_grad(val.numpy(), std.numpy(), t, theta.numpy(), 0.25, relative=True, unc_weighting=unc)
        assert max_rel(lin.cpu().numpy(), o["linloss"]) < 1e-5 and max_abs_over_max(grad.cpu().numpy(), o["grad_lin"]) < 1e-5
    with pytest.raises(ValueError):
        kernels.pair_fused(dv, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True)                  # odd plane
    val6, _, t6 = ct.synthetic.make_stack(3, 3, 8, 8, bits=8, seed=1)
    i6, j6, r6 = ct.common.get_valid_exposure_pairs(torch.from_numpy(t6), 0.1)
    with pytest.raises(ValueError):
        kernels.pair_fused(val6.to(DEV), i6, j6, r6, theta.to(DEV), 1 / 255, 254 / 255, True)    # more than one pair


def test_integer_code_batches_reject_what_they_cannot_take(ct):
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import ExposureStackDataset, StdSpec, custom_collate
    codes = torch.randint(0, 255, (3, 3, 5, 7), dtype=torch.uint8)                # 315 elements: not a multiple of 4
    with pytest.raises(ValueError):
        kernels.expand_codes(codes.to(DEV))
    with pytest.raises(TypeError):
        kernels.expand_codes(codes.to(torch.float32).to(DEV))
    with pytest.raises(TypeError):
        kernels.expand_codes(torch.zeros(8, dtype=torch.uint8, device=DEV), std=0.05)
    # a LOOKUP model cannot carry a std — synthesised or not — through measure_linearity (the reference's autograd raises)
    good = torch.randint(0, 255, (3, 3, 8, 8), dtype=torch.uint8)
    ds = ExposureStackDataset(list(good), StdSpec("multiplier", 0.05), [0.01, 0.02, 0.04])
    model = ct.ICRFModelDirect(256, 3, ct.InterpMode.LOOKUP).to(DEV)
    with pytest.raises(RuntimeError):
        ct.measure_linearity(DataLoader(ds, batch_size=3, collate_fn=custom_collate), DEV, True, True, model)
    # a StdSpec needs integer codes to be evaluated from
    ds_f = ExposureStackDataset(list(good.to(torch.float32) / 255), StdSpec("multiplier", 0.05), [0.01, 0.02, 0.04])
    with pytest.raises(ValueError):
        ct.measure_linearity(DataLoader(ds_f, batch_size=3, collate_fn=custom_collate), DEV, True, True, None)


def test_grid_sizes_do_not_change_results(ct):
    """The launch grids are tuned per kernel (blocks that walk several tiles, grids of 1x .. 32x the resident blocks): results
    must not depend on them.  Elementwise kernels bit for bit, the pair sums to the rounding of their fp32 partial sums."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    lib = ct._native.load()

    def with_knobs(fn, **knobs):
        try:
            for key, value in knobs.items():
                ct._native.check(lib.clair_set_tuning(key.encode(), value), "tune")
            return fn()
        finally:
            for key in knobs:
                lib.clair_set_tuning(key.encode(), 0)

    val, std, t = ct.synthetic.make_stack(6, 3, 310, 404, bits=16, seed=77, device=DEV)       # 125 240 pixels: tiles with a tail
    theta = ct.synthetic.reference_curve(3).to(DEV)
    base = kernels.linearize(val, std, theta)
    for blocks in (-1, 1, 3, 50):
        got = with_knobs(lambda: kernels.linearize(val, std, theta), fwd_blocks=blocks)
        assert all(torch.equal(a, b) for a, b in zip(got, base))
    codes = torch.round(val * 65535.0).to(torch.int32).to(torch.uint16)
    base_codes = kernels.expand_codes(codes, StdSpec("multiplier", 0.05), 65535.0)
    for waves in (1, 7):
        got = with_knobs(lambda: kernels.expand_codes(codes, StdSpec("multiplier", 0.05), 65535.0), aux_waves=waves)
        assert all(torch.equal(a, b) for a, b in zip(got, base_codes))
    dark = torch.rand_like(val) * 0.1
    dark_std = dark * 0.1 + 1e-3
    base_dark = kernels.dark_field_mix(val, std, dark, dark_std)
    for waves in (1, 3):
        got = with_knobs(lambda: kernels.dark_field_mix(val, std, dark, dark_std), aux_waves=waves)
        assert all(torch.equal(a, b) for a, b in zip(got, base_dark))
    i, j, r = orc.exposure_pairs(t, 0.2)
    base_sums = kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True)
    for waves in (1, 2, 8):
        got = with_knobs(lambda: kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True), stats_waves=waves)
        assert max_rel(got.cpu().numpy(), base_sums.cpu().numpy(), 1e-300) < 5e-6      # fp32 partial sums per (lane, block)
    base_merge = kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True, radiance_dtype=torch.float32)
    for waves in (1, 5):
        got = with_knobs(lambda: kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True,
                                                          radiance_dtype=torch.float32), hdr_waves=waves)
        assert all(torch.equal(a, b) for a, b in zip(got, base_merge))
