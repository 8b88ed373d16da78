"""GPU tests at the BASELINE.json sizes: direct comparison with the C/OpenMP oracle (fast enough at full size) and
size-independent properties of the domain (exposure scaling, frame-order invariance, N=1 merge == linearisation,
row-band shards == whole image, linearity of the table gradient in the upstream)."""
import numpy as np
import pytest
import torch

from _helpers import max_abs_over_max, max_rel
from oracle import c_oracle as corc
from oracle import clair_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


@pytest.fixture(scope="module")
def ct():
    import clair_torch_b200 as pkg
    pkg._native.load()
    return pkg


def _merge(ct, val, std, t, theta, gaussian=True, dtype=torch.float32, row_base=None):
    from clair_torch_b200 import kernels
    return kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, gaussian, True, radiance_dtype=dtype,
                                    row_base=row_base)


def test_c1_full_size_against_c_oracle(ct):
    """Config 1: 5 x 8-bit RGB 1920x1080, 256-entry ICRF with distinct rows, Gaussian weights, uncertainty."""
    val, std, t = ct.synthetic.make_stack(5, 3, 1080, 1920, bits=8, seed=1234, device=DEV)
    theta = ct.synthetic.reference_curve(3)
    rad, sig = _merge(ct, val, std, t, theta.to(DEV))
    o_rad, o_sig = corc.hdr_merge(val.cpu().numpy(), std.cpu().numpy(), t, theta.numpy(), True)
    assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6
    assert max_rel(sig.cpu().numpy(), o_sig) < 5e-6
    assert torch.isfinite(rad).all() and torch.isfinite(sig).all() and (sig >= 0).all()


def test_c4_full_stack_against_c_oracle(ct):
    """Config 4: one whole 9-frame 24 MP (4000 x 6000) 16-bit stack, fp32 val + std (the N = 9 register kernel the bench
    times), every pixel against the C oracle; and the same stack as uint16 camera codes is bit-identical."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    val, std, t = ct.synthetic.make_stack(9, 3, 4000, 6000, bits=16, seed=4567, device=DEV)
    theta = ct.synthetic.reference_curve(3)
    rad, sig = _merge(ct, val, std, t, theta.to(DEV))
    o_rad, o_sig = corc.hdr_merge(val.cpu().numpy(), std.cpu().numpy(), t, theta.numpy(), True)
    assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6
    assert max_rel(sig.cpu().numpy(), o_sig) < 5e-6
    del o_rad, o_sig, std
    codes = torch.round(val * 65535.0).to(torch.int32).to(torch.uint16)
    del val
    rad_c, sig_c = kernels.hdr_merge_update(kernels.HdrMergeState(), codes, StdSpec("multiplier", 0.05), t, theta.to(DEV), True, True,
                                            radiance_dtype=torch.float32)
    assert torch.equal(rad_c, rad) and torch.equal(sig_c, sig)


def test_merge_properties_at_c1_size(ct):
    val, std, t = ct.synthetic.make_stack(5, 3, 1080, 1920, bits=8, seed=77, device=DEV)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    rad, sig = _merge(ct, val, std, t, theta)
    # (1) exposure scaling: t -> 4t (a power of two, exact in fp32) divides radiance and sigma by exactly 4
    rad4, sig4 = _merge(ct, val, std, t * 4.0, theta)
    assert torch.equal(rad4 * 4.0, rad) and torch.equal(sig4 * 4.0, sig)
    # (2) frame order inside the batch does not matter beyond fp32 summation order
    perm = [3, 0, 4, 1, 2]
    radp, sigp = _merge(ct, val[perm].contiguous(), std[perm].contiguous(), t[perm], theta)
    assert max_rel(radp.cpu().numpy(), rad.cpu().numpy()) < 1e-6
    assert max_rel(sigp.cpu().numpy(), sig.cpu().numpy(), 1e-30) < 1e-5
    # (3) row-band shards with the right table-row offsets reproduce the whole image bit for bit
    from clair_torch_b200 import kernels
    for r0, r1 in ((0, 135), (135, 541), (541, 1080)):
        rb = kernels.shard_row_base(3, 1080, 1920, r0)
        pr, ps = _merge(ct, val[:, :, r0:r1].contiguous(), std[:, :, r0:r1].contiguous(), t, theta, row_base=rb)
        assert torch.equal(pr, rad[:, r0:r1]) and torch.equal(ps, sig[:, r0:r1])
    # (4) float64 radiance output is the float32 one widened (single batch)
    rad64, _ = _merge(ct, val, std, t, theta, dtype=torch.float64)
    assert torch.equal(rad64.to(torch.float32), rad)


def test_single_frame_merge_equals_linearisation(ct):
    """N = 1, unit weights: mean = f(x)/t * (1/(1+1e-6)), sigma = |f'(x)| s /(t (1+1e-6)) — ties the merge kernel to
    the lineariser and the forward kernel."""
    from clair_torch_b200 import kernels
    val, std, t = ct.synthetic.make_stack(1, 3, 1080, 1920, bits=16, seed=5, device=DEV)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    rad, sig = _merge(ct, val, std, t, theta, gaussian=False)
    lin, lsig = kernels.linearize(val, std, theta)
    y, dydx = kernels.icrf_forward(val, theta, want_derivative=True)
    assert torch.equal(y, lin) and torch.equal(lsig, (dydx * std).abs())
    scale = 1.0 / (float(t[0]) * (1.0 + 1e-6))
    assert max_rel(rad.cpu().numpy(), lin[0].cpu().numpy().astype(np.float64) * scale, 1e-30) < 1e-6
    assert max_rel(sig.cpu().numpy(), lsig[0].cpu().numpy().astype(np.float64) * scale, 1e-30) < 1e-6


def test_multibatch_state_at_c1_size(ct):
    """Three DataLoader-style batches (2+2+1 frames) through the running state vs the C oracle's batch semantics."""
    from clair_torch_b200 import kernels
    val, std, t = ct.synthetic.make_stack(5, 3, 540, 1920, bits=8, seed=9, device=DEV)
    theta = ct.synthetic.reference_curve(3)
    st = kernels.HdrMergeState()
    out = None
    for a, b in ((0, 2), (2, 4), (4, 5)):
        out = kernels.hdr_merge_update(st, val[a:b].contiguous(), std[a:b].contiguous(), t[a:b], theta.to(DEV), True, b == 5)
    o_rad, o_sig = corc.hdr_merge(val.cpu().numpy(), std.cpu().numpy(), t, theta.numpy(), True, batch_size=2)
    assert out[0].dtype == torch.float64
    assert max_rel(out[0].cpu().numpy(), o_rad) < 2e-6
    assert max_rel(out[1].cpu().numpy(), o_sig) < 5e-6


def test_c2_training_step_full_size_against_c_oracle(ct):
    """Config 2: 10 x 1080p, script settings (P = 17): loss, spatial means and table gradient of one step."""
    from clair_torch_b200.training import linearity_loss_and_table_grad
    val, std, t = ct.synthetic.make_stack(10, 3, 1080, 1920, bits=8, seed=2345, device=DEV)
    theta = torch.stack([torch.linspace(0, 1, 256) ** (2.5 + 0.15 * c) for c in range(3)])
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.25)
    assert len(i) == 17
    for unc in (False, True):
        lin, spatial, grad = linearity_loss_and_table_grad(val, std, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, unc)
        o_lin, o_mean, o_grad = corc.train_grad(val.cpu().numpy(), std.cpu().numpy(), i.numpy(), j.numpy(), r.numpy(),
                                                theta.numpy(), relative=True, unc_weighting=unc)
        assert max_rel(lin.cpu().numpy(), o_lin) < 2e-6
        assert max_rel(spatial.cpu().numpy(), o_mean) < 2e-6
        assert max_abs_over_max(grad.cpu().numpy(), o_grad) < TOL


def _mask_counts_torch(val, i, j, lo, hi):
    """Population count of get_pairwise_valid_pixel_mask (common/general_functions.py:276-312) per (pair, channel) with
    plain torch compares on the device: inclusive bounds, compared in fp32."""
    lo32, hi32 = float(np.float32(lo)), float(np.float32(hi))
    ok = (val >= lo32) & (val <= hi32)
    return torch.stack([(ok[int(a)] & ok[int(b)]).sum(dim=(1, 2)) for a, b in zip(i, j)]).to(torch.float64)


def test_c3_linearity_full_size_against_c_oracle(ct):
    """Config 3: 16 exposures of 4K (2160 x 3840) 16-bit RGB, P = 29, uncertainty-weighted relative loss — all three
    outputs against the C oracle at full size, and the validity-mask population counts exactly."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    val, std, t = ct.synthetic.make_stack(16, 3, 2160, 3840, bits=16, seed=3456, device=DEV)
    theta = ct.synthetic.reference_curve(3)
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.2)
    assert len(i) == 29
    sums = kernels.pair_stats(val, std, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, True)
    mean, sd, err = spatial_statistics(sums, True)
    o_mean, o_sd, o_err = corc.pair_stats(val.cpu().numpy(), std.cpu().numpy(), i.numpy(), j.numpy(), r.numpy(), theta.numpy())
    assert max_rel(mean.cpu().numpy(), o_mean) < 2e-6
    assert max_rel(sd.cpu().numpy(), o_sd) < 2e-6
    assert max_rel(err.cpu().numpy(), o_err) < 2e-6
    assert torch.equal(sums[..., 4], _mask_counts_torch(val, i, j, 1 / 255, 254 / 255))
    # and against the numpy oracle's boolean mask on a crop
    mask = orc.pair_valid_mask(val[:, :, :64].cpu().numpy(), i.numpy(), j.numpy(), 1 / 255, 254 / 255)
    part = kernels.pair_stats(val[:, :, :64].contiguous(), std[:, :, :64].contiguous(), i, j, r, theta.to(DEV), 1 / 255,
                              254 / 255, True, True)
    assert np.array_equal(part[..., 4].cpu().numpy(), mask.sum(axis=(2, 3)).astype(np.float64))


def test_c5_exposure_pair_full_size_against_c_oracle(ct):
    """Config 5: one 100.7 MP (8192 x 12288) 16-bit exposure pair, script settings (P = 1): loss, spatial mean, mask count
    and table gradient of the whole image against the C oracle; a 4-band split with all-reduced (here: added) sums and
    gradient reproduces it — what the data-parallel step computes."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.training import linearity_loss_and_table_grad
    h, w = 8192, 12288
    val, std, _ = ct.synthetic.make_stack(2, 3, h, w, bits=16, seed=5678, device=DEV)
    t = np.array([0.01, 0.02])
    theta = torch.stack([torch.linspace(0, 1, 256) ** (2.5 + 0.15 * c) for c in range(3)])
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.25)
    assert len(i) == 1
    lin, spatial, grad = linearity_loss_and_table_grad(val, std, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, False)
    o_lin, o_mean, o_grad = corc.train_grad(val.cpu().numpy(), std.cpu().numpy(), i.numpy(), j.numpy(), r.numpy(), theta.numpy(),
                                            relative=True, unc_weighting=False)
    assert max_rel(lin.cpu().numpy(), o_lin) < 2e-6
    assert max_rel(spatial.cpu().numpy(), o_mean) < 2e-6
    assert max_abs_over_max(grad.cpu().numpy(), o_grad) < TOL
    sums = kernels.pair_stats(val, std, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, False)
    assert torch.equal(sums[..., 4], _mask_counts_torch(val, i, j, 1 / 255, 254 / 255))
    # row bands: sums added, upstream from the total, gradients added
    bands = [(0, 2048), (2048, 4100), (4100, 6001), (6001, h)]
    acc = torch.zeros((1, 3, 5), dtype=torch.float64, device=DEV)
    parts = []
    for r0, r1 in bands:
        rb = kernels.shard_row_base(3, h, w, r0)
        bv, bs = val[:, :, r0:r1].contiguous(), std[:, :, r0:r1].contiguous()
        parts.append((bv, bs, rb))
        kernels.pair_stats(bv, bs, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, False, row_base=rb, out=acc, means_only=True)
    lin_b, mean_b, up, mfg = kernels.pair_upstream(acc)
    g_acc = torch.zeros((3, 256), dtype=torch.float64, device=DEV)
    for bv, bs, rb in parts:
        kernels.pair_grad(bv, bs, i, j, r, theta.to(DEV), 1 / 255, 254 / 255, True, False, up, mfg, row_base=rb, out=g_acc)
    assert max_rel(lin_b.cpu().numpy(), o_lin) < 2e-6
    assert max_abs_over_max(g_acc.cpu().numpy(), o_grad) < TOL


def test_row_band_statistics_add_up(ct):
    """Sums of row bands (with their table-row offsets) add to the whole-image sums: the invariant the multi-GPU
    all-reduce relies on."""
    from clair_torch_b200 import kernels
    val, std, t = ct.synthetic.make_stack(6, 3, 301, 1000, bits=16, seed=21, device=DEV)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.1)
    whole = kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True)
    acc = torch.zeros_like(whole)
    for r0, r1 in ((0, 100), (100, 217), (217, 301)):
        rb = kernels.shard_row_base(3, 301, 1000, r0)
        kernels.pair_stats(val[:, :, r0:r1].contiguous(), std[:, :, r0:r1].contiguous(), i, j, r, theta, 1 / 255, 254 / 255,
                           True, True, row_base=rb, out=acc)
    assert torch.equal(acc[..., 4], whole[..., 4])
    # fp32 partial sums of <= 32 terms per lane, merged in float64: tilings differ by a few 1e-9 relative
    assert max_rel(acc.cpu().numpy()[..., :4], whole.cpu().numpy()[..., :4], 1e-300) < 1e-7


def test_lookup_and_forward_full_frame(ct):
    from clair_torch_b200 import kernels
    val, _, _ = ct.synthetic.make_stack(2, 3, 2160, 3840, bits=16, seed=31, device=DEV)
    theta = ct.synthetic.reference_curve(3)
    y = kernels.icrf_forward(val, theta.to(DEV))
    f, _, _ = corc.icrf_linear(val.cpu().numpy(), theta.numpy())
    assert np.array_equal(y.cpu().numpy(), f)
    yl = kernels.icrf_forward(val, theta.to(DEV), ct._native.INTERP_LOOKUP)
    fl, _ = corc.icrf_lookup(val.cpu().numpy(), theta.numpy())
    assert np.array_equal(yl.cpu().numpy(), fl)


def test_dark_corrected_c1_full_size(ct):
    """Config 1 with a dark field.  The mix itself (pre-pass kernel) against the numpy oracle at full size; the merge that
    mixes in its load (column-strip walk, rows per band chosen for the launch) against pre-pass + merge, which evaluates the
    same fp32 expressions (an fp64 mix would put a few of the 31 M mixed values on the other side of a table sample, where
    the slope f' — and with it sigma — jumps); bit for bit against itself at other band heights (the arithmetic per pixel
    does not depend on the band); and against the grid-stride form of the same kernel.  Also three frames and seven frames
    (no next-row register set) at odd sizes."""
    from clair_torch_b200 import kernels
    lib = ct._native.load()

    def with_knobs(fn, **knobs):
        try:
            for key, value in knobs.items():
                ct._native.check(lib.clair_set_tuning(key.encode(), value), "tune")
            return fn()
        finally:
            for key in knobs:
                lib.clair_set_tuning(key.encode(), 0)

    for n, h, w in ((5, 1080, 1920), (3, 431, 614), (7, 257, 362)):
        val, std, t = ct.synthetic.make_stack(n, 3, h, w, bits=16, seed=1234 + n, device=DEV)
        gen = torch.Generator(device=DEV).manual_seed(n)
        dark = torch.rand(val.shape, device=DEV, generator=gen) * 0.02
        hot = torch.rand(val.shape, device=DEV, generator=gen) < 0.08
        dark = torch.where(hot, 0.03 + 0.4 * torch.rand(val.shape, device=DEV, generator=gen), dark)
        dark_std = 0.1 * dark + 1e-3
        theta = ct.synthetic.reference_curve(3).to(DEV)
        mixed, seff = kernels.dark_field_mix(val, std, dark, dark_std)
        o_mixed, o_seff = orc.dark_field_mix(val.cpu().numpy(), std.cpu().numpy(), dark.cpu().numpy(), dark_std.cpu().numpy())
        assert max_rel(mixed.cpu().numpy(), o_mixed, 1e-12) < 2e-6 and max_rel(seff.cpu().numpy(), o_seff, 1e-12) < TOL   # (B(x) - x cancels)
        run = lambda: kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True, radiance_dtype=torch.float32,
                                               dark=(dark, dark_std))
        rad, sig = run()
        p_rad, p_sig = _merge(ct, mixed, seff, t, theta)
        e_rad, e_sig = max_rel(rad.cpu().numpy(), p_rad.cpu().numpy()), max_rel(sig.cpu().numpy(), p_sig.cpu().numpy())
        assert e_rad < 1e-6 and e_sig < 5e-6, (n, h, w, e_rad, e_sig)
        for rows in (6, 17, 128):
            r2, s2 = with_knobs(run, dark_rows=rows)
            assert torch.equal(r2, rad) and torch.equal(s2, sig)
        r3, s3 = with_knobs(run, dark_strip=-1)
        e_rad, e_sig = max_rel(r3.cpu().numpy(), rad.cpu().numpy()), max_rel(s3.cpu().numpy(), sig.cpu().numpy())
        assert e_rad < 1e-6 and e_sig < 5e-6, ("grid-stride form", n, h, w, e_rad, e_sig)
