"""GPU parity tests: the sm_100a kernels, called through the public API / C ABI, against
(a) the fixtures written by the unmodified reference (tests/golden/) and (b) the pinned CPU oracle on larger
seeded inputs.  Bars (BASELINE.json north_star): indices, masks and pure table look-ups bit-exact;
radiance, uncertainty, loss statistics within 1e-5 relative; table gradient within 1e-5 of max|grad|.
"""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from _helpers import SIGMA_TOL, golden, golden_names, linear_golden_names, max_abs_over_max, max_rel, mode_of
from oracle import clair_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ct():
    import clair_torch_b200 as pkg
    pkg._native.load()          # fail loudly if the extension is missing
    return pkg


def _opt(z, k):
    return z[k] if k in z else None


def _loader(ct, z, batch_size):
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    vals = [torch.from_numpy(v) for v in z["val"]]
    stds = None if "std" not in z else [torch.from_numpy(s) for s in z["std"]]
    ds = ExposureStackDataset(vals, stds, list(z["exposure"]))
    return DataLoader(ds, batch_size=batch_size, shuffle=False, collate_fn=custom_collate)


def _model(ct, theta, mode=None):
    from clair_torch_b200 import ICRFModelDirect, InterpMode
    if isinstance(mode, str):
        mode = {"linear": InterpMode.LINEAR, "lookup": InterpMode.LOOKUP, "catmull": InterpMode.CATMULL}[mode]
    return ICRFModelDirect(icrf=torch.from_numpy(theta).clone(), interpolation_mode=mode or InterpMode.LINEAR).to(DEV)


# ---- ICRF evaluation ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("forward_linear"))
def test_forward_linear_bit_exact(ct, name):
    z = golden(name)
    model = _model(ct, z["theta"])
    x = torch.from_numpy(z["x"]).to(DEV).requires_grad_(True)
    y = model(x)
    (dydx,) = torch.autograd.grad(y, x, torch.ones_like(y))
    assert np.array_equal(y.detach().cpu().numpy(), z["y"])
    assert np.array_equal(dydx.cpu().numpy(), z["dydx"])


@pytest.mark.parametrize("name", golden_names("forward_lookup"))
def test_forward_lookup_bit_exact(ct, name):
    z = golden(name)
    model = _model(ct, z["theta"], ct.InterpMode.LOOKUP)
    y = model(torch.from_numpy(z["x"]).to(DEV))
    assert np.array_equal(y.cpu().numpy(), z["y"])


@pytest.mark.parametrize("name", golden_names("forward_codes"))
def test_forward_every_code(ct, name):
    z = golden(name)
    x = torch.from_numpy(z["x"]).to(DEV)
    assert np.array_equal(_model(ct, z["theta"])(x).cpu().numpy(), z["y_linear"])
    assert np.array_equal(_model(ct, z["theta"], ct.InterpMode.LOOKUP)(x).cpu().numpy(), z["y_lookup"])


@pytest.mark.parametrize("name", golden_names("forward_catmull"))
def test_forward_catmull_golden(ct, name):
    """CATMULL mode: value bit-exact, derivative and table gradient against the reference's CPU autograd."""
    z = golden(name)
    c = z["theta"].shape[0]
    model = ct.ICRFModelDirect(256, c, ct.InterpMode.CATMULL).to(DEV)
    with torch.no_grad():
        for k, p in enumerate(model.direct_params):
            p.copy_(torch.from_numpy(z["theta"][k]))
    model.update_icrf()
    x = torch.from_numpy(z["x"]).to(DEV).requires_grad_(True)
    y = model(x)
    (dydx,) = torch.autograd.grad(y, x, torch.ones_like(y), retain_graph=True)
    (y * torch.from_numpy(z["upstream"]).to(DEV)).sum().backward()
    gtheta = torch.stack([p.grad for p in model.direct_params]).cpu().numpy()
    assert np.array_equal(y.detach().cpu().numpy(), z["y"])
    assert max_abs_over_max(dydx.cpu().numpy(), z["dydx"]) < 1e-4
    assert max_abs_over_max(gtheta, z["grad_theta"]) < TOL


def test_pca_model_table_and_training_step(ct):
    """ICRFModelPCA builds a (C, L) table from exponents + PCA coefficients and trains through the fused step."""
    from clair_torch_b200 import ICRFModelPCA, train_icrf_step
    L, K, C = 256, 3, 3
    x = torch.linspace(0, 1, L)
    basis = torch.stack([torch.sin((k + 1) * torch.pi * x) for k in range(K)], dim=1).unsqueeze(-1).repeat(1, 1, C) * 0.01
    model = ICRFModelPCA(basis).to(DEV)
    model.update_icrf()
    assert tuple(model.icrf.shape) == (C, L)
    want = x.clamp(min=1e-6).pow(2.0)
    assert torch.allclose(model.icrf[0].cpu(), want, atol=1e-6)
    val, std, t = ct.synthetic.make_stack(5, 3, 48, 64, seed=3)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-2) for c in range(C)]
    before = [p.detach().clone() for p in model.parameters()]
    for _ in range(3):
        loss = train_icrf_step(model, opts, val.to(DEV), std.to(DEV), torch.from_numpy(t))
    assert loss.shape == (3,) and torch.isfinite(loss).all()
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, model.parameters()))


def test_forward_table_gradient_matches_scatter(ct):
    """d/d table of sum(g * f(x)) equals the two-tap scatter of the oracle (index_put of models/base.py:176)."""
    rng = np.random.default_rng(5)
    x = rng.uniform(-0.05, 1.05, size=(3, 3, 37, 53)).astype(np.float32)
    g = rng.normal(size=x.shape).astype(np.float32)
    theta = ct.synthetic.reference_curve(3).numpy()
    model = _model(ct, theta)
    model.update_icrf()   # connect the table to the parameters
    with torch.no_grad():
        for c, p in enumerate(model.direct_params):
            p.copy_(torch.from_numpy(theta[c]))
    model.update_icrf()
    y = model(torch.from_numpy(x).to(DEV))
    y.backward(torch.from_numpy(g).to(DEV))
    got = torch.stack([p.grad for p in model.direct_params]).cpu().numpy().astype(np.float64)
    _, _, x0, rows = orc.icrf_linear(x, theta)
    xs = np.clip(x * np.float32(255), 0, 255).astype(np.float32)
    w = (xs - x0.astype(np.float32)).astype(np.float64)
    want = np.zeros((3, 256))
    np.add.at(want, (rows.ravel(), x0.ravel()), (g * (1 - w)).ravel())
    np.add.at(want, (rows.ravel(), np.minimum(x0 + 1, 255).ravel()), (g * w).ravel())
    assert max_abs_over_max(got, want) < TOL


# ---- HDR merge ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("hdr_"))
def test_hdr_merge_golden(ct, name):
    z = golden(name)
    model = _model(ct, z["theta"], mode_of(z)) if "theta" in z else None
    weight_fn = (lambda img: img) if int(z["gaussian"]) else None
    rad, sig = ct.compute_hdr_image(_loader(ct, z, int(z["batch_size"])), DEV, model, weight_fn)
    assert rad.dtype == torch.float64 and tuple(rad.shape) == z["radiance"].shape
    assert max_rel(rad.cpu().numpy(), z["radiance"]) < TOL
    if "sigma" in z:
        assert sig.dtype == torch.float32
        assert max_rel(sig.cpu().numpy(), z["sigma"]) < SIGMA_TOL[mode_of(z)]      # reference noise per mode, _helpers.py
    else:
        assert sig is None


@pytest.mark.parametrize("mode,batch,gaussian", [("lookup", None, True), ("lookup", 3, True), ("catmull", None, True),
                                                 ("catmull", 4, True), ("catmull", None, False)])
def test_hdr_merge_modes_vs_oracle(ct, mode, batch, gaussian):
    """LOOKUP / CATMULL models through the all-modes kernel against the float64 closed form, at the LINEAR bar (1e-5):
    the kernel takes the cancelling difference v_n - mean_B in float64."""
    val, std, t = ct.synthetic.make_stack(7, 3, 61, 83, bits=16, seed=77)
    theta = ct.synthetic.reference_curve(3).numpy()
    want_r, want_s = orc.hdr_merge(val.numpy(), std.numpy(), t, theta, gaussian, batch, mode=mode)
    z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
    rad, sig = ct.compute_hdr_image(_loader(ct, z, batch or 7), DEV, _model(ct, theta, mode), (lambda i: i) if gaussian else None)
    assert max_rel(rad.cpu().numpy(), want_r) < TOL
    assert max_rel(sig.cpu().numpy(), want_s) < TOL


def test_hdr_merge_lookup_without_weights_raises_like_reference(ct):
    val, std, t = ct.synthetic.make_stack(3, 3, 8, 12, seed=1)
    z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
    model = _model(ct, ct.synthetic.reference_curve(3).numpy(), "lookup")
    with pytest.raises(RuntimeError):
        ct.compute_hdr_image(_loader(ct, z, 3), DEV, model, None)
    z.pop("std")
    rad, sig = ct.compute_hdr_image(_loader(ct, z, 3), DEV, model, None)          # no std: nothing to differentiate, fine
    want, _ = orc.hdr_merge(val.numpy(), None, t, model.icrf.cpu().numpy(), False, mode="lookup")
    assert sig is None and max_rel(rad.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("bits,batch", [(8, None), (16, None), (8, 2), (16, 4)])
def test_hdr_merge_vs_oracle_larger(ct, bits, batch):
    val, std, t = ct.synthetic.make_stack(5 if bits == 8 else 6, 3, 135, 240, bits=bits, seed=99 + bits)
    theta = ct.synthetic.reference_curve(3)
    z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
    rad, sig = ct.compute_hdr_image(_loader(ct, z, batch or len(t)), DEV, _model(ct, theta.numpy()), max)
    o_rad, o_sig = orc.hdr_merge(z["val"], z["std"], t, theta.numpy(), True, batch)
    assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6
    assert max_rel(sig.cpu().numpy(), o_sig) < 5e-6


def test_hdr_merge_fp32_radiance_and_flags(ct):
    val, std, t = ct.synthetic.make_stack(4, 3, 32, 48, seed=3)
    z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
    theta = ct.synthetic.reference_curve(3)
    for gaussian in (True, False):
        for with_model in (True, False):
            model = _model(ct, theta.numpy()) if with_model else None
            rad, sig = ct.compute_hdr_image(_loader(ct, z, 4), DEV, model, (max if gaussian else None),
                                            radiance_dtype=torch.float32)
            assert rad.dtype == torch.float32
            o_rad, o_sig = orc.hdr_merge(z["val"], z["std"], t, theta.numpy() if with_model else None, gaussian)
            assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6
            assert max_rel(sig.cpu().numpy(), o_sig) < 5e-6


def test_hdr_merge_odd_sizes_use_scalar_path(ct):
    """H*W not divisible by 4 or 2 exercises the VEC=2 / VEC=1 kernels and the k-mod-C row arithmetic."""
    theta = ct.synthetic.reference_curve(3)
    for h, w in ((7, 13), (6, 9), (5, 5)):
        val, std, t = ct.synthetic.make_stack(3, 3, h, w, seed=h * w)
        z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
        rad, sig = ct.compute_hdr_image(_loader(ct, z, 3), DEV, _model(ct, theta.numpy()), max)
        o_rad, o_sig = orc.hdr_merge(z["val"], z["std"], t, theta.numpy(), True)
        assert max_rel(rad.cpu().numpy(), o_rad) < 2e-6
        assert max_rel(sig.cpu().numpy(), o_sig) < 5e-6


def test_hdr_merge_row_band_shards_equal_whole_image(ct):
    """Spatial shards with the right curve_row_base reproduce the whole-image result bit for bit."""
    from clair_torch_b200 import kernels
    val, std, t = ct.synthetic.make_stack(4, 3, 40, 50, seed=8)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    full = kernels.hdr_merge_update(kernels.HdrMergeState(), val.to(DEV), std.to(DEV), t, theta, True, True)
    for r0, r1 in ((0, 13), (13, 27), (27, 40)):
        rb = kernels.shard_row_base(3, 40, 50, r0)
        part = kernels.hdr_merge_update(kernels.HdrMergeState(), val[:, :, r0:r1].contiguous().to(DEV),
                                        std[:, :, r0:r1].contiguous().to(DEV), t, theta, True, True, row_base=rb)
        assert torch.equal(part[0], full[0][:, r0:r1])
        assert torch.equal(part[1], full[1][:, r0:r1])


# ---- linearisation -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("linearize_"))
def test_linearize_bit_exact(ct, name):
    z = golden(name)
    model = _model(ct, z["theta"], mode_of(z))
    outs = list(ct.linearize_dataset_generator(_loader(ct, z, 1), DEV, model))
    assert len(outs) == z["val"].shape[0]
    for n, (lin, sig, meta) in enumerate(outs):
        assert lin.device.type == "cpu" and sig.device.type == "cpu"
        assert np.array_equal(lin.numpy(), z["linearized"][n])
        if mode_of(z) == "catmull":      # closed-form derivative vs the reference's fp32 autograd of the cubic weights
            assert np.max(np.abs(sig.numpy() - z["sigma"][n])) <= 5e-5 * np.max(z["sigma"][n])
        else:
            assert np.array_equal(sig.numpy(), z["sigma"][n])
        assert float(meta["exposure_time"][0]) == float(z["exposure"][n])


@pytest.mark.parametrize("shape,bands", [((2, 3, 37, 53), 4), ((1, 3, 64, 96), 16), ((3, 1, 5, 7), 2), ((1, 3, 40, 1000), 64)])
def test_linearize_staged_host_pipeline_is_bit_identical(ct, shape, bands):
    """Pinned host frames in, pinned host results out through the three-stream band pipeline (and through the zero-copy
    kernel): the same bits as the device-resident call, for every interpolation mode."""
    from clair_torch_b200 import kernels
    n, c, h, w = shape
    val, std, _ = ct.synthetic.make_stack(n, c, h, w, bits=16, seed=h)
    theta = ct.synthetic.reference_curve(c).to(DEV)
    for mode in (ct._native.INTERP_LINEAR, ct._native.INTERP_CATMULL, ct._native.INTERP_LOOKUP):
        s_dev = None if mode == ct._native.INTERP_LOOKUP else std.to(DEV)
        s_host = None if mode == ct._native.INTERP_LOOKUP else std.pin_memory()
        want = kernels.linearize(val.to(DEV), s_dev, theta, interp_mode=mode)
        for staged in (True, False):
            got = kernels.linearize(val.pin_memory(), s_host, theta, device=torch.device(DEV), pinned_out=True, interp_mode=mode,
                                    staged=staged, bands=bands)
            torch.cuda.current_stream().synchronize()
            assert got[0].is_pinned() and torch.equal(got[0], want[0].cpu()) and torch.equal(got[1], want[1].cpu())


# ---- linearity measurement ---------------------------------------------------------------------------
def _linearity_flags(name):
    if name == "linearity_u16_w11":
        return True, True
    return "rel1" in name, "unc1" in name


@pytest.mark.parametrize("name", golden_names("linearity_"))
def test_measure_linearity_golden(ct, name):
    z = golden(name)
    rel, unc = _linearity_flags(name)
    model = _model(ct, z["theta"], mode_of(z)) if "theta" in z else None
    ratio, mean, std, err = ct.measure_linearity(_loader(ct, z, len(z["exposure"])), DEV, unc, rel, model)
    assert np.array_equal(ratio.cpu().numpy(), z["ratio"])
    assert max_rel(mean.cpu().numpy(), z["mean"]) < TOL
    assert max_rel(std.cpu().numpy(), z["stddev"]) < TOL
    if "errmean" in z:
        assert max_rel(err.cpu().numpy(), z["errmean"]) < TOL
    else:
        assert err is None


@pytest.mark.parametrize("bits,n,relative,unc", [(8, 7, True, True), (16, 8, True, True), (16, 6, False, True),
                                                  (8, 10, True, False)])
def test_pair_stats_vs_oracle_larger(ct, bits, n, relative, unc):
    from clair_torch_b200 import kernels
    val, std, t = ct.synthetic.make_stack(n, 3, 96, 130, bits=bits, seed=7 * n)
    theta = ct.synthetic.reference_curve(3)
    i_idx, j_idx, ratio = orc.exposure_pairs(t, 0.1)
    sums = kernels.pair_stats(val.to(DEV), std.to(DEV), i_idx, j_idx, ratio, theta.to(DEV), 1 / 255, 254 / 255,
                              relative, unc).cpu().numpy()
    # validity mask: bit-exact through its per-(pair, channel) population count
    mask = orc.pair_valid_mask(val.numpy(), i_idx, j_idx, 1 / 255, 254 / 255)
    assert np.array_equal(sums[..., 4], mask.sum(axis=(2, 3)).astype(np.float64))
    _, o_mean, o_std, o_err = orc.linearity_stats(val.numpy(), std.numpy(), t, theta.numpy(), 0.1, relative=relative,
                                                  unc_weighting=unc)
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    mean, sd, err = spatial_statistics(torch.from_numpy(sums), True)
    assert max_rel(mean.numpy(), o_mean) < 2e-6
    assert max_rel(sd.numpy(), o_std) < 2e-6
    assert max_rel(err.numpy(), o_err) < 2e-6


def test_pair_stats_degenerate_spread(ct):
    """Identity linearisation of gamma-encoded 16-bit data: the relative loss is almost the same number at every
    pixel (std << mean), the case a one-pass sum-of-squares loses without float64 accumulation."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    val, std, t = ct.synthetic.make_stack(5, 3, 64, 64, bits=16, seed=12)
    i_idx, j_idx, ratio = orc.exposure_pairs(t, 0.2)
    sums = kernels.pair_stats(val.to(DEV), std.to(DEV), i_idx, j_idx, ratio, None, 1 / 255, 254 / 255, True, True)
    mean, sd, err = spatial_statistics(sums.cpu(), True)
    _, o_mean, o_std, o_err = orc.linearity_stats(val.numpy(), std.numpy(), t, None, 0.2)
    assert max_rel(mean.numpy(), o_mean) < 2e-6
    # here std/mean ~ 3e-4: the fp32 rounding of each per-pixel loss (~3e-8 absolute at l ~ 0.46) shows up in the
    # spread, so the bound is stated relative to the mean the rounding scales with
    assert np.max(np.abs(sd.numpy() - o_std) / (o_std + 2e-4 * o_mean)) < TOL


# ---- training step -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("trainstep_catmull"))
def test_train_step_catmull_golden(ct, name):
    """A CATMULL model trains through the same fused pair kernels as a LINEAR one (four-tap evaluation and scatter): loss
    and the updated table of every recorded step, teacher-forced like the LINEAR test below."""
    from clair_torch_b200 import ICRFModelDirect, InterpMode, train_icrf_step
    z = golden(name)
    val, std = torch.from_numpy(z["val"]).to(DEV), torch.from_numpy(z["std"]).to(DEV)
    exposure = torch.from_numpy(z["exposure"])
    a, b, g, d = (float(k) for k in z["coeffs"])
    for step in range(int(z["n_steps"])):
        theta = z["theta0"] if step == 0 else z[f"theta_after_{step - 1}"]
        model = ICRFModelDirect(256, 3, InterpMode.CATMULL).to(DEV)
        with torch.no_grad():
            for c, p in enumerate(model.direct_params):
                p.copy_(torch.from_numpy(theta[c]))
        model.update_icrf()
        opts = [torch.optim.SGD(model.channel_params(c), lr=1.0) for c in range(3)]     # theta - grad: exposes the gradient
        loss = train_icrf_step(model, opts, val, std, exposure, use_relative_linearity_loss=bool(z["rel"]),
                               use_uncertainty_weighting=bool(z["unc"]), alpha=a, beta=b, gamma=g, delta=d,
                               exposure_ratio_threshold=float(z["thr"]))
        assert max_rel(loss.cpu().numpy(), z[f"loss_{step}"]) < TOL
        grad = theta.astype(np.float64) - model.icrf.detach().cpu().numpy().astype(np.float64)
        want = z[f"grad_theta_{step}"]
        assert np.max(np.abs(grad - want)) < 1e-5 * np.max(np.abs(want)) + 2e-7       # fp32 parameter update rounding


@pytest.mark.parametrize("mode,with_std", [("catmull", True), ("catmull", False), ("lookup", False)])
@pytest.mark.parametrize("shape", [(6, 3, 40, 64), (5, 3, 21, 35), (7, 1, 30, 50), (4, 4, 17, 26)])
def test_pair_kernels_all_modes_vs_oracle(ct, mode, with_std, shape):
    """The fused pair kernels with LOOKUP / CATMULL models (models/base.py:138-158, :184-226) against the numpy oracle:
    statistics in all flag combinations and the table gradient — packed kernels (H*W % 4 == 0), the scalar ones (odd
    planes) and channel counts other than 3."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    n, c, h, w = shape
    val, std, t = ct.synthetic.make_stack(n, c, h, w, bits=16, seed=n * 100 + w)
    if not with_std:
        std = None
    theta = ct.synthetic.reference_curve(c)
    code = {"lookup": ct._native.INTERP_LOOKUP, "catmull": ct._native.INTERP_CATMULL}[mode]
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.1)
    dv, ds = val.to(DEV), None if std is None else std.to(DEV)
    for relative in (True, False):
        for unc in (True, False):
            sums = kernels.pair_stats(dv, ds, i_idx, j_idx, ratio, theta.to(DEV), 1 / 255, 254 / 255, relative, unc, interp_mode=code)
            mean, sd, err = spatial_statistics(sums, with_std)
            _, o_mean, o_sd, o_err = orc.linearity_stats(val.numpy(), None if std is None else std.numpy(), t, theta.numpy(), 0.1,
                                                         relative=relative, unc_weighting=unc, mode=mode)
            assert max_rel(mean.cpu().numpy(), o_mean) < TOL and max_rel(sd.cpu().numpy(), o_sd) < TOL
            if with_std:
                assert max_rel(err.cpu().numpy(), o_err) < TOL
            from clair_torch_b200.training import linearity_loss_and_table_grad
            lin, spatial, grad = linearity_loss_and_table_grad(dv, ds, i_idx, j_idx, ratio, theta.to(DEV), 1 / 255, 254 / 255,
                                                               relative, unc, interp_mode=code)
            o = orc.train_loss_and_grad(val.numpy(), None if std is None else std.numpy(), t, theta.numpy(), 0.1, relative=relative,
                                        unc_weighting=unc, mode=mode)
            assert max_rel(lin.cpu().numpy(), o["linloss"]) < TOL
            assert max_abs_over_max(grad.cpu().numpy(), o["grad_lin"]) < TOL


@pytest.mark.parametrize("mode", ["catmull", "lookup"])
def test_row_band_and_graphed_training_in_every_mode(ct, mode):
    """Row-band sharding (sums and gradient added over bands, the all-reduce of the data-parallel step) and the CUDA-graph
    step work for LOOKUP / CATMULL models like they do for LINEAR ones."""
    from clair_torch_b200 import ICRFModelDirect, InterpMode, kernels
    from clair_torch_b200.training import linearity_loss_and_table_grad
    imode = {"catmull": InterpMode.CATMULL, "lookup": InterpMode.LOOKUP}[mode]
    code = {"lookup": ct._native.INTERP_LOOKUP, "catmull": ct._native.INTERP_CATMULL}[mode]
    val, std, t = ct.synthetic.make_stack(6, 3, 60, 100, bits=16, seed=91)
    dv = val.to(DEV)
    ds = std.to(DEV) if mode == "catmull" else None
    theta = ct.synthetic.reference_curve(3).to(DEV)
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.1)
    lin_w, sp_w, g_w = linearity_loss_and_table_grad(dv, ds, i_idx, j_idx, ratio, theta, 1 / 255, 254 / 255, True, True, interp_mode=code)
    acc = torch.zeros((len(i_idx), 3, 5), dtype=torch.float64, device=DEV)
    bands = [(0, 17), (17, 41), (41, 60)]
    for r0, r1 in bands:
        rb = kernels.shard_row_base(3, 60, 100, r0)
        kernels.pair_stats(dv[:, :, r0:r1].contiguous(), None if ds is None else ds[:, :, r0:r1].contiguous(), i_idx, j_idx, ratio, theta,
                           1 / 255, 254 / 255, True, True, row_base=rb, out=acc, means_only=True, interp_mode=code)
    lin_b, mean_b, up, mfg = kernels.pair_upstream(acc)
    g_b = torch.zeros((3, 256), dtype=torch.float64, device=DEV)
    for r0, r1 in bands:
        rb = kernels.shard_row_base(3, 60, 100, r0)
        kernels.pair_grad(dv[:, :, r0:r1].contiguous(), None if ds is None else ds[:, :, r0:r1].contiguous(), i_idx, j_idx, ratio, theta,
                          1 / 255, 254 / 255, True, True, up, mfg, row_base=rb, out=g_b, interp_mode=code)
    assert torch.allclose(lin_b, lin_w, rtol=1e-7, atol=0) and (g_b - g_w).abs().max() <= 1e-6 * g_w.abs().max()
    # graph replay == eager steps (8-bit codes sit exactly on table samples, so no table entry receives a gradient that is
    # pure rounding noise — Adam's first steps move every entry by +-lr whatever the magnitude of its gradient)
    def fresh():
        m = ICRFModelDirect(256, 3, imode, 2.5).to(DEV)
        return m, [torch.optim.Adam(m.channel_params(c), lr=1e-3, capturable=True) for c in range(3)]
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=True, alpha=10.0, exposure_ratio_threshold=0.2)
    val, std, t = ct.synthetic.make_stack(6, 3, 96, 128, bits=8, seed=77)
    dv = val.to(DEV)
    ds = std.to(DEV) if mode == "catmull" else None
    m1, o1 = fresh()
    eager = [ct.train_icrf_step(m1, o1, dv, ds, torch.from_numpy(t), **kw) for _ in range(5)]
    m2, o2 = fresh()
    first = [ct.train_icrf_step(m2, o2, dv, ds, torch.from_numpy(t), **kw) for _ in range(2)]
    step = ct.GraphedTrainStep(m2, o2, dv, ds, torch.from_numpy(t), **kw)
    replayed = [step() for _ in range(3)]
    # (four-tap CATMULL weights of both signs: the order of the fp32 reductions shows in the 6th digit after a few Adam steps)
    for a, b in zip(eager, first + replayed):
        assert max_rel(b.cpu().numpy(), a.cpu().numpy()) < 1e-5
    assert max_abs_over_max(m2.icrf.detach().cpu().numpy(), m1.icrf.detach().cpu().numpy()) < 1e-5


@pytest.mark.parametrize("mode", ["linear", "catmull", "lookup"])
@pytest.mark.parametrize("relative", [True, False])
def test_single_pair_fused_pass_matches_two_pass_and_oracle(ct, mode, relative):
    """One exposure pair without uncertainty weights takes clair_pair_fused (statistics + un-normalised gradient in one pass,
    upstream applied afterwards): equal to the two-pass kernels and to the numpy oracle; row bands accumulate into one
    buffer (the single all-reduce of the data-parallel step)."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.training import linearity_loss_and_table_grad
    code = {"linear": ct._native.INTERP_LINEAR, "lookup": ct._native.INTERP_LOOKUP, "catmull": ct._native.INTERP_CATMULL}[mode]
    h, w = 50, 84
    val, std, _ = ct.synthetic.make_stack(2, 3, h, w, bits=16, seed=17)
    if mode == "lookup":
        std = None
    t = np.array([0.01, 0.02])
    theta = ct.synthetic.reference_curve(3).to(DEV)
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.25)
    assert len(i_idx) == 1
    dv, ds = val.to(DEV), None if std is None else std.to(DEV)
    assert kernels.can_fuse_pair(dv, ds, 1, False)
    lin, spatial, grad = linearity_loss_and_table_grad(dv, ds, i_idx, j_idx, ratio, theta, 1 / 255, 254 / 255, relative, False,
                                                       interp_mode=code)
    o = orc.train_loss_and_grad(val.numpy(), None if std is None else std.numpy(), t, theta.cpu().numpy(), 0.25, relative=relative,
                                unc_weighting=False, mode=mode)
    assert max_rel(lin.cpu().numpy(), o["linloss"]) < TOL and max_rel(spatial.cpu().numpy(), o["spatial"]) < TOL
    assert max_abs_over_max(grad.cpu().numpy(), o["grad_lin"]) < TOL
    # the two-pass kernels (what more than one pair takes)
    sums = kernels.pair_stats(dv, ds, i_idx, j_idx, ratio, theta, 1 / 255, 254 / 255, relative, False, means_only=True, interp_mode=code)
    lin2, mean2, up, mfg = kernels.pair_upstream(sums)
    grad2 = kernels.pair_grad(dv, ds, i_idx, j_idx, ratio, theta, 1 / 255, 254 / 255, relative, False, up, mfg, interp_mode=code)
    assert max_rel(lin.cpu().numpy(), lin2.cpu().numpy()) < 1e-6 and max_abs_over_max(grad.cpu().numpy(), grad2.cpu().numpy()) < 2e-6
    # row bands into one buffer
    buf = None
    for r0, r1 in ((0, 13), (13, 30), (30, h)):
        rb = kernels.shard_row_base(3, h, w, r0)
        buf = kernels.pair_fused(dv[:, :, r0:r1].contiguous(), i_idx, j_idx, ratio, theta, 1 / 255, 254 / 255, relative, row_base=rb,
                                 out=buf, interp_mode=code)
    lin_b, _, grad_b = kernels.pair_fused_combine(buf, 3, 256)
    assert max_rel(lin_b.cpu().numpy(), lin.cpu().numpy()) < 1e-7 and max_abs_over_max(grad_b.cpu().numpy(), grad.cpu().numpy()) < 1e-6
    # with uncertainty weights (and std images) the step keeps the two-pass kernels
    assert not kernels.can_fuse_pair(dv, dv, 1, True) and not kernels.can_fuse_pair(dv, ds, 2, False)


def test_lookup_model_errors_and_table_gradient(ct):
    """LOOKUP has no image edge: drivers that propagate std images through the model raise like the reference's autograd
    call; the table edge (a gather) exists and matches a scatter of the upstream gradient."""
    from clair_torch_b200 import ICRFModelDirect, InterpMode
    val, std, t = ct.synthetic.make_stack(3, 3, 12, 20, seed=2)
    z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
    model = _model(ct, ct.synthetic.reference_curve(3).numpy(), "lookup")
    with pytest.raises(RuntimeError):
        list(ct.linearize_dataset_generator(_loader(ct, z, 1), DEV, model))
    with pytest.raises(RuntimeError):
        ct.measure_linearity(_loader(ct, z, 3), DEV, True, True, model)
    m = ICRFModelDirect(256, 3, InterpMode.LOOKUP).to(DEV)
    m.update_icrf()
    g = torch.randn(val.shape, device=DEV)
    m(val.to(DEV)).backward(g)
    got = torch.stack([p.grad for p in m.direct_params]).cpu().numpy()
    _, idx = orc.icrf_lookup(val.numpy(), m.icrf.detach().cpu().numpy())
    want = np.zeros((3, 256))
    chan = np.broadcast_to(np.arange(3).reshape(1, 3, 1, 1), idx.shape)
    np.add.at(want, (chan.ravel(), idx.ravel()), g.cpu().numpy().astype(np.float64).ravel())
    assert max_abs_over_max(got, want) < TOL


@pytest.mark.parametrize("name", linear_golden_names("trainstep_"))
def test_train_step_golden(ct, name):
    """Loss, spatial means, table gradient and the Adam-updated table of every recorded step, each step started
    from the reference's own table (teacher forcing: a sign-like first Adam step amplifies 1e-8-sized gradient
    differences into +-lr, so free-running trajectories are not comparable bin by bin)."""
    from clair_torch_b200.training import linearity_loss_and_table_grad
    z = golden(name)
    val = torch.from_numpy(z["val"]).to(DEV)
    std = torch.from_numpy(z["std"]).to(DEV) if "std" in z else None
    exposure = torch.from_numpy(z["exposure"])
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(exposure, float(z["thr"]))
    coeffs = tuple(z["coeffs"])
    for step in range(int(z["n_steps"])):
        theta = z["theta0"] if step == 0 else z[f"theta_after_{step - 1}"]
        lin, spatial, grad = linearity_loss_and_table_grad(val, std, i_idx, j_idx, ratio, torch.from_numpy(theta).to(DEV),
                                                           1 / 255, 254 / 255, bool(z["rel"]), bool(z["unc"]))
        assert max_rel(lin.cpu().numpy(), z[f"linloss_{step}"]) < TOL
        assert max_rel(spatial.cpu().numpy(), z[f"spatial_{step}"]) < TOL
        pens, gpens = orc.curve_penalties(theta)
        total = grad.cpu().numpy() + sum(k * g for k, g in zip(coeffs, gpens))
        assert max_abs_over_max(total, z[f"grad_theta_{step}"]) < TOL


def test_train_icrf_step_updates_like_reference(ct):
    """train_icrf_step through the model / optimiser objects: first step moves nothing, afterwards
    the table follows Adam on the reference's gradient."""
    z = golden("trainstep_script")
    from clair_torch_b200 import ICRFModelDirect, train_icrf_step
    model = ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3) for c in range(3)]
    val, std = torch.from_numpy(z["val"]).to(DEV), torch.from_numpy(z["std"]).to(DEV)
    exposure = torch.from_numpy(z["exposure"])
    before = model.icrf.detach().clone()
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=False, alpha=10.0, beta=1.0, gamma=1.0,
              delta=1.0, exposure_ratio_threshold=0.25)
    loss0 = train_icrf_step(model, opts, val, std, exposure, **kw)
    assert torch.equal(model.icrf.detach(), before) and model.icrf.requires_grad
    # now reproduce the recorded first connected step: load the recorded start table into the parameters
    with torch.no_grad():
        for c, p in enumerate(model.direct_params):
            p.copy_(torch.from_numpy(z["theta0"][c]))
    model.update_icrf()
    loss1 = train_icrf_step(model, opts, val, std, exposure, **kw)
    assert max_rel(loss1.cpu().numpy(), z["loss_0"]) < TOL
    # Adam's first step is lr * g / (|g| + 1e-8): compare only bins whose gradient is far above the 1e-8 knee
    g = z["grad_theta_0"]
    strong = np.abs(g) > 1e-4
    diff = np.abs(model.icrf.detach().cpu().numpy() - z["theta_after_0"])
    assert diff[strong].max() < 5e-7
    assert loss0.shape == (3,)


@pytest.mark.parametrize("unc", [False, True])
def test_graphed_train_step_equals_eager_steps(ct, unc):
    """GraphedTrainStep (one captured CUDA graph, replayed) walks the same trajectory as eager train_icrf_step calls with
    the same capturable optimisers: losses and tables agree step by step (the float64 atomics of the statistics kernel
    make the last bits run-to-run noise, hence a tolerance instead of equality)."""
    val, std, t = ct.synthetic.make_stack(6, 3, 96, 128, bits=8, seed=77, device=DEV)
    exposure = torch.from_numpy(t)
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=unc, alpha=10.0, beta=1.0, gamma=1.0, delta=1.0,
              exposure_ratio_threshold=0.2)

    def fresh():
        model = ct.ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
        opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3, capturable=True) for c in range(3)]
        return model, opts

    m_e, o_e = fresh()
    eager = [ct.train_icrf_step(m_e, o_e, val, std, exposure, **kw).cpu().numpy() for _ in range(7)]
    m_g, o_g = fresh()
    losses = [ct.train_icrf_step(m_g, o_g, val, std, exposure, **kw).cpu().numpy() for _ in range(2)]
    before = ct._native.launch_count()
    step = ct.GraphedTrainStep(m_g, o_g, val, std, exposure, **kw)
    assert ct._native.launch_count() - before >= 4          # the fused kernels were captured, not executed eagerly
    losses += [step().cpu().numpy() for _ in range(5)]
    # Adam's first steps move every table entry by about +-lr whatever the size of its gradient, so an entry whose gradient
    # is fp32 reduction-order noise can go either way from run to run; the loss sees that in its 6th digit
    for a, b in zip(eager, losses):
        assert max_rel(b, a) < 1e-5
    assert max_abs_over_max(m_g.icrf.detach().cpu().numpy(), m_e.icrf.detach().cpu().numpy()) < 1e-2
    assert m_g.icrf.requires_grad
    # a plain (host-stepped) Adam cannot be captured
    plain = [torch.optim.Adam(m_g.channel_params(c), lr=1e-3) for c in range(3)]
    with pytest.raises(ValueError):
        ct.GraphedTrainStep(m_g, plain, val, std, exposure, **kw)


def test_train_icrf_replays_a_graph_for_device_resident_batches(ct):
    """train_icrf with its default (capturable) optimisers and a dataset that lives on the device: from the third epoch
    on the step is a graph replay; the result matches the eager loop."""
    from torch.utils.data import DataLoader
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    val, std, t = ct.synthetic.make_stack(6, 3, 64, 96, bits=8, seed=78, device=DEV)

    class OneBatch(torch.utils.data.Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return torch.arange(6), val, std, {"exposure_time": torch.from_numpy(t)}

    loader = DataLoader(OneBatch(), batch_size=None, shuffle=False)
    out = {}
    for graphed in (False, True):
        torch.manual_seed(0)
        model = ct.ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
        ct.train_icrf(loader, 6, DEV, model, use_uncertainty_weighting=False, epochs=12, verbose=False, use_cuda_graph=graphed)
        out[graphed] = model.icrf.detach().cpu().numpy()
    assert max_abs_over_max(out[True], out[False]) < 1e-6
    assert np.abs(out[True] - np.linspace(0, 1, 256) ** 2.5).max() > 1e-4      # it did train


def test_train_icrf_on_one_exposure_pair_takes_the_single_pass_step_in_a_graph(ct):
    """Batches of two frames (one exposure pair, BASELINE config 5) without uncertainty weighting: train_icrf's step is the
    single-pass kernel + combine, eager and captured; both loops end at the same table, and the loss of a step equals the
    two-pass formulation's."""
    from torch.utils.data import DataLoader
    val, std, _ = ct.synthetic.make_stack(2, 3, 96, 128, bits=8, seed=79, device=DEV)
    t = np.array([0.01, 0.02])

    class OneBatch(torch.utils.data.Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return torch.arange(2), val, std, {"exposure_time": torch.from_numpy(t)}

    loader = DataLoader(OneBatch(), batch_size=None, shuffle=False)
    out = {}
    for graphed in (False, True):
        model = ct.ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
        before = ct._native.launch_count()
        ct.train_icrf(loader, 2, DEV, model, use_uncertainty_weighting=False, epochs=10, verbose=False, use_cuda_graph=graphed,
                      exposure_ratio_threshold=0.25)
        out[graphed] = model.icrf.detach().cpu().numpy()
        launches = ct._native.launch_count() - before
        # eager: the first step only connects the table (means pass, upstream, penalties), the other nine are 4 library
        # launches each (single pass, its finalize, combine, penalties) instead of the two-pass step's 6
        assert (launches == 39 if not graphed else launches == 11), launches     # graphed: 3 + 4 eager, 4 at capture, replays are not counted
    assert max_abs_over_max(out[True], out[False]) < 1e-5
    assert np.abs(out[True] - np.linspace(0, 1, 256) ** 2.5).max() > 1e-4      # it did train


def test_train_icrf_recaptures_when_a_scheduler_changes_the_learning_rate(ct):
    """A captured optimiser step bakes the learning rate into its kernels: when a scheduler changes it, train_icrf must
    capture a new step (the batch key includes the optimisers' hyper-parameters) — same trajectory as the eager loop."""
    from torch.utils.data import DataLoader
    val, std, t = ct.synthetic.make_stack(5, 3, 64, 96, bits=8, seed=79, device=DEV)

    class OneBatch(torch.utils.data.Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return torch.arange(5), val, std, {"exposure_time": torch.from_numpy(t)}

    class Halve:                       # ReduceLROnPlateau's calling convention: step(metric)
        def __init__(self, opt):
            self.opt, self.calls = opt, 0

        def step(self, metric):
            self.calls += 1
            if self.calls % 4 == 0:
                for pg in self.opt.param_groups:
                    pg["lr"] = pg["lr"] * 0.5

    loader = DataLoader(OneBatch(), batch_size=None, shuffle=False)
    out = {}
    for graphed in (False, True):
        model = ct.ICRFModelDirect(256, 3, initial_power=2.5).to(DEV)
        opts = [torch.optim.Adam(model.channel_params(c), lr=2e-3, capturable=True) for c in range(3)]
        scheds = [Halve(o) for o in opts]
        ct.train_icrf(loader, 5, DEV, model, opts, scheds, use_uncertainty_weighting=False, epochs=14, verbose=False,
                      use_cuda_graph=graphed)
        out[graphed] = model.icrf.detach().cpu().numpy()
        assert opts[0].param_groups[0]["lr"] == pytest.approx(2e-3 / 8)
    assert max_abs_over_max(out[True], out[False]) < 1e-6


@pytest.mark.parametrize("staged", [True, False])
def test_hdr_merge_pinned_host_stack(ct, staged):
    """Pinned host batches are either streamed band by band by the copy engine while the kernel merges the previous band
    (staged) or read by the kernel over PCIe (zero-copy), and results can be written straight to pinned host buffers:
    bit-identical to the device-resident path, for one batch and for several."""
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    val, std, t = ct.synthetic.make_stack(5, 3, 90, 161, seed=17)
    theta = ct.synthetic.reference_curve(3)
    model = _model(ct, theta.numpy())
    z = {"val": val.numpy(), "std": std.numpy(), "exposure": t}
    ref_rad, ref_sig = ct.compute_hdr_image(_loader(ct, z, 5), DEV, model, max, radiance_dtype=torch.float32)
    ds = ExposureStackDataset(list(val), list(std), list(t))
    for bs in (5, 2):
        loader = DataLoader(ds, batch_size=bs, shuffle=False, collate_fn=custom_collate, pin_memory=True)
        rad_h = torch.empty((3, 90, 161), dtype=torch.float32).pin_memory()
        sig_h = torch.empty((3, 90, 161), dtype=torch.float32).pin_memory()
        rad, sig = ct.compute_hdr_image(loader, DEV, model, max, radiance_dtype=torch.float32, host_out=(rad_h, sig_h),
                                        staged=staged)
        torch.cuda.synchronize()
        assert rad.data_ptr() == rad_h.data_ptr() and not rad.is_cuda
        if bs == 5:
            assert torch.equal(rad, ref_rad.cpu()) and torch.equal(sig, ref_sig.cpu())
        else:
            o_rad, o_sig = orc.hdr_merge(z["val"], z["std"], t, theta.numpy(), True, bs)
            assert max_rel(rad.numpy(), o_rad) < 2e-6 and max_rel(sig.numpy(), o_sig) < 5e-6
    with pytest.raises(ValueError):
        ct.compute_hdr_image(loader, DEV, model, max, radiance_dtype=torch.float32,
                             host_out=(torch.empty((3, 90, 161)), torch.empty((3, 90, 161))))


@pytest.mark.parametrize("shape,bands", [((4, 3, 7, 11), 8), ((5, 3, 64, 48), 3), ((9, 3, 40, 52), 64), ((3, 1, 33, 31), 2)])
def test_hdr_merge_staged_bands_equal_one_launch(ct, shape, bands):
    """The band split of the staged entry point (any band count, planes smaller than one band, N in the register / parked
    kernels, float64 radiance) is invisible in the result; uint8 codes and a LOOKUP model take the same route."""
    n, c, h, w = shape
    val, std, t = ct.synthetic.make_stack(n, c, h, w, seed=5 + n)
    theta = ct.synthetic.reference_curve(c).to(DEV)
    whole = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), val.to(DEV), std.to(DEV), t, theta, True, True)
    banded = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), val.pin_memory(), std.pin_memory(), t, theta, True, True,
                                         device=torch.device(DEV), staged=True, bands=bands)
    assert torch.equal(whole[0], banded[0]) and torch.equal(whole[1], banded[1])
    for mode in (ct._native.INTERP_LOOKUP, ct._native.INTERP_CATMULL):
        a = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), val.to(DEV), std.to(DEV), t, theta, True, True, interp_mode=mode)
        b = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), val.pin_memory(), std.pin_memory(), t, theta, True, True,
                                        device=torch.device(DEV), staged=True, bands=bands, interp_mode=mode)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    if (h * w) % 4 == 0:
        codes = torch.round(val * 255).to(torch.uint8)
        from clair_torch_b200.datasets import StdSpec
        spec = StdSpec("multiplier", 0.05)
        a = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), codes.to(DEV), spec, t, theta, True, True)
        b = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), codes.pin_memory(), spec, t, theta, True, True,
                                        device=torch.device(DEV), staged=True, bands=bands)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("bits", [8, 16])
def test_hdr_merge_integer_ingest_is_bit_identical(ct, bits):
    """Raw uint8 / uint16 codes + in-kernel CastTo/Normalize/std synthesis give exactly what the CPU-transformed fp32
    images give (fl32(code)/fl32(max) with an IEEE division), for every std mode, one batch and two."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    n = 5 if bits == 8 else 9          # 9 frames exercises the N > 8 kernel
    val, _, t = ct.synthetic.make_stack(n, 3, 48, 64, bits=bits, seed=bits)
    maxval = float(2 ** bits - 1)
    codes = torch.round(val * maxval).to(torch.uint8 if bits == 8 else torch.uint16)
    assert torch.equal(codes.to(torch.float32) / maxval, val)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    mult = torch.tensor(0.05)
    cases = [("multiplier", StdSpec("multiplier", 0.05), val * mult), ("constant", StdSpec("constant", 0.01), torch.full_like(val, 0.01)),
             ("tensor", (val * 0.03 + 0.001), (val * 0.03 + 0.001)), ("none", None, None)]
    for name, std_codes, std_f32 in cases:
        for split in (None, 3):
            def run(v, s, **kw):
                st = kernels.HdrMergeState()
                bounds = [(0, n)] if split is None else [(0, split), (split, n)]
                out = None
                for a, b in bounds:
                    sb = s[a:b].contiguous().to(DEV) if torch.is_tensor(s) else s
                    out = kernels.hdr_merge_update(st, v[a:b].contiguous().to(DEV), sb, t[a:b], theta, True, b == n,
                                                   radiance_dtype=torch.float32, **kw)
                return out
            got = run(codes, std_codes)
            want = run(val, std_f32)
            assert torch.equal(got[0], want[0]), (name, split)
            if want[1] is None:
                assert got[1] is None
            else:
                assert torch.equal(got[1], want[1]), (name, split)


@pytest.mark.parametrize("bits,n", [(8, 5), (16, 4), (8, 9), (16, 12)])
def test_hdr_merge_interleaved_bgr_codes_equal_planar(ct, bits, n):
    """code_layout='hwc_bgr': (N, H, W, 3) BGR camera buffers give the same bits as the planar RGB codes the reference's
    CvToTorch would have produced from them — register, parked and all-modes kernels, device and pinned host memory."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    val, _, t = ct.synthetic.make_stack(n, 3, 30, 44, bits=bits, seed=bits + n)
    maxval = float(2 ** bits - 1)
    planar = torch.round(val * maxval).to(torch.uint8 if bits == 8 else torch.uint16)
    # what cv2.imread hands over: height x width x (B, G, R)
    camera = torch.stack([planar[:, 2], planar[:, 1], planar[:, 0]], dim=-1).contiguous()
    assert camera.shape == (n, 30, 44, 3)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    std_tensor = (val * 0.03 + 0.001)
    for std_p in (StdSpec("multiplier", 0.05), StdSpec("constant", 0.02), std_tensor.to(DEV), None):
        for mode in (ct._native.INTERP_LINEAR, ct._native.INTERP_LOOKUP):
            want = kernels.hdr_merge_update(kernels.HdrMergeState(), planar.to(DEV), std_p, t, theta, True, True, interp_mode=mode)
            got = kernels.hdr_merge_update(kernels.HdrMergeState(), camera.to(DEV), std_p, t, theta, True, True, interp_mode=mode,
                                           code_layout="hwc_bgr")
            assert torch.equal(got[0], want[0])
            assert (got[1] is None and want[1] is None) or torch.equal(got[1], want[1])
    dev = kernels.hdr_merge_update(kernels.HdrMergeState(), planar.to(DEV), StdSpec("multiplier", 0.05), t, theta, True, True)
    for staged, bands in ((None, 16), (True, 3), (False, 1)):      # default = staged band pipeline; False = read in place
        host = kernels.hdr_merge_update(kernels.HdrMergeState(), camera.pin_memory(), StdSpec("multiplier", 0.05), t, theta, True, True,
                                        device=torch.device(DEV), code_layout="hwc_bgr", staged=staged, bands=bands)
        assert torch.equal(host[0], dev[0]) and torch.equal(host[1], dev[1])
    with pytest.raises(ValueError):
        kernels.hdr_merge_update(kernels.HdrMergeState(), val.to(DEV), None, t, theta, True, True, code_layout="hwc_bgr")


@pytest.mark.parametrize("bits,n,shape", [(16, 9, (36, 48)), (8, 12, (36, 48)), (16, 16, (40, 64)), (16, 9, (1088, 1936)), (8, 10, (544, 1936))])
def test_hdr_merge_camera_codes_through_the_bulk_copy_stage(ct, bits, n, shape):
    """9..16 frames of (N, H, W, 3) camera codes with H*W % 16 == 0 take the register kernel whose codes arrive in shared
    memory by cp.async.bulk (two-stage ring, one trip ahead): bit-identical to the planar codes and to the per-thread-load
    form of the same kernel — single trips with a partial tail, and many trips per block (one resident wave)."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    h, w = shape
    assert (h * w) % 16 == 0
    val, _, t = ct.synthetic.make_stack(n, 3, h, w, bits=bits, seed=3 * bits + n, device=DEV)
    maxval = float(2 ** bits - 1)
    planar = torch.round(val * maxval).to(torch.int32).to(torch.uint8 if bits == 8 else torch.uint16)
    camera = torch.stack([planar[:, 2], planar[:, 1], planar[:, 0]], dim=-1).contiguous()
    theta = ct.synthetic.reference_curve(3).to(DEV)
    lib = ct._native.load()
    spec = StdSpec("multiplier", 0.05)
    want = kernels.hdr_merge_update(kernels.HdrMergeState(), planar, spec, t, theta, True, True)
    try:
        for waves in (0, 1):
            ct._native.check(lib.clair_set_tuning(b"hdr_waves", waves), "tune")
            got = kernels.hdr_merge_update(kernels.HdrMergeState(), camera, spec, t, theta, True, True, code_layout="hwc_bgr")
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        ct._native.check(lib.clair_set_tuning(b"hdr_tma", -1), "tune")
        plain = kernels.hdr_merge_update(kernels.HdrMergeState(), camera, spec, t, theta, True, True, code_layout="hwc_bgr")
        assert torch.equal(plain[0], want[0]) and torch.equal(plain[1], want[1])
    finally:
        lib.clair_set_tuning(b"hdr_waves", 0)
        lib.clair_set_tuning(b"hdr_tma", 0)
    # std as a tensor, two batches, and the staged host pipeline (bands of larger planes)
    std = (val * 0.03 + 0.001)
    st_a, st_b = kernels.HdrMergeState(), kernels.HdrMergeState()
    half = n // 2
    if n - half >= 9:
        for sl, last in ((slice(0, half), False), (slice(half, n), True)):
            a = kernels.hdr_merge_update(st_a, planar[sl].contiguous(), std[sl].contiguous(), t[sl], theta, True, last)
            b = kernels.hdr_merge_update(st_b, camera[sl].contiguous(), std[sl].contiguous(), t[sl], theta, True, last, code_layout="hwc_bgr")
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    host = kernels.hdr_merge_update(kernels.HdrMergeState(), camera.cpu().pin_memory(), spec, t, theta, True, True, device=torch.device(DEV),
                                    code_layout="hwc_bgr", staged=True, bands=3)
    assert torch.equal(host[0], want[0]) and torch.equal(host[1], want[1])


@pytest.mark.parametrize("n", [9, 10, 11, 13, 14, 15, 16])
def test_hdr_merge_integer_ingest_9_to_16_frames(ct, n):
    """9..16 frames of uint8 / uint16 codes (planar and BGR camera layout) give the bits of the fp32 stack, which takes the
    fp32 register kernel, and stay within tolerance of the oracle — whichever kernel the frame count selects for codes
    (the 2-code register kernel or the shared-memory-parked one), one batch and two."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    theta = ct.synthetic.reference_curve(3).to(DEV)
    for bits in (16, 8):
        val, _, t = ct.synthetic.make_stack(n, 3, 72, 100, bits=bits, seed=100 * bits + n)
        maxval = float(2 ** bits - 1)
        codes = torch.round(val * maxval).to(torch.uint8 if bits == 8 else torch.uint16)
        assert torch.equal(codes.to(torch.float32) / maxval, val)
        camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous()
        std_tensor = val * 0.03 + 0.001
        for std_codes, std_f32 in ((StdSpec("multiplier", 0.05), val * torch.tensor(0.05)), (std_tensor, std_tensor)):
            for split in (None, 7):
                def run(v, s, **kw):
                    st = kernels.HdrMergeState()
                    out = None
                    for a, b in ([(0, n)] if split is None else [(0, split), (split, n)]):
                        sb = s[a:b].contiguous().to(DEV) if torch.is_tensor(s) else s
                        out = kernels.hdr_merge_update(st, v[a:b].contiguous().to(DEV), sb, t[a:b], theta, True, b == n,
                                                       radiance_dtype=torch.float32, **kw)
                    return out
                want = run(val, std_f32)
                if split is None:
                    single = want
                for name, got in (("planar", run(codes, std_codes)), ("camera", run(camera, std_codes, code_layout="hwc_bgr"))):
                    assert torch.equal(got[0], want[0]), (bits, name, split)
                    assert torch.equal(got[1], want[1]), (bits, name, split)
        o_rad, o_sig = orc.hdr_merge(val.numpy(), std_tensor.numpy(), np.asarray(t, dtype=np.float64), theta.cpu().numpy())
        assert max_rel(single[0].cpu().numpy(), o_rad) < TOL      # the one-batch merge with std as a tensor
        assert max_rel(single[1].cpu().numpy(), o_sig) < TOL


@pytest.mark.parametrize("code_max", [65535.0, 4095.0, 1023.0, 60000.0, 65536.0, 1e5])
def test_hdr_merge_every_16_bit_code_normalises_like_the_ieee_division(ct, code_max):
    """The register kernels divide 16-bit codes by code_max as q0 = a*rcp corrected by one exact residual
    (normalise_code16, clair_merge.cuh); the reference's Normalize is an IEEE division.  Every one of the 65 536 codes, in
    every frame (each frame a different permutation of them; 0..code_max for the narrower ranges), through the 4-code (3, 8 frames) and 2-code (9, 16 frames)
    register kernels, planar and camera layout: the bits of the CPU-divided fp32 stack."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    theta = ct.synthetic.reference_curve(3).to(DEV)
    idx = torch.arange(65536, dtype=torch.int64)
    for n in (3, 8, 9, 16):
        n_codes = min(65536, int(code_max) + 1)                  # a 12-bit camera delivers 0..4095
        frames = [(((idx * (2 * k + 1) + 977 * c) % 65536) % n_codes).reshape(256, 256) for k in range(n) for c in range(3)]
        codes = torch.stack(frames).reshape(n, 3, 256, 256).to(torch.int32).to(torch.uint16)
        val = codes.to(torch.float32) / code_max                 # CastTo + Normalize on the CPU: a true division
        t = ct.synthetic.exposure_times(n, 1e-3)
        camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous()
        want = kernels.hdr_merge_update(kernels.HdrMergeState(), val.to(DEV), (val * torch.tensor(0.05)).to(DEV), t, theta, True, True,
                                        radiance_dtype=torch.float32)
        for buf, kw in ((codes, {}), (camera, {"code_layout": "hwc_bgr"})):
            got = kernels.hdr_merge_update(kernels.HdrMergeState(), buf.to(DEV), StdSpec("multiplier", 0.05), t, theta, True, True,
                                           radiance_dtype=torch.float32, code_max=code_max, **kw)
            for g, w in zip(got, want):
                assert torch.equal(torch.nan_to_num(g, nan=-1.0), torch.nan_to_num(w, nan=-1.0)), (n, kw)


@pytest.mark.parametrize("name", golden_names("ingest_"))
def test_hdr_merge_camera_codes_against_reference_transform_fixture(ct, name):
    """The camera buffers of the ingest fixtures (every code of the range) handed over untouched — code_layout hwc_bgr, code_max,
    StdSpec — merge to the bits of the frames the reference's CvToTorch + CastTo + Normalize made of them, with the std images
    its datasets synthesise."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import StdSpec
    z = golden(name)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    shifts = (0, 5, 11, 2, 7, 13, 3, 9, 1, 6)                   # frame k = the fixture's frame rolled along the width
    for n in (3, 10):
        t = ct.synthetic.exposure_times(n, 1e-3)
        camera = torch.from_numpy(np.stack([np.roll(z["camera"], s, axis=1) for s in shifts[:n]]).astype(np.int32))
        camera = camera.to(torch.uint8 if z["camera"].dtype == np.uint8 else torch.uint16)

        def frames(key):
            return torch.from_numpy(np.stack([np.roll(z[key], s, axis=2) for s in shifts[:n]]))
        for spec, std_key in ((StdSpec("multiplier", float(z["multiplier"])), "std_multiplier"),
                              (StdSpec("constant", float(z["constant"])), "std_constant")):
            want = kernels.hdr_merge_update(kernels.HdrMergeState(), frames("val").to(DEV), frames(std_key).to(DEV), t, theta, True, True,
                                            radiance_dtype=torch.float32)
            got = kernels.hdr_merge_update(kernels.HdrMergeState(), camera.to(DEV), spec, t, theta, True, True,
                                           radiance_dtype=torch.float32, code_max=float(z["max_val"]), code_layout="hwc_bgr")
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), (n, std_key)


def test_hdr_merge_integer_ingest_against_reference_fixture(ct):
    """The 16-bit reference fixture fed as uint16 codes through the public API (DataLoader + StdSpec)."""
    from clair_torch_b200.datasets import ExposureStackDataset, StdSpec, custom_collate
    z = golden("hdr_u16")
    codes = torch.round(torch.from_numpy(z["val"]) * 65535.0).to(torch.uint16)
    assert np.array_equal((codes.to(torch.float32) / 65535.0).numpy(), z["val"])
    ds = ExposureStackDataset(list(codes), StdSpec("multiplier", 0.05), list(z["exposure"]))
    loader = DataLoader(ds, batch_size=len(codes), collate_fn=custom_collate)
    rad, sig = ct.compute_hdr_image(loader, DEV, _model(ct, z["theta"]), max)
    assert rad.dtype == torch.float64
    assert max_rel(rad.cpu().numpy(), z["radiance"]) < TOL
    assert max_rel(sig.cpu().numpy(), z["sigma"]) < TOL


@pytest.mark.parametrize("bits", [8, 16])
def test_linearize_integer_ingest_is_bit_identical(ct, bits):
    """clair_linearize_codes: raw uint8 / uint16 codes + in-kernel CastTo / Normalize / std synthesis give exactly what the
    CPU-transformed fp32 images give through clair_linearize — std as a tensor, as value * m, as a constant, and absent;
    device-resident and pinned-host codes; through kernels.linearize and through linearize_dataset_generator."""
    from clair_torch_b200.datasets import ExposureStackDataset, StdSpec
    maxval = 255.0 if bits == 8 else 65535.0
    val, std, t = ct.synthetic.make_stack(3, 3, 46, 64, bits=bits, seed=31 + bits)
    codes = torch.round(val * maxval).to(torch.uint8 if bits == 8 else torch.uint16)
    x = codes.to(torch.float32) / maxval                                    # CastTo + Normalize on the CPU
    assert torch.equal(x, val)
    theta = ct.synthetic.reference_curve(3).to(DEV)
    m = float(np.float32(0.05))
    cases = [(std, std), (StdSpec("multiplier", 0.05), x * m), (StdSpec("constant", 0.01), torch.full_like(x, 0.01)), (None, None)]
    for spec, std_f32 in cases:
        ref_lin, ref_sig = ct.kernels.linearize(x.to(DEV), None if std_f32 is None else std_f32.to(DEV), theta)
        s_dev = spec.to(DEV) if torch.is_tensor(spec) else spec
        lin, sig = ct.kernels.linearize(codes.to(DEV), s_dev, theta)
        assert torch.equal(lin, ref_lin) and torch.equal(sig, ref_sig)
        s_pin = spec.pin_memory() if torch.is_tensor(spec) else spec
        lin_h, sig_h = ct.kernels.linearize(codes.pin_memory(), s_pin, theta, device=torch.device(DEV), pinned_out=True)
        torch.cuda.synchronize()
        assert not lin_h.is_cuda and torch.equal(lin_h, ref_lin.cpu()) and torch.equal(sig_h, ref_sig.cpu())
    # a 12-bit camera in 16-bit words: code_max = 4095
    if bits == 16:
        c12 = (codes.to(torch.int32) >> 4).to(torch.uint16)
        x12 = c12.to(torch.float32) / 4095.0
        a = ct.kernels.linearize(c12.to(DEV), StdSpec("multiplier", 0.05), theta, code_max=4095.0)
        b = ct.kernels.linearize(x12.to(DEV), (x12 * m).to(DEV), theta)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # the generator: one image per batch, codes in pinned host memory, results on the host
    model = _model(ct, theta.cpu().numpy())
    ds = ExposureStackDataset(list(codes.pin_memory()), StdSpec("multiplier", 0.05), list(t))
    view = lambda b: (torch.tensor([b[0][0]]), b[0][1].unsqueeze(0), b[0][2], {"exposure_time": torch.tensor([b[0][3]["exposure_time"]])})
    ref_lin, ref_sig = ct.kernels.linearize(x.to(DEV), (x * m).to(DEV), theta)
    for k, (lin, sig, meta) in enumerate(ct.linearize_dataset_generator(DataLoader(ds, batch_size=1, collate_fn=view), DEV, model)):
        assert torch.equal(lin, ref_lin[k].cpu()) and torch.equal(sig, ref_sig[k].cpu())
    # LOOKUP / CATMULL models take the same integer-ingest kernel
    for mode, spec, std_f32 in ((ct._native.INTERP_CATMULL, StdSpec("multiplier", 0.05), x * m), (ct._native.INTERP_CATMULL, None, None),
                                (ct._native.INTERP_LOOKUP, None, None)):
        a = ct.kernels.linearize(codes.to(DEV), spec, theta, interp_mode=mode)
        b = ct.kernels.linearize(x.to(DEV), None if std_f32 is None else std_f32.to(DEV), theta, interp_mode=mode)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # errors: StdSpec without codes, odd plane, std images through a LOOKUP model (as the reference's autograd call)
    with pytest.raises(ValueError):
        ct.kernels.linearize(x.to(DEV), StdSpec("multiplier", 0.05), theta)
    with pytest.raises(ValueError):
        ct.kernels.linearize(codes[:, :, :3, :5].contiguous().to(DEV), None, theta)
    with pytest.raises(RuntimeError):
        ct.kernels.linearize(codes.to(DEV), StdSpec("multiplier", 0.05), theta, interp_mode=ct._native.INTERP_LOOKUP)


@pytest.mark.parametrize("bits", [8, 16])
def test_pair_drivers_from_integer_codes_are_bit_identical(ct, bits):
    """measure_linearity and a train_icrf step fed with raw uint8 / uint16 codes + StdSpec (normalised on the device by
    clair_expand_codes) equal the same calls on the CPU-transformed fp32 images; every code of the 8-bit range and a sweep
    of the 16-bit range normalises like the IEEE division."""
    from clair_torch_b200 import kernels
    from clair_torch_b200.datasets import ExposureStackDataset, StdSpec, custom_collate
    maxval = 255.0 if bits == 8 else 65535.0
    dt = torch.uint8 if bits == 8 else torch.uint16
    val, _, t = ct.synthetic.make_stack(6, 3, 40, 64, bits=bits, seed=21 + bits, std_multiplier=None)
    codes = torch.round(val * maxval).to(torch.int32).to(dt)
    x = codes.to(torch.float32) / maxval
    m = float(np.float32(0.05))
    xv, xs = kernels.expand_codes(codes.pin_memory(), StdSpec("multiplier", 0.05), device=DEV)
    assert torch.equal(xv.cpu(), x) and torch.equal(xs.cpu(), x * m)
    every = torch.arange(0, 256 if bits == 8 else 65536, dtype=torch.int32).to(dt)
    ev, es = kernels.expand_codes(every.to(DEV), StdSpec("constant", 0.01))
    assert torch.equal(ev.cpu(), every.to(torch.float32) / maxval) and torch.equal(es.cpu(), torch.full((every.numel(),), 0.01))
    c12 = kernels.expand_codes(every.to(DEV), None, code_max=4095.0)[0]
    assert torch.equal(c12.cpu(), every.to(torch.float32) / 4095.0)
    model = _model(ct, ct.synthetic.reference_curve(3).numpy())
    ds_codes = ExposureStackDataset(list(codes), StdSpec("multiplier", 0.05), list(t))
    ds_f32 = ExposureStackDataset(list(x), list(x * m), list(t))
    got = ct.measure_linearity(DataLoader(ds_codes, batch_size=6, collate_fn=custom_collate), DEV, True, True, model)
    want = ct.measure_linearity(DataLoader(ds_f32, batch_size=6, collate_fn=custom_collate), DEV, True, True, model)
    assert torch.equal(got[0], want[0])
    for a, b in zip(got[1:], want[1:]):
        assert max_rel(a.cpu().numpy(), b.cpu().numpy()) < 1e-12          # same kernels on the same bits; float64 atomics order
    losses = []
    for ds in (ds_codes, ds_f32):
        mdl = ct.ICRFModelDirect(256, 3, initial_power=2.2).to(DEV)
        ct.train_icrf(DataLoader(ds, batch_size=6, collate_fn=custom_collate), 6, DEV, mdl, epochs=2, verbose=False, use_cuda_graph=False)
        losses.append(mdl.icrf.detach().cpu().numpy())
    assert max_rel(losses[0], losses[1], 1e-6) < 1e-5


def test_measure_linearity_pipelines_large_pinned_code_batches(ct):
    """A page-locked code batch of more than 64 MB is cut into row bands whose copies overlap the expansion and the statistics of
    the bands already on the device: same result as the whole batch (summation order aside), for a band count that divides
    the height and with the fallback when none does."""
    from clair_torch_b200.datasets import StdSpec
    import importlib
    ml = importlib.import_module("clair_torch_b200.inference.measure_linearity")     # (the package re-exports the function)
    model = _model(ct, ct.synthetic.reference_curve(3).numpy())
    for h, w, expect in ((1024, 1408, 8), (1021, 1416, 0)):
        val, _, t = ct.synthetic.make_stack(8, 3, h, w, bits=16, seed=h, std_multiplier=None)
        codes = torch.round(val * 65535.0).to(torch.int32).to(torch.uint16)
        del val
        spec = StdSpec("multiplier", 0.05)
        exposures = {"exposure_time": torch.as_tensor(np.asarray(t, dtype=np.float64))}
        assert ml._code_bands(codes.pin_memory(), spec) == expect and ml._code_bands(codes, spec) == 0
        batch = lambda v: DataLoader([0], batch_size=1, collate_fn=lambda _: (torch.arange(8), v, spec, exposures))
        got = ct.measure_linearity(batch(codes.pin_memory()), DEV, True, True, model)
        want = ct.measure_linearity(batch(codes.to(DEV)), DEV, True, True, model)
        assert torch.equal(got[0], want[0])
        for a, b in zip(got[1:], want[1:]):
            assert max_rel(a.cpu().numpy(), b.cpu().numpy()) < 2e-6


# ---- streaming frame statistics -------------------------------------------------
def test_wbomeanvar_golden(ct):
    from clair_torch_b200.common.statistics import WBOMeanVar
    from clair_torch_b200.common.enums import VarianceMode
    z = golden("framestats_wbomeanvar")
    val, wts = torch.from_numpy(z["val"]).to(DEV), torch.from_numpy(z["weights"]).to(DEV)
    b = z["bounds"]
    for tag, weighted in (("w", True), ("u", False)):
        h = WBOMeanVar(dim=0, variance_mode=VarianceMode.RELIABILITY_WEIGHTS)
        for a, e in zip(b[:-1], b[1:]):
            mean, m2 = h.update_values(val[a:e].contiguous(), wts[a:e].contiguous() if weighted else None)
        assert tuple(mean.shape) == z[f"mean_{tag}"].shape
        assert max_rel(mean.cpu().numpy(), z[f"mean_{tag}"]) < TOL
        assert max_rel(m2.cpu().numpy(), z[f"m2_{tag}"]) < 3e-5        # the reference's own fp32 two-pass sum
        assert max_rel(h.variance().cpu().numpy(), z[f"var_rel_{tag}"]) < 3e-5
        o_mean, o_m2, _, _ = orc.frame_stats(z["val"], z["weights"] if weighted else None, b)
        assert max_rel(mean.cpu().numpy()[0], o_mean) < 2e-6 and max_rel(m2.cpu().numpy()[0], o_m2) < TOL


@pytest.mark.parametrize("name", golden_names("framestats_video"))
def test_compute_video_mean_and_std_golden(ct, name):
    z = golden(name)
    model = _model(ct, z["theta"]) if "theta" in z else None
    loader = _loader(ct, z, int(z["batch_size"]))
    mean, sem = ct.compute_video_mean_and_std(loader, DEV, model)
    assert max_rel(mean.cpu().numpy(), z["mean"]) < TOL
    assert max_rel(sem.cpu().numpy(), z["sem"]) < 3e-5


@pytest.mark.parametrize("mode", ["catmull", "lookup"])
def test_video_statistics_with_lookup_and_catmull_models(ct, mode):
    """compute_video_mean_and_std linearises inside the statistics kernel for any InterpMode: equal to running the model's
    own forward kernel first and the statistics on its output."""
    from clair_torch_b200.common.statistics import WBOMeanVar
    from clair_torch_b200.datasets import ExposureStackDataset, custom_collate
    val, _, t = ct.synthetic.make_stack(9, 3, 24, 40, bits=8, seed=3, std_multiplier=None)
    model = _model(ct, ct.synthetic.reference_curve(3).numpy(), mode)
    loader = DataLoader(ExposureStackDataset(list(val), None, list(t)), batch_size=4, shuffle=False, collate_fn=custom_collate)
    mean, sem = ct.compute_video_mean_and_std(loader, DEV, model)
    h = WBOMeanVar(dim=0, variance_mode=ct.common.VarianceMode.SAMPLE_FREQUENCY)
    for _, vb, _, _ in loader:
        h.update_values(model(vb.to(DEV)).detach(), None)
    assert torch.equal(mean, h.mean.squeeze())
    assert torch.equal(sem, torch.sqrt(h.variance().squeeze()) / 3.0)


def test_frame_stats_large_vs_oracle(ct):
    """64 frames of 540p in four uneven batches, linearised in the same pass."""
    from clair_torch_b200.common.statistics import WBOMeanVar
    g = torch.Generator().manual_seed(1)
    val = torch.rand((64, 3, 135, 240), generator=g)
    theta = ct.synthetic.reference_curve(3)
    h = WBOMeanVar(dim=0)
    bounds = [0, 7, 30, 31, 64]
    for a, e in zip(bounds[:-1], bounds[1:]):
        h.update_values(val[a:e].contiguous().to(DEV), None, table=theta.to(DEV))
    o_mean, o_m2, o_w, _ = orc.frame_stats(val.numpy(), None, bounds, theta.numpy())
    assert max_rel(h.mean.cpu().numpy()[0], o_mean) < 2e-6
    assert max_rel(h.m2.cpu().numpy()[0], o_m2) < TOL
    assert torch.equal(h.sum_of_weights.cpu(), torch.full((1, 3, 135, 240), 64.0))


# ---- dark-field / flat-field corrections -----------------------------------------
def _artefact_datasets(z):
    from clair_torch_b200.datasets import InMemoryArtefactDataset
    dark = flat = None
    if "dark" in z:
        dark = InMemoryArtefactDataset([torch.from_numpy(d) for d in z["dark"]], [torch.from_numpy(d) for d in z["dark_std"]],
                                       list(z["exposure"]))
    if "flat" in z:
        flat = InMemoryArtefactDataset([torch.from_numpy(z["flat"][0])], [torch.from_numpy(z["flat_std"][0])])
    return dark, flat


@pytest.mark.parametrize("name", golden_names("artefact_"))
def test_artefact_corrections_golden(ct, name):
    z = golden(name)
    dark, flat = _artefact_datasets(z)
    model = _model(ct, z["theta"])
    rad, sig = ct.compute_hdr_image(_loader(ct, z, int(z["batch_size"])), DEV, model, max, flat, None, dark)
    assert rad.dtype == torch.float64
    assert max_rel(rad.cpu().numpy(), z["radiance"]) < TOL
    assert max_rel(sig.cpu().numpy(), z["sigma"]) < TOL
    outs = list(ct.linearize_dataset_generator(_loader(ct, z, 1), DEV, model, flat, None, dark))
    for n, (lin, lsig, _) in enumerate(outs):
        assert max_rel(lin.numpy(), z["linearized"][n]) < TOL
        assert max_rel(lsig.numpy(), z["lin_sigma"][n]) < TOL


@pytest.mark.parametrize("shape", [(4, 3, 50, 132), (5, 3, 33, 130), (3, 1, 2, 2), (8, 3, 21, 70), (2, 3, 64, 64)])
def test_dark_field_row_groups_and_fused_merge(ct, shape):
    """Even widths take the row-group pre-pass kernel (4 or 2 pixels per thread, neighbours by lane shuffle) and, inside
    compute_hdr_image, the merge kernel with the dark-field mix fused into its load: both against the oracle, and the
    fused merge against pre-pass + merge (same arithmetic, so 1e-6), for one batch and for several."""
    from clair_torch_b200 import kernels
    n, c, h, w = shape
    rng = np.random.default_rng(n * 100 + w)
    val, std, t = ct.synthetic.make_stack(n, c, h, w, bits=16, seed=w)
    dark = rng.uniform(0, 0.02, size=val.shape).astype(np.float32)
    hot = rng.random(val.shape) < 0.1
    dark[hot] = rng.uniform(0.03, 0.4, size=int(hot.sum())).astype(np.float32)
    dark_std = (0.1 * dark + 1e-3).astype(np.float32)
    d_val, d_std = torch.from_numpy(dark).to(DEV), torch.from_numpy(dark_std).to(DEV)
    o_mixed, o_seff = orc.dark_field_mix(val.numpy(), std.numpy(), dark, dark_std)
    mixed, seff = kernels.dark_field_mix(val.to(DEV), std.to(DEV), d_val, d_std)
    assert max_rel(mixed.cpu().numpy(), o_mixed, 1e-12) < 2e-6
    assert max_rel(seff.cpu().numpy(), o_seff, 1e-12) < 5e-6
    only_val, none = kernels.dark_field_mix(val.to(DEV), None, d_val, None)
    assert none is None and torch.equal(only_val, mixed)
    theta = ct.synthetic.reference_curve(c).to(DEV)
    assert kernels.can_fuse_dark(val.to(DEV), std.to(DEV), d_std)
    for batch in (n, max(1, n // 2)):
        st_f, st_p = kernels.HdrMergeState(), kernels.HdrMergeState()
        for a in range(0, n, batch):
            sl = slice(a, min(a + batch, n))
            last = sl.stop == n
            fused = kernels.hdr_merge_update(st_f, val[sl].to(DEV), std[sl].to(DEV), t[sl], theta, True, last, dark=(d_val[sl], d_std[sl]))
            pre = kernels.hdr_merge_update(st_p, mixed[sl], seff[sl], t[sl], theta, True, last)
        assert max_rel(fused[0].cpu().numpy(), pre[0].cpu().numpy()) < 1e-6
        assert max_rel(fused[1].cpu().numpy(), pre[1].cpu().numpy()) < 5e-6     # different FMA contraction in the two kernels
        want_r, want_s = orc.hdr_merge(o_mixed, o_seff, t, theta.cpu().numpy(), True, batch)
        assert max_rel(fused[0].cpu().numpy(), want_r) < TOL and max_rel(fused[1].cpu().numpy(), want_s) < TOL


def test_dark_mix_and_flat_field_vs_oracle_larger(ct):
    from clair_torch_b200 import kernels
    rng = np.random.default_rng(9)
    val, std, t = ct.synthetic.make_stack(4, 3, 97, 131, bits=16, seed=5)
    dark = rng.uniform(0, 0.02, size=val.shape).astype(np.float32)
    hot = rng.random(val.shape) < 0.05
    dark[hot] = rng.uniform(0.03, 0.4, size=int(hot.sum())).astype(np.float32)
    dark_std = (0.1 * dark + 1e-3).astype(np.float32)
    mixed, seff = kernels.dark_field_mix(val.to(DEV), std.to(DEV), torch.from_numpy(dark).to(DEV), torch.from_numpy(dark_std).to(DEV))
    o_mixed, o_seff = orc.dark_field_mix(val.numpy(), std.numpy(), dark, dark_std)
    assert max_rel(mixed.cpu().numpy(), o_mixed, 1e-12) < 2e-6
    assert max_rel(seff.cpu().numpy(), o_seff, 1e-12) < 5e-6
    flat = rng.uniform(0.5, 1.0, size=(3, 97, 131)).astype(np.float32)
    flat_std = (0.02 * flat).astype(np.float32)
    for dtype, in_graph in ((torch.float64, True), (torch.float32, False)):
        value = torch.from_numpy(rng.uniform(0.1, 5.0, size=(3, 97, 131))).to(dtype).to(DEV)
        sigma = torch.from_numpy(rng.uniform(0.01, 0.1, size=(3, 97, 131)).astype(np.float32)).to(DEV)
        want_v, want_var = orc.flat_field_correct(value.cpu().numpy(), sigma.cpu().numpy().astype(np.float64) ** 2, flat, flat_std, in_graph)
        kernels.flat_field_correct_(value, sigma, torch.from_numpy(flat), torch.from_numpy(flat_std), in_graph)
        # float64 values are corrected in float64, fp32 ones in fp32 (1-ulp reciprocal, a few roundings per element)
        assert max_rel(value.cpu().numpy(), want_v) < (1e-12 if dtype == torch.float64 else 5e-7)
        assert max_rel(sigma.cpu().numpy(), np.sqrt(want_var)) < (2e-7 if dtype == torch.float64 else 6e-7)
