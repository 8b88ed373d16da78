"""Pins the CPU oracle (oracle/clair_oracle.py) to the reference.

Every fixture under tests/golden/ was written by tests/golden/make_golden.py from the UNMODIFIED
reference; this file checks the oracle's closed forms against them.  Bit-exact where the result is an
index, a mask or a pure fp32 table evaluation; 1e-5 relative (the north-star tolerance) elsewhere, with
the tighter bound the oracle actually meets asserted next to it.
"""
import numpy as np
import pytest

from _helpers import SIGMA_TOL, golden, golden_names, linear_golden_names, max_abs_over_max, max_rel, mode_of
from oracle import clair_oracle as orc

TOL = 1e-5   # BASELINE.json north_star: radiance, uncertainty and loss within 1e-5 relative


def _opt(z, k):
    return z[k] if k in z else None


# ---- known-answer vectors held by the reference's own unit tests ------------------------------
def test_known_exposure_pairs():
    z = golden("known_answers")
    i, j, r = orc.exposure_pairs(z["pairs_exposure"])
    assert np.array_equal(i, z["pairs_i"]) and np.array_equal(j, z["pairs_j"])
    assert np.array_equal(r.astype(np.float32), z["pairs_r"])
    assert np.array_equal(i, [0, 0, 1]) and np.array_equal(j, [1, 2, 2])          # test_general_functions.py:298-300
    i, j, r = orc.exposure_pairs(z["pairs_exposure"], float(z["pairs_thr"]))
    assert np.array_equal(i, z["pairs_thr_i"]) and np.array_equal(j, z["pairs_thr_j"])
    assert np.array_equal(i, [0, 1]) and np.array_equal(j, [1, 2])                # :313-315
    i, j, r = orc.exposure_pairs(z["pairs6_exposure"], float(z["pairs6_thr"]))
    assert np.array_equal(i, z["pairs6_i"]) and np.array_equal(j, z["pairs6_j"]) and np.array_equal(r, z["pairs6_r"])


def test_known_valid_masks():
    z = golden("known_answers")
    m = orc.pair_valid_mask(z["mask_stack"], z["mask_i"], z["mask_j"], float(z["mask_lo"]), float(z["mask_hi"]))
    assert np.array_equal(m, z["mask_expected"])
    k16 = (np.arange(65536, dtype=np.float32) / np.float32(65535.0))
    v16 = orc.frame_valid(k16, 1 / 255, 254 / 255)
    assert np.array_equal(v16, z["valid16"])
    assert v16[257] and not v16[256] and v16[65278] and not v16[65279]
    k8 = (np.arange(256, dtype=np.float32) / np.float32(255.0))
    assert np.array_equal(orc.frame_valid(k8, 1 / 255, 254 / 255), z["valid8"])


def test_known_gaussian_weights():
    z = golden("known_answers")
    assert max_rel(orc.gaussian_value_weights(z["gw_x"], 30.0), z["gw_30"]) < 5e-7
    assert max_rel(orc.gaussian_value_weights(z["gw_x"], 10.0), z["gw_10"]) < 5e-7
    assert orc.gaussian_value_weights(np.float32(0.5)) == 1.0                      # test_losses.py: peak at 0.5


def test_known_running_weighted_mean():
    """WBOMean (common/statistics.py) equals the direct weighted mean; HdrState uses the same merge."""
    z = golden("known_answers")
    v, w = z["wbo_values"], z["wbo_weights"]
    mean, wsum = 0.0, 0.0
    for a in range(0, 12, 5):
        wb = w[a:a + 5].sum(0)
        mb = (w[a:a + 5] * v[a:a + 5]).sum(0) / (wb + 1e-6)
        tot = wsum + wb
        mean = mean + (wb / tot) * (mb - mean)
        wsum = tot
    assert max_rel(mean, z["wbo_mean"][0]) < 1e-12
    assert max_rel(wsum, z["wbo_wsum"][0]) < 1e-12


# ---- ICRF evaluation ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("forward_linear"))
def test_forward_linear_bit_exact(name):
    z = golden(name)
    f, fp, x0, rows = orc.icrf_linear(z["x"], z["theta"])
    assert np.array_equal(f, z["y"])
    assert np.array_equal(fp, z["dydx"])


@pytest.mark.parametrize("name", golden_names("forward_catmull"))
def test_forward_catmull(name):
    z = golden(name)
    f, fp, taps, w, rows = orc.icrf_catmull(z["x"], z["theta"])
    assert np.array_equal(f, z["y"])
    assert max_abs_over_max(fp, z["dydx"]) < 1e-4      # the reference's fp32 autograd vs the closed form
    gt = np.zeros(z["theta"].shape)
    for ix, wi in zip(taps, w):
        np.add.at(gt, (rows.ravel(), ix.ravel()), (z["upstream"] * wi).ravel().astype(np.float64))
    assert max_abs_over_max(gt, z["grad_theta"]) < 1e-6


@pytest.mark.parametrize("name", golden_names("forward_lookup"))
def test_forward_lookup_bit_exact(name):
    z = golden(name)
    y, _ = orc.icrf_lookup(z["x"], z["theta"])
    assert np.array_equal(y, z["y"])


@pytest.mark.parametrize("name", golden_names("forward_codes"))
def test_forward_all_codes(name):
    z = golden(name)
    f, _, x0, _ = orc.icrf_linear(z["x"], z["theta"])
    y, idx = orc.icrf_lookup(z["x"], z["theta"])
    assert np.array_equal(f, z["y_linear"]) and np.array_equal(y, z["y_lookup"])
    if name.endswith("u8"):   # fl32(k/255)*255 == k exactly, so x0 == k
        assert np.array_equal(x0[0, 0, 0], np.arange(256))


# ---- HDR merge ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("hdr_"))
def test_hdr_merge(name):
    z = golden(name)
    rad, sig = orc.hdr_merge(z["val"], _opt(z, "std"), z["exposure"], _opt(z, "theta"), bool(z["gaussian"]),
                             int(z["batch_size"]), mode=mode_of(z))
    assert max_rel(rad.reshape(z["radiance"].shape), z["radiance"]) < 5e-7
    if "sigma" in z:
        # the reference's own fp32 autograd carries ~5e-6 of cancellation noise at bright codes; more in the
        # other modes (SIGMA_TOL in _helpers.py)
        assert max_rel(sig.reshape(z["sigma"].shape), z["sigma"]) < SIGMA_TOL[mode_of(z)]
    else:
        assert sig is None


@pytest.mark.parametrize("name", golden_names("linearize_"))
def test_linearize_bit_exact(name):
    z = golden(name)
    for n in range(z["val"].shape[0]):
        std = None if "std" not in z else z["std"][n:n + 1]
        f, s = orc.linearize(z["val"][n:n + 1], std, z["theta"], mode=mode_of(z))
        assert np.array_equal(f[0], z["linearized"][n])
        if mode_of(z) == "catmull":     # the oracle's derivative is the closed form, the reference's its own fp32 autograd
            assert np.max(np.abs(s[0] - z["sigma"][n])) <= 5e-5 * np.max(z["sigma"][n])
        else:
            assert np.array_equal(s[0], z["sigma"][n])


# ---- linearity measurement ---------------------------------------------------------------------
def _linearity_flags(name):
    if name == "linearity_u16_w11":
        return True, True
    return "rel1" in name, "unc1" in name


@pytest.mark.parametrize("name", golden_names("linearity_"))
def test_linearity_stats(name):
    z = golden(name)
    rel, unc = _linearity_flags(name)
    ratio, m, sd, em = orc.linearity_stats(z["val"], _opt(z, "std"), z["exposure"], _opt(z, "theta"), 0.2,
                                           relative=rel, unc_weighting=unc, mode=mode_of(z))
    assert np.array_equal(ratio, z["ratio"])
    assert max_rel(m, z["mean"]) < 1e-7
    assert max_rel(sd, z["stddev"]) < 1e-7
    if "errmean" in z:
        assert max_rel(em, z["errmean"]) < 1e-12
    else:
        assert em is None


# ---- training step -----------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("trainstep_"))
def test_train_step(name):
    z = golden(name)
    c = z["theta0"].shape[0]
    adams = [orc.Adam(z["theta0"][k].shape) for k in range(c)]
    for step in range(int(z["n_steps"])):
        theta = z["theta0"] if step == 0 else z[f"theta_after_{step - 1}"]
        out = orc.train_loss_and_grad(z["val"], _opt(z, "std"), z["exposure"], theta, float(z["thr"]),
                                      relative=bool(z["rel"]), unc_weighting=bool(z["unc"]),
                                      coeffs=tuple(z["coeffs"]), mode=mode_of(z))
        assert max_rel(out["loss"], z[f"loss_{step}"]) < 1e-6
        assert max_rel(out["linloss"], z[f"linloss_{step}"]) < 1e-6
        assert max_rel(out["spatial"], z[f"spatial_{step}"]) < 1e-6
        assert max_abs_over_max(out["grad"], z[f"grad_theta_{step}"]) < TOL
        # Adam on the reference's own gradient reproduces the reference's next table
        nxt = np.stack([adams[k].step(theta[k], z[f"grad_theta_{step}"][k]) for k in range(c)])
        assert np.max(np.abs(nxt - z[f"theta_after_{step}"])) < 2e-7


# ---- the C (OpenMP) twin of the oracle: same fixtures, same bars ---------------------------------------
from oracle import c_oracle as corc  # noqa: E402


@pytest.mark.parametrize("name", golden_names("forward_linear") + golden_names("forward_lookup"))
def test_c_forward_bit_exact(name):
    z = golden(name)
    if "lookup" in name:
        y, _ = corc.icrf_lookup(z["x"], z["theta"])
        assert np.array_equal(y, z["y"])
    else:
        f, fp, x0 = corc.icrf_linear(z["x"], z["theta"])
        assert np.array_equal(f, z["y"]) and np.array_equal(fp, z["dydx"])
        assert np.array_equal(x0, orc.icrf_linear(z["x"], z["theta"])[2])


@pytest.mark.parametrize("name", linear_golden_names("hdr_"))
def test_c_hdr_merge(name):
    z = golden(name)
    rad, sig = corc.hdr_merge(z["val"], _opt(z, "std"), z["exposure"], _opt(z, "theta"), bool(z["gaussian"]),
                              int(z["batch_size"]))
    assert max_rel(rad.reshape(z["radiance"].shape), z["radiance"]) < 5e-7
    if "sigma" in z:
        assert max_rel(sig.reshape(z["sigma"].shape), z["sigma"]) < TOL


@pytest.mark.parametrize("name", linear_golden_names("linearize_"))
def test_c_linearize_bit_exact(name):
    z = golden(name)
    lin, sig = corc.linearize(z["val"], _opt(z, "std"), z["theta"])
    assert np.array_equal(lin, z["linearized"]) and np.array_equal(sig, z["sigma"])


@pytest.mark.parametrize("name", linear_golden_names("linearity_"))
def test_c_linearity_stats(name):
    z = golden(name)
    rel, unc = _linearity_flags(name)
    i, j, r = orc.exposure_pairs(z["exposure"], 0.2)
    m, sd, em = corc.pair_stats(z["val"], _opt(z, "std"), i, j, r, _opt(z, "theta"), relative=rel, unc_weighting=unc)
    assert max_rel(m, z["mean"]) < 1e-7 and max_rel(sd, z["stddev"]) < 1e-7
    if "errmean" in z:
        assert max_rel(em, z["errmean"]) < 1e-12


@pytest.mark.parametrize("name", linear_golden_names("trainstep_"))
def test_c_train_grad(name):
    z = golden(name)
    i, j, r = orc.exposure_pairs(z["exposure"], float(z["thr"]))
    for step in range(int(z["n_steps"])):
        theta = z["theta0"] if step == 0 else z[f"theta_after_{step - 1}"]
        lin, mean, grad = corc.train_grad(z["val"], _opt(z, "std"), i, j, r, theta, relative=bool(z["rel"]),
                                          unc_weighting=bool(z["unc"]))
        assert max_rel(lin, z[f"linloss_{step}"]) < 1e-6
        assert max_rel(mean, z[f"spatial_{step}"]) < 1e-6
        _, gpens = orc.curve_penalties(theta)
        total = grad + sum(k * g for k, g in zip(tuple(z["coeffs"]), gpens))
        assert max_abs_over_max(total, z[f"grad_theta_{step}"]) < TOL


# ---- streaming frame statistics -------------------------------------------------
def test_frame_stats_wbomeanvar():
    z = golden("framestats_wbomeanvar")
    for tag, w in (("w", z["weights"]), ("u", None)):
        mean, m2, wsum, wsq = orc.frame_stats(z["val"], w, z["bounds"])
        assert max_rel(mean, z[f"mean_{tag}"][0]) < 1e-6 and max_rel(m2, z[f"m2_{tag}"][0]) < 2e-5
        assert max_rel(wsum, z[f"wsum_{tag}"][0]) < 1e-6 and max_rel(wsq, z[f"wsq_{tag}"][0]) < 1e-6
        assert max_rel(m2 / (wsum - wsq / wsum), z[f"var_rel_{tag}"][0]) < 2e-5


@pytest.mark.parametrize("name", golden_names("framestats_video"))
def test_frame_stats_video(name):
    z = golden(name)
    n = z["val"].shape[0]
    order = np.argsort(z["exposure"], kind="stable")
    bounds = list(range(0, n, int(z["batch_size"]))) + [n]
    mean, m2, wsum, _ = orc.frame_stats(z["val"][order], None, bounds, _opt(z, "theta"))
    assert max_rel(mean, z["mean"]) < 1e-6
    assert max_rel(np.sqrt(m2 / (wsum - 1)) / np.sqrt(n), z["sem"]) < 2e-5


# ---- dark-field / flat-field corrections -----------------------------------------
@pytest.mark.parametrize("name", golden_names("artefact_"))
def test_artefact_corrections(name):
    """Closed forms of the dark-field mix (gradient taken wrt the mixed image, both variance terms folded into one
    effective std) and of the flat-field correction (mean inside / outside the graph) against the reference."""
    z = golden(name)
    val, std = z["val"], z["std"]
    if "dark" in z:
        mixed, seff = orc.dark_field_mix(val, std, z["dark"], z["dark_std"])
        mixed, seff = mixed.astype(np.float32), seff.astype(np.float32)
    else:
        mixed, seff = val, std
    rad, sig = orc.hdr_merge(mixed, seff, z["exposure"], z["theta"], True, int(z["batch_size"]))
    var = sig ** 2
    if "flat" in z:
        rad, var = orc.flat_field_correct(rad, var, z["flat"][0], z["flat_std"][0], True)
    assert max_rel(rad, z["radiance"]) < 2e-6
    assert max_rel(np.sqrt(var), z["sigma"]) < TOL
    for i in range(val.shape[0]):
        f, s = orc.linearize(mixed[i:i + 1], seff[i:i + 1], z["theta"])
        lin, v = f.astype(np.float64), s.astype(np.float64) ** 2
        if "flat" in z:
            lin, v = orc.flat_field_correct(lin, v, z["flat"][0], z["flat_std"][0], False)
        assert max_rel(lin[0], z["linearized"][i]) < 2e-6
        assert max_rel(np.sqrt(v[0]), z["lin_sigma"][i]) < TOL


@pytest.mark.parametrize("name", golden_names("ingest_"))
def test_ingest_chain_bit_exact(name):
    """CvToTorch + CastTo(float32) + Normalize(max, 0) on a camera buffer holding every code of its range, and
    the std the datasets synthesise — the oracle's restatement against what the reference's transforms returned, to the bit.
    For min 0 and the default target range the chain is fl32(code) / fl32(max): the quotient the CUDA ingest kernels form."""
    z = golden(name)
    val = orc.cast_normalize(orc.cv_to_torch(z["camera"]), float(z["max_val"]))
    assert val.dtype == np.float32 and np.array_equal(val, z["val"])
    planar_codes = orc.cv_to_torch(z["camera"]).astype(np.float32)
    assert np.array_equal(planar_codes / np.float32(z["max_val"]), z["val"])
    assert np.array_equal(orc.missing_std(val, "multiplier", float(z["multiplier"])), z["std_multiplier"])
    assert np.array_equal(orc.missing_std(val, "constant", float(z["constant"])), z["std_constant"])
    assert orc.missing_std(val, "none", 0.0) is None
    with pytest.raises(ValueError):
        orc.cast_normalize(z["camera"], 5.0, 5.0)
    with pytest.raises(ValueError):
        orc.cv_to_torch(np.zeros((4, 4, 2), dtype=np.uint8))
    assert orc.cv_to_torch(np.zeros((4, 5), dtype=np.uint8)).shape == (1, 4, 5)
    # a non-trivial range exercises the general form: ((x - min) / (max - min)) * span + target_min in fp32
    x = orc.cv_to_torch(z["camera"]).astype(np.float32)
    want = ((x - np.float32(3.0)) / np.float32(float(z["max_val"]) - 3.0)) * np.float32(2.0) + np.float32(-1.0)
    assert np.array_equal(orc.cast_normalize(orc.cv_to_torch(z["camera"]), float(z["max_val"]), 3.0, (-1.0, 1.0)), want)


# ---- the reference itself, run from oracle/_ref (bench.py's CPU arm) ----------------------------------------------
def test_reference_runner_reproduces_the_fixtures():
    """oracle/reference_runner.py drives the UNMODIFIED reference vendored into oracle/_ref (make -C oracle ref); the
    fixtures were written by the same code imported from /root/reference, so the runner must reproduce them to the bit.
    This pins what bench.py times as `cpu_baseline.kind = "reference"` and `--impl reference`."""
    from oracle import reference_runner as rr
    if not rr.available() and not rr.vendor():
        pytest.skip("oracle/_ref is not vendored here and /root/reference is absent")
    z = golden("hdr_c1small_2batches")
    rad, sig, _ = rr.hdr_merge(z["val"], z["std"], z["exposure"], z["theta"], bool(z["gaussian"]), int(z["batch_size"]))
    assert np.array_equal(rad, z["radiance"]) and np.array_equal(sig, z["sigma"])
    z = golden("linearity_rel1_unc1_std1_model1")
    ratio, mean, sd, err, _ = rr.measure_linearity(z["val"], z["std"], z["exposure"], z["theta"], True, True)
    assert np.array_equal(ratio, z["ratio"]) and np.array_equal(mean, z["mean"]) and np.array_equal(sd, z["stddev"])
    assert np.array_equal(err, z["errmean"])
    z = golden("trainstep_script")
    a, b, g, d = (float(k) for k in z["coeffs"])
    step = rr.TrainStep(z["val"], z["std"], z["exposure"], theta=z["theta0"], relative=bool(z["rel"]), unc_weighting=bool(z["unc"]),
                        alpha=a, beta=b, gamma=g, delta=d, threshold=float(z["thr"]))
    for k in range(2):
        loss, lin, grad, _ = step()
        assert np.array_equal(loss, z[f"loss_{k}"]) and np.array_equal(lin, z[f"linloss_{k}"])
        assert np.array_equal(grad, z[f"grad_theta_{k}"])
