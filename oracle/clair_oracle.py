"""CPU oracle (numpy) for the per-pixel radiometric hot path of samivout/clair-torch.

TEST INFRASTRUCTURE ONLY.  This module is a closed-form CPU restatement of the
reference's algorithm; it exists to check the CUDA path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it.
The product package (`clair_torch_b200`) never imports anything from `oracle/`.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function
here against fixtures under `tests/golden/` that were produced by running the
unmodified reference (see `tests/golden/make_golden.py`), including the
known-answer vectors of the reference's own unit tests.

The reference evaluates everything through chains of ATen ops plus
`torch.autograd`; here the same quantities are written in closed form (no
autograd).  Where a result decides an integer (LUT index, validity mask) the
arithmetic is done in float32 in the reference's op order so that it is
bit-exact; everything else is float64, which the reference's own mixed
fp32/fp64 results agree with to ~5e-6.

All `file:line` citations are relative to the reference repository root.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
F64 = np.float64

HDR_WEIGHT_SCALE = 30.0   # training/losses.py:193 default, used by inference/hdr_merge.py:95
PAIR_WEIGHT_SCALE = 10.0  # training/losses.py:212 default


# ----------------------------------------------------------------------------------------------
# ICRF evaluation
# ----------------------------------------------------------------------------------------------
def curve_rows(shape, flat_offset=0):
    """Row of the (C, L) table each element of a contiguous (..., C, H, W) batch reads in LINEAR mode.

    models/base.py:173-176 pairs the NCHW-flattened indices with `arange(C).repeat(N*H*W)`, so the
    element with flat index k uses row `k mod C` rather than its own channel.
    `flat_offset` is the flat index of this array's first element inside the full batch tensor
    (non-zero only for a spatial shard of a larger image).
    """
    n_elem = int(np.prod(shape))
    c = shape[-3]
    return ((np.arange(n_elem, dtype=np.int64) + int(flat_offset)) % c).reshape(shape)


def icrf_linear(x, theta, flat_offset=0, rows=None):
    """LINEAR-mode ICRF, models/base.py:160-182, and its autograd derivative d f / d x.

    Returns (f, fprime, x0, rows): f and fprime float32 (bit-exact with the reference), x0 the lower
    LUT index (int64), rows the table row used per element.
    """
    x = np.ascontiguousarray(x, dtype=F32)
    theta = np.asarray(theta, dtype=F32)
    n_rows, l = theta.shape
    lm1 = F32(l - 1)
    xs_raw = x * lm1                                     # :167 (fp32 multiply)
    xs = np.minimum(np.maximum(xs_raw, F32(0)), lm1)     # :167 clamp_
    x0 = np.floor(xs).astype(np.int64)                   # :169
    x1 = np.minimum(x0 + 1, l - 1)                       # :170
    w = xs - x0.astype(F32)                              # :171
    if rows is None:
        rows = curve_rows(x.shape, flat_offset)
    g0 = theta[rows, x0]
    g1 = theta[rows, x1]
    f = g0 * (F32(1.0) - w) + g1 * w                     # :182, separate fp32 roundings
    inside = (xs_raw >= F32(0)) & (xs_raw <= lm1)        # clamp backward passes grad on the closed range
    fprime = np.where(inside, (g1 - g0) * lm1, F32(0)).astype(F32)
    return f.astype(F32), fprime, x0, rows


def icrf_catmull(x, theta, flat_offset=0, rows=None):
    """CATMULL-mode ICRF, models/base.py:184-226 (four-tap Catmull-Rom, rows per the same k-mod-C rule as LINEAR).

    Returns (f float32 — same op order as the reference, fprime float64 closed form, taps (4 index arrays), weights
    (4 float32 arrays), rows).
    """
    x = np.ascontiguousarray(x, dtype=F32)
    theta = np.asarray(theta, dtype=F32)
    l = theta.shape[1]
    lm1 = F32(l - 1)
    xs_raw = x * lm1
    xs = np.minimum(np.maximum(xs_raw, F32(0)), lm1)                       # :190
    x0 = np.floor(xs).astype(np.int64)                                     # :193
    taps = [np.clip(x0 + d, 0, l - 1) for d in (-1, 0, 1, 2)]              # :194-199
    t = np.minimum(np.maximum(xs - x0.astype(F32), F32(0)), F32(1))        # :202
    t2 = t * t
    t3 = t2 * t
    w = [F32(-0.5) * t3 + t2 - F32(0.5) * t,                               # :208-211, left-to-right fp32
         F32(1.5) * t3 - F32(2.5) * t2 + F32(1.0),
         F32(-1.5) * t3 + F32(2.0) * t2 + F32(0.5) * t,
         F32(0.5) * t3 - F32(0.5) * t2]
    if rows is None:
        rows = curve_rows(x.shape, flat_offset)
    g = [theta[rows, ix] for ix in taps]
    f = ((w[0] * g[0] + w[1] * g[1]) + w[2] * g[2]) + w[3] * g[3]          # stack(...).sum(dim=0), :224
    td = t.astype(F64)
    dw = [-1.5 * td * td + 2 * td - 0.5, 4.5 * td * td - 5 * td, -4.5 * td * td + 4 * td + 0.5, 1.5 * td * td - td]
    inside = (xs_raw >= F32(0)) & (xs_raw <= lm1)
    fprime = np.where(inside, sum(d * gi.astype(F64) for d, gi in zip(dw, g)) * float(l - 1), 0.0)
    return f.astype(F32), fprime, taps, w, rows


def icrf_lookup(x, theta):
    """LOOKUP-mode ICRF, models/base.py:138-158: round-half-even index, true channel row."""
    x = np.ascontiguousarray(x, dtype=F32)
    theta = np.asarray(theta, dtype=F32)
    l = theta.shape[1]
    idx = np.clip(np.rint(x * F32(l - 1)), 0, l - 1).astype(np.int64)   # :145
    c = x.shape[-3]
    chan = np.arange(c).reshape((c, 1, 1))
    chan = np.broadcast_to(chan, x.shape)
    return theta[chan, idx], idx


def icrf_eval(x, theta, mode="linear", flat_offset=0):
    """ICRFModelBase.forward in any of the three modes plus the derivative autograd hands back for d f / d x.

    Returns dict(f32, fp (float64), taps [index arrays], weights [float64 arrays], rows): `taps` / `weights` are what the
    backward pass scatters the upstream gradient to (two for LINEAR, four for CATMULL, one with weight 1 for LOOKUP,
    whose image edge does not exist: fp = 0).
    """
    if mode == "linear":
        f32, fp, x0, rows = icrf_linear(x, theta, flat_offset)
        l = np.asarray(theta).shape[1]
        xs = np.minimum(np.maximum(np.asarray(x, dtype=F32) * F32(l - 1), F32(0)), F32(l - 1))
        w = (xs - x0.astype(F32)).astype(F64)
        return {"f32": f32, "fp": fp.astype(F64), "taps": [x0, np.minimum(x0 + 1, l - 1)], "weights": [1.0 - w, w], "rows": rows}
    if mode == "catmull":
        f32, fp, taps, w, rows = icrf_catmull(x, theta, flat_offset)
        return {"f32": f32, "fp": np.asarray(fp, dtype=F64), "taps": taps, "weights": [wi.astype(F64) for wi in w], "rows": rows}
    if mode == "lookup":
        f32, idx = icrf_lookup(x, theta)
        c = np.asarray(x).shape[-3]
        rows = np.broadcast_to(np.arange(c).reshape((c, 1, 1)), np.asarray(x).shape)
        return {"f32": f32.astype(F32), "fp": np.zeros(np.asarray(x).shape, dtype=F64), "taps": [idx],
                "weights": [np.ones(np.asarray(x).shape, dtype=F64)], "rows": rows}
    raise ValueError(f"unknown mode {mode!r}")


def gaussian_value_weights(x, scale=HDR_WEIGHT_SCALE):
    """training/losses.py:193-205 in float32."""
    x = np.asarray(x, dtype=F32)
    d = x - F32(0.5)
    return np.exp(F32(-scale) * (d * d)).astype(F32)


# ----------------------------------------------------------------------------------------------
# HDR merge with first-order uncertainty
# ----------------------------------------------------------------------------------------------
class HdrState:
    """Running state of the merge: WBOMean's (mean, sum_of_weights) plus the running variance.

    common/statistics.py:27-29 starts both at the python float 0.0; inference/hdr_merge.py:55
    starts the variance at None.
    """

    def __init__(self):
        self.mean = None     # float64 (C,H,W)
        self.wsum = None     # float64 here; float32 in the reference
        self.var = None      # float64 here; float32 in the reference


def hdr_merge_update(state, val, std, exposure, theta=None, gaussian=True, flat_offset=0, mode="linear"):
    """One DataLoader batch of compute_hdr_image (inference/hdr_merge.py:95-128).

    val, std: (N,C,H,W) float32 (std may be None); exposure: (N,) float64 seconds; theta (C,L) or None
    (`icrf_model=None` => identity, :99-100).  `gaussian` mirrors `weight_fn is not None` (:95).

    Closed form of the autograd pass at :107-115:
      w_n = exp(-30 (x_n-.5)^2) | 1,   v_n = f(x_n)/t_n,   W_B = sum w_n,   mean_B = sum w_n v_n / (W_B+1e-6)
      W = W_A + W_B,   mean = mean_A + (W_B/W)(mean_B - mean_A)                    (statistics.py:74-109)
      g_n = d mean / d x_n = (W_B/W) [w_n f'_n/t_n + w'_n (v_n-mean_B)]/(W_B+1e-6) + w'_n (W_A/W^2)(mean_B-mean_A)
      var += sum_n (g_n s_n)^2                                                        (hdr_merge.py:114-115)
    """
    val = np.ascontiguousarray(val, dtype=F32)
    n = val.shape[0]
    t = np.asarray(exposure, dtype=F64).reshape(n, 1, 1, 1)
    if theta is not None:
        ev = icrf_eval(val, theta, mode, flat_offset)
        f = ev["f32"].astype(F64)
        fp = ev["fp"]
    else:
        f = val.astype(F64)
        fp = np.ones_like(f)
    x = val.astype(F64)
    if gaussian:
        w = gaussian_value_weights(val, HDR_WEIGHT_SCALE).astype(F64)
        wp = -2.0 * HDR_WEIGHT_SCALE * (x - 0.5) * w
    else:
        w = np.ones_like(f)
        wp = np.zeros_like(f)
    v = f / t
    # statistics.py:75-77.  The reference sums its float32 weights in float32; here W_B is float64: the difference (6e-8
    # relative) only matters for a LOOKUP model, whose whole uncertainty is the cancelling term w'(v_n - mean_B) below and
    # therefore inherits the rounding of W_B amplified by mean_B / (v_n - mean_B) (SIGMA_TOL in tests/_helpers.py).
    w_b = w.sum(axis=0)
    w_be = w_b + 1e-6
    mean_b = (w * v).sum(axis=0) / w_be
    if state.mean is None:
        w_a = np.zeros_like(w_b)
        mean_a = np.zeros_like(mean_b)
    else:
        w_a, mean_a = state.wsum, state.mean
    w_tot = w_a + w_b
    frac = w_b / w_tot
    mean_new = mean_a + frac * (mean_b - mean_a)
    if std is not None:
        s = np.asarray(std, dtype=F32).astype(F64)
        dmean_b = (w * fp / t + wp * (v - mean_b)) / w_be
        g = frac * dmean_b + wp * (w_a / (w_tot * w_tot)) * (mean_b - mean_a)
        upd = ((g * s) ** 2).sum(axis=0)
        state.var = upd if state.var is None else state.var + upd
    state.mean, state.wsum = mean_new, w_tot
    return state


def hdr_merge(val, std, exposure, theta=None, gaussian=True, batch_size=None, flat_offset=0, mode="linear"):
    """compute_hdr_image (inference/hdr_merge.py:19-155) without flat/dark-field branches.

    The stack is assumed sorted by ascending exposure inside each batch (datasets/collate.py:23).
    Returns (radiance float64 (C,H,W), sigma float64 or None).
    """
    n = val.shape[0]
    bs = n if batch_size is None else int(batch_size)
    st = HdrState()
    for a in range(0, n, bs):
        sl = slice(a, min(a + bs, n))
        hdr_merge_update(st, val[sl], None if std is None else std[sl], np.asarray(exposure)[sl], theta,
                         gaussian, flat_offset, mode)
    sigma = None if st.var is None else np.sqrt(st.var)
    return st.mean, sigma


def linearize(val, std, theta, flat_offset=0, mode="linear"):
    """linearize_dataset_generator core for one (1,C,H,W) image (inference/linearization.py:94-106,132).

    Returns (f float32, sigma float32) with sigma = sqrt((f'(x) s)^2); zeros when std is None (:97).
    """
    if mode != "linear":
        ev = icrf_eval(val, theta, mode, flat_offset)
        f = ev["f32"]
        if std is None:
            return f, np.zeros_like(f)
        g = ev["fp"].astype(F32) * np.asarray(std, dtype=F32)
        return f, np.sqrt(g * g).astype(F32)
    f, fp, _, _ = icrf_linear(val, theta, flat_offset)
    if std is None:
        return f, np.zeros_like(f)
    g = fp * np.asarray(std, dtype=F32)
    return f, np.sqrt(g * g).astype(F32)


# ----------------------------------------------------------------------------------------------
# Image ingest: what the reference's transform chain and datasets do to a camera buffer before the hot path sees it
# ----------------------------------------------------------------------------------------------
def cv_to_torch(x):
    """CvToTorch (common/transforms.py:67-84, common/general_functions.py:315-336): an OpenCV image (H,W) or (H,W,3) BGR
    becomes (1,H,W) or (3,H,W) RGB."""
    x = np.asarray(x)
    if x.ndim == 2:
        return x[None]
    if x.ndim == 3 and x.shape[2] == 3:
        return np.transpose(x[:, :, [2, 1, 0]], (2, 0, 1))
    raise ValueError(f"Unexpected image shape: {x.shape}")


def cast_normalize(codes, max_val, min_val=0.0, target_range=(0.0, 1.0)):
    """CastTo(float32) then Normalize(max_val, min_val) (common/transforms.py:107-131,161-183;
    common/general_functions.py:373-388): ((x - min) / (max - min)) * span + target_min, every step an fp32 op with the
    Python scalars rounded to fp32 — for min 0 and the default range that is fl32(code) / fl32(max), an IEEE division."""
    x = np.asarray(codes).astype(F32)
    denominator = max_val - min_val                      # Python floats (:376)
    if denominator == 0:
        raise ValueError("Normalization range is zero (min == max); cannot normalize.")
    x = (x - F32(min_val)) / F32(denominator)
    lo, hi = target_range
    return (x * F32(hi - lo) + F32(lo)).astype(F32)


def missing_std(val, mode, value):
    """The std image a dataset makes up when no std file exists (datasets/base.py:35,128-133): None, a constant plane, or
    val * fl32(value)."""
    if mode == "none":
        return None
    if mode == "constant":
        return np.full_like(np.asarray(val, dtype=F32), F32(value))
    if mode == "multiplier":
        return (np.asarray(val, dtype=F32) * F32(value)).astype(F32)
    raise ValueError(f"Unsupported MissingStdMode: {mode}")


# ----------------------------------------------------------------------------------------------
# Exposure pairs, validity mask, pair loss statistics
# ----------------------------------------------------------------------------------------------
def exposure_pairs(exposure, threshold=None):
    """get_valid_exposure_pairs, common/general_functions.py:242-272 (row-major upper triangle)."""
    t = np.asarray(exposure)
    n = t.shape[0]
    i_idx, j_idx = np.triu_indices(n, k=1)
    ratio = t[i_idx] / t[j_idx]
    if threshold is not None:
        keep = ratio >= threshold
        i_idx, j_idx, ratio = i_idx[keep], j_idx[keep], ratio[keep]
    return i_idx.astype(np.int64), j_idx.astype(np.int64), ratio


def frame_valid(val, lo, hi):
    """Per-frame half of get_pairwise_valid_pixel_mask (general_functions.py:303-305).

    The python-float thresholds are compared in the tensor's dtype, i.e. after rounding to float32.
    """
    val = np.asarray(val, dtype=F32)
    return (val >= F32(lo)) & (val <= F32(hi))


def pair_valid_mask(val, i_idx, j_idx, lo, hi):
    fv = frame_valid(val, lo, hi)
    return fv[i_idx] & fv[j_idx]


def _pair_terms(val, std, i_idx, j_idx, ratio, theta, lo, hi, relative, unc_weighting, flat_offset, mode="linear"):
    """Everything per pair-element that losses.py:13-67,70-108 and general_functions.py:118-178 form.

    Returns a dict of float64 (P,C,H,W) arrays plus the per-frame ICRF pieces.
    """
    val = np.ascontiguousarray(val, dtype=F32)
    ev = None
    if theta is not None:
        ev = icrf_eval(val, theta, mode, flat_offset)
        f32, fp32_ = ev["f32"], ev["fp"].astype(F32)
    else:
        f32, fp32_ = val, np.ones_like(val)
    f = f32.astype(F64)
    have_std = std is not None
    if have_std:
        sig32 = np.abs(fp32_ * np.asarray(std, dtype=F32))            # icrf_training.py:124 (fp32)
        sig = sig32.astype(F64)
    r = np.asarray(ratio, dtype=F64).reshape(-1, 1, 1, 1)
    a, b = f[i_idx], f[j_idx]
    e = b * r                                                         # losses.py:40 (promotes to fp64)
    d = a - e
    out = {"f32": f32, "fp32": fp32_, "ev": ev, "a": a, "b": b, "r": r}
    if relative:
        es = e + 1e-6                                                 # :45
        q = d / es
        out["es"] = es
    else:
        q = d
    ell = np.abs(q)
    out["sgn"] = np.sign(q)
    err = None
    if have_std:
        sa, sb = sig[i_idx], sig[j_idx]
        if relative:
            bs = np.maximum(f32[j_idx], F32(1e-6)).astype(F64)        # :55 clamp in fp32
            num = (f32[i_idx] * sig32[j_idx]).astype(F64)             # :58 fp32 product
            term = (sa / es) ** 2 + (num / (es * bs)) ** 2 + 1e-6
            err = np.sqrt(term)
            out.update(bs=bs, sa=sa, sb=sb)
        else:
            err = np.sqrt((sig32[i_idx] ** 2).astype(F64) + (r * sb) ** 2)   # :62
    gw = (gaussian_value_weights(val, PAIR_WEIGHT_SCALE)[i_idx]
          + gaussian_value_weights(val, PAIR_WEIGHT_SCALE)[j_idx]).astype(F64)   # losses.py:229-234 (fp32 add)
    wt = gw.copy()
    if err is not None and unc_weighting:
        wt = wt + 1.0 / (err + 1e-6)                                  # losses.py:97
    m = pair_valid_mask(val, i_idx, j_idx, lo, hi).astype(F64)
    out.update(ell=ell, err=err, wt=wt, mask=m)
    return out


def linearity_stats(val, std, exposure, theta=None, threshold=0.2, lo=1 / 255, hi=254 / 255, relative=True,
                    unc_weighting=True, flat_offset=0, pairs=None, mode="linear"):
    """measure_linearity (inference/measure_linearity.py:41-74).

    Returns (ratio (P,), mean (P,C), stddev (P,C), errmean (P,C) or None), float64.
    """
    i_idx, j_idx, ratio = exposure_pairs(exposure, threshold) if pairs is None else pairs
    tm = _pair_terms(val, std, i_idx, j_idx, ratio, theta, lo, hi, relative, unc_weighting, flat_offset, mode)
    m, wt, ell = tm["mask"], tm["wt"], tm["ell"]
    mw = m * wt
    dsum = np.maximum(mw.sum(axis=(2, 3)), 1e-8)                      # general_functions.py:156
    mean = (mw * ell).sum(axis=(2, 3)) / dsum
    dev = (ell * m - mean[:, :, None, None]) ** 2                     # values were masked at :149
    stddev = np.sqrt((dev * mw).sum(axis=(2, 3)) / dsum)
    errmean = None
    if tm["err"] is not None:
        errmean = (tm["err"] * m).sum(axis=(2, 3)) / np.maximum(m.sum(axis=(2, 3)), 1e-8)
    return np.asarray(ratio, dtype=F64), mean, stddev, errmean


# ----------------------------------------------------------------------------------------------
# ICRF training step
# ----------------------------------------------------------------------------------------------
def curve_penalties(theta):
    """The four per-channel penalties of training/losses.py:111-190 and their gradients wrt theta (C,L)."""
    th = np.asarray(theta, dtype=F64)
    df = th[:, 1:] - th[:, :-1]
    neg = (df <= 0).astype(F64)
    mono = (neg * df * df).sum(axis=1)
    g_mono = np.zeros_like(th)
    g_mono[:, 1:] += 2 * neg * df
    g_mono[:, :-1] -= 2 * neg * df
    sd = th[:, :-2] - 2 * th[:, 1:-1] + th[:, 2:]
    smooth = (sd * sd).sum(axis=1)
    g_smooth = np.zeros_like(th)
    g_smooth[:, :-2] += 2 * sd
    g_smooth[:, 1:-1] -= 4 * sd
    g_smooth[:, 2:] += 2 * sd
    rng = (np.maximum(-th, 0) + np.maximum(th - 1, 0)).sum(axis=1)
    g_rng = -(th < 0).astype(F64) + (th > 1).astype(F64)
    endp = th[:, 0] ** 2 + (th[:, -1] - 1) ** 2
    g_end = np.zeros_like(th)
    g_end[:, 0] = 2 * th[:, 0]
    g_end[:, -1] = 2 * (th[:, -1] - 1)
    return (mono, rng, endp, smooth), (g_mono, g_rng, g_end, g_smooth)


def train_loss_and_grad(val, std, exposure, theta, threshold=0.1, lo=1 / 255, hi=254 / 255, relative=True,
                        unc_weighting=True, coeffs=(1.0, 1.0, 1.0, 1.0), flat_offset=0, mode="linear"):
    """Loss and d(sum_c Loss_c)/d theta of one train_icrf step (training/icrf_training.py:105-149).

    Returns dict(loss (C,), linloss (C,), spatial (P,C), grad (C,L) float64, grad_lin (C,L)).
    Closed form of the C `backward` calls.
    """
    theta = np.asarray(theta, dtype=F32)
    n_rows, l = theta.shape
    i_idx, j_idx, ratio = exposure_pairs(exposure, threshold)
    tm = _pair_terms(val, std, i_idx, j_idx, ratio, theta, lo, hi, relative, unc_weighting, flat_offset, mode)
    m, wt, ell, sgn, r = tm["mask"], tm["wt"], tm["ell"], tm["sgn"], tm["r"]
    a, b = tm["a"], tm["b"]
    mw = m * wt
    wsum = mw.sum(axis=(2, 3))
    clamped = wsum < 1e-8
    dsum = np.maximum(wsum, 1e-8)
    spatial = (mw * ell).sum(axis=(2, 3)) / dsum                      # (P,C)
    linloss = np.sqrt((spatial ** 2).sum(axis=0))                     # icrf_training.py:136
    with np.errstate(divide="ignore", invalid="ignore"):
        up = np.where(linloss[None, :] > 0, spatial / linloss[None, :], 0.0) / dsum     # U_{p,c}
    up = up[:, :, None, None]
    if relative:
        es = tm["es"]
        dl_da = sgn / es
        dl_db = -sgn * r * (a + 1e-6) / (es * es)
    else:
        dl_da = sgn * np.ones_like(a)
        dl_db = -sgn * r
    g_a = mw * dl_da * up
    g_b = mw * dl_db * up
    if relative and unc_weighting and tm["err"] is not None:
        # weights depend on a and b through err (losses.py:55-60 are inside the graph)
        err, bs, sa, sb = tm["err"], tm["bs"], tm["sa"], tm["sb"]
        dm_dwt = m * (ell - np.where(clamped, 0.0, spatial)[:, :, None, None]) * up
        dm_dwt = np.where(clamped[:, :, None, None], 0.0, dm_dwt)
        dwt_dt = -1.0 / (err + 1e-6) ** 2 / (2.0 * err)               # dWt/derr * derr/dT
        dt_da = 2 * a * sb * sb / (es * bs) ** 2
        dt_des = -2 * sa * sa / es ** 3 - 2 * a * a * sb * sb / (es ** 3 * bs ** 2)
        dt_dbs = -2 * a * a * sb * sb / (es ** 2 * bs ** 3)
        dbs_db = (tm["f32"][j_idx] >= F32(1e-6)).astype(F64)
        g_a = g_a + dm_dwt * dwt_dt * dt_da
        g_b = g_b + dm_dwt * dwt_dt * (dt_des * r + dt_dbs * dbs_db)
    # gather per-frame upstream, then scatter to the taps of each element (models/base.py:173-182, :194-224)
    g_frame = np.zeros(tm["f32"].shape, dtype=F64)
    np.add.at(g_frame, i_idx, g_a)
    np.add.at(g_frame, j_idx, g_b)
    rows = tm["ev"]["rows"]
    grad_lin = np.zeros((n_rows, l), dtype=F64)
    for tap, w in zip(tm["ev"]["taps"], tm["ev"]["weights"]):
        np.add.at(grad_lin, (rows.ravel(), tap.ravel()), (g_frame * w).ravel())
    pens, gpens = curve_penalties(theta)
    loss = linloss.copy()
    grad = grad_lin.copy()
    for k, (p, g) in zip(coeffs, zip(pens, gpens)):
        loss = loss + k * p
        grad = grad + k * g
    return {"loss": loss, "linloss": linloss, "spatial": spatial, "grad": grad, "grad_lin": grad_lin,
            "pairs": (i_idx, j_idx, ratio)}


class Adam:
    """torch.optim.Adam(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, amsgrad=False) on one float32 array."""

    def __init__(self, shape, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.m = np.zeros(shape, dtype=F32)
        self.v = np.zeros(shape, dtype=F32)
        self.t = 0
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    def step(self, param, grad):
        g = np.asarray(grad, dtype=F32)
        self.t += 1
        self.m = (self.m + (g - self.m) * F32(1 - self.b1)).astype(F32)       # lerp_
        self.v = (self.v * F32(self.b2) + F32(1 - self.b2) * g * g).astype(F32)
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        denom = (np.sqrt(self.v) / F32(np.sqrt(bc2)) + F32(self.eps)).astype(F32)
        return (np.asarray(param, dtype=F32) - F32(self.lr / bc1) * (self.m / denom)).astype(F32)


# ----------------------------------------------------------------------------------------------
# Streaming weighted mean / second moment over frames
# ----------------------------------------------------------------------------------------------
def frame_stats(val, weights=None, bounds=None, theta=None):
    """WBOMeanVar over dim 0 of an (N,C,H,W) stack fed in the batches [bounds[k], bounds[k+1])
    (common/statistics.py:209-259), frames optionally linearised first
    (inference/inferential_statistics.py:43-46).  Returns float64 (mean, m2, wsum, wsq)."""
    val = np.asarray(val, dtype=F32)
    n = val.shape[0]
    bounds = [0, n] if bounds is None else list(bounds)
    v_all = icrf_linear(val, theta)[0].astype(F64) if theta is not None else val.astype(F64)
    mean = m2 = wsum = wsq = 0.0
    for a, b in zip(bounds[:-1], bounds[1:]):
        v = v_all[a:b]
        if weights is not None:
            w = np.asarray(weights, dtype=F32)[a:b].astype(F64)
            w_b = w.sum(0)
            wsq_b = (w * w).sum(0)
            mean_b = (w * v).sum(0) / (w_b + 1e-6)
            m2_b = (w * (v - mean_b) ** 2).sum(0)
        else:
            mean_b = v.mean(0)
            w_b = np.full_like(mean_b, float(b - a))
            wsq_b = w_b
            m2_b = ((v - mean_b) ** 2).sum(0)
        w_tot = wsum + w_b
        m2 = m2 + m2_b + (wsum * w_b / w_tot) * (mean_b - mean) ** 2
        mean = mean + (w_b / w_tot) * (mean_b - mean)
        wsum = w_tot
        wsq = wsq + wsq_b
    return mean, m2, wsum, wsq


# ----------------------------------------------------------------------------------------------
# Dark-field and flat-field corrections
# ----------------------------------------------------------------------------------------------
def gaussian_blur3(x):
    """torchvision GaussianBlur(kernel_size=3, sigma=1.0) on (..., H, W): reflect padding, taps
    exp(-0.5 d^2) / sum (common/general_functions.py:462-463)."""
    k1 = np.exp(-0.5 * np.array([-1.0, 0.0, 1.0]) ** 2)
    k1 = (k1 / k1.sum()).astype(F32).astype(F64)
    x = np.asarray(x, dtype=F64)
    p = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(1, 1), (1, 1)], mode="reflect")
    h, w = x.shape[-2:]
    out = np.zeros_like(x)
    for dy in range(3):
        for dx in range(3):
            out += k1[dy] * k1[dx] * p[..., dy:dy + h, dx:dx + w]
    return out


def dark_field_mix(val, std, dark, dark_std, threshold=0.05, alpha=50.0):
    """conditional_gaussian_blur(images, dark, 0.05, 3, differentiable=True) (common/general_functions.py:440-486) as
    used at inference/hdr_merge.py:89-92, plus the effective per-pixel std that carries BOTH variance terms of the
    drivers: the gradient is taken wrt the MIXED image (hdr_merge.py:107-113 re-uses the reassigned `images`), so
      var = sum (g' std)^2 + sum (g' dx'/dD dark_std)^2 = sum (g' std_eff)^2,   std_eff = sqrt(std^2 + (dx'/dD dark_std)^2),
      dx'/dD = (blur(x) - x) * alpha * m (1 - m),   m = sigmoid(alpha (D - threshold)).
    Returns (mixed float64, std_eff float64)."""
    x = np.asarray(val, dtype=F32).astype(F64)
    d = np.asarray(dark, dtype=F32).astype(F64)
    m = 1.0 / (1.0 + np.exp(-(d - threshold) * alpha))
    b = gaussian_blur3(x)
    mixed = m * b + (1.0 - m) * x
    dmix = (b - x) * alpha * m * (1.0 - m)
    s = np.asarray(std, dtype=F32).astype(F64)
    ds = np.asarray(dark_std, dtype=F32).astype(F64)
    return mixed, np.sqrt(s * s + (dmix * ds) ** 2)


def flat_field_correct(value, var, flat, flat_std, mean_in_graph=True):
    """flatfield_correction with the whole-image mean (flat_field_mean(F, 1.0)) and its variance term
    (common/general_functions.py:182-238; inference/hdr_merge.py:131-153 with the mean inside the autograd graph,
    inference/linearization.py:50-57,118-130 with the mean a constant — SURVEY.md Q11).  The existing variance is not
    rescaled by the gain.  value/var: (..., C, H, W); flat, flat_std: (C, H, W)."""
    value = np.asarray(value, dtype=F64)
    f = np.asarray(flat, dtype=F32).astype(F64)
    mu = f.mean(axis=(-1, -2), keepdims=True)
    corrected = value / (f + 1e-6) * mu
    g = -value * mu / (f + 1e-6) ** 2
    if mean_in_graph:
        g = g + (value / (f + 1e-6)).sum(axis=(-1, -2), keepdims=True) / (f.shape[-1] * f.shape[-2])
    fs = np.asarray(flat_std, dtype=F32).astype(F64)
    return corrected, np.asarray(var, dtype=F64) + (g * fs) ** 2
