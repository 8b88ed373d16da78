#!/usr/bin/env python
"""Generate the golden input/output fixtures in this directory by running the
UNMODIFIED reference (samivout/clair-torch, mounted read-only at /root/reference).

The reference is Python/torch, so it can be imported in the build container but
cannot travel to the GPU box; this script is committed together with the
`.npz` files it wrote.  Nothing under tests/ (other than this script) reads
/root/reference at run time.

Usage (build container only):   python tests/golden/make_golden.py

Each fixture stores the exact inputs (fp32 value/std stacks, float64 exposure
times, the ICRF table) and what the reference returned for them:

  forward_*.npz      ICRFModelBase.forward in LINEAR / LOOKUP mode + autograd d/dx
                     (clair_torch/models/base.py:135-182)
  hdr_*.npz          compute_hdr_image (clair_torch/inference/hdr_merge.py:19-155)
  linearize_*.npz    linearize_dataset_generator (clair_torch/inference/linearization.py:17-132)
  linearity_*.npz    measure_linearity (clair_torch/inference/measure_linearity.py:17-74)
  trainstep_*.npz    the step body of train_icrf (clair_torch/training/icrf_training.py:96-156),
                     rebuilt from the reference's own functions because train_icrf
                     itself crashes on CPU at :92
  ingest_*.npz       CvToTorch + CastTo + Normalize on a camera buffer of integer codes, and the synthesised std
                     (clair_torch/common/transforms.py:67-190, clair_torch/datasets/base.py:128-133)
  known_answers.npz  the known-answer vectors the reference's own unit tests pin
                     (tests/unit/common/test_general_functions.py:292-391,
                      tests/unit/common/test_statistics.py, tests/unit/training/test_losses.py)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_reference():
    def _stub(name, classes=()):
        m = types.ModuleType(name)
        for c in classes:
            setattr(m, c, type(c, (), {}))
        sys.modules[name] = m
        return m

    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    for fn in ("ion", "subplots", "pause", "tight_layout", "savefig", "clf"):
        setattr(plt, fn, lambda *a, **k: None)
    _stub("matplotlib.figure", ["Figure"])
    _stub("matplotlib.axes", ["Axes"])
    _stub("matplotlib.lines", ["Line2D"])
    sys.path.insert(0, REF)
    import clair_torch  # noqa: F401


_import_reference()

import torch  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

import clair_torch.inference  # noqa: E402,F401
from clair_torch.common.enums import InterpMode  # noqa: E402
from clair_torch.common.general_functions import (  # noqa: E402
    get_pairwise_valid_pixel_mask, get_valid_exposure_pairs, weighted_mean_and_std)
from clair_torch.common.statistics import WBOMean  # noqa: E402
from clair_torch.datasets.collate import custom_collate  # noqa: E402
from clair_torch.datasets.image_dataset import ImageMapDataset  # noqa: E402
from clair_torch.inference.hdr_merge import compute_hdr_image  # noqa: E402
from clair_torch.inference.linearization import linearize_dataset_generator  # noqa: E402
from clair_torch.inference.measure_linearity import measure_linearity  # noqa: E402
from clair_torch.models.icrf_model import ICRFModelDirect  # noqa: E402
from clair_torch.training.losses import (  # noqa: E402
    combined_gaussian_pair_weights, compute_endpoint_penalty, compute_monotonicity_penalty,
    compute_range_penalty, compute_smoothness_penalty, compute_spatial_linearity_loss,
    gaussian_value_weights, pixelwise_linearity_loss)

# the package re-exports functions under the same names as its submodules, so go through sys.modules
for _m in ("clair_torch.inference.hdr_merge", "clair_torch.inference.measure_linearity"):
    sys.modules[_m].tqdm = lambda it, **k: it
torch.set_num_threads(4)


class SyntheticStack(ImageMapDataset):
    """In-memory stand-in for a file-backed ImageMapDataset (typeguard requires the type)."""

    def __init__(self, vals, stds, exposures):
        self.vals, self.stds, self.exposures = vals, stds, exposures
        self.files = tuple(range(len(vals)))

    def __len__(self):
        return len(self.vals)

    def __getitem__(self, i):
        std = None if self.stds is None else self.stds[i].clone()
        return i, self.vals[i].clone(), std, {"exposure_time": float(self.exposures[i])}


def loader(vals, stds, exposures, batch_size):
    return DataLoader(SyntheticStack(vals, stds, exposures), batch_size=batch_size, shuffle=False,
                      collate_fn=custom_collate)


def make_stack(seed, n, c, h, w, bits, std_mult=0.05, first_exposure=1e-3):
    """Synthetic exposure stack, SURVEY.md §8(d): log-uniform scene, doubling exposures, gamma 1/2.2."""
    rng = np.random.default_rng(seed)
    scene = np.exp(rng.uniform(-7.0, 0.0, size=(c, h, w)))
    t = first_exposure * 2.0 ** np.arange(n)
    maxval = float(2 ** bits - 1)
    vals = []
    for k in range(n):
        v = np.clip(4.0 * scene * t[k] / t[-1], 0.0, 1.0) ** (1.0 / 2.2)
        q = np.round(maxval * v).astype(np.float32)
        vals.append(q / np.float32(maxval))
    vals = np.stack(vals).astype(np.float32)
    stds = None if std_mult is None else (vals * np.float32(std_mult)).astype(np.float32)
    return vals, stds, t.astype(np.float64)


def curve(c, l=256, powers=(2.2, 2.0, 2.4, 1.8)):
    x = torch.linspace(0, 1, l)
    return torch.stack([x ** powers[i % len(powers)] for i in range(c)]).to(torch.float32)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if v is None:
            continue
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


# ----------------------------------------------------------------------------------------------
def gen_forward():
    g = torch.Generator().manual_seed(11)
    for c, h, w in ((3, 5, 7), (3, 4, 6), (1, 3, 5), (2, 3, 5), (4, 3, 5)):
        n = 2
        x = torch.rand((n, c, h, w), generator=g) * 1.2 - 0.1          # outside [0,1] on purpose
        flat = x.view(-1)
        specials = torch.tensor([0.0, 1.0, 0.5, 1.0 / 255, 254.0 / 255, 1.5 / 255, 2.5 / 255, 0.5 / 255,
                                 -0.0, 1.0000001, 128.0 / 255, 65535.0 / 65535, 257.0 / 65535])
        flat[: len(specials)] = specials
        theta = curve(c)
        for mode, tag in ((InterpMode.LINEAR, "linear"), (InterpMode.LOOKUP, "lookup")):
            model = ICRFModelDirect(icrf=theta.clone(), interpolation_mode=mode)
            xi = x.clone().requires_grad_(mode is InterpMode.LINEAR)
            y = model(xi)
            dydx = None
            if mode is InterpMode.LINEAR:
                dydx = torch.autograd.grad(y, xi, torch.ones_like(y))[0]
            save(f"forward_{tag}_c{c}_w{w}", x=x, theta=theta, y=y, dydx=dydx)
    # every 8-bit and a sweep of 16-bit code values through LINEAR and LOOKUP with L=256
    k8 = torch.arange(256, dtype=torch.float32) / 255.0
    k16 = torch.arange(0, 65536, 7, dtype=torch.float32) / 65535.0
    for tag, x in (("u8", k8), ("u16", k16)):
        xs = x.view(1, 1, 1, -1).repeat(1, 3, 1, 1).contiguous()
        theta = curve(3)
        yl = ICRFModelDirect(icrf=theta.clone(), interpolation_mode=InterpMode.LINEAR)(xs)
        yn = ICRFModelDirect(icrf=theta.clone(), interpolation_mode=InterpMode.LOOKUP)(xs)
        save(f"forward_codes_{tag}", x=xs, theta=theta, y_linear=yl, y_lookup=yn)


def gen_catmull():
    """CATMULL mode (models/base.py:184-226) on the CPU (the reference's CUDA path fails at :218): value, autograd
    derivative wrt the image, and the table gradient for a fixed upstream."""
    g = torch.Generator().manual_seed(17)
    for c, h, w in ((3, 5, 7), (3, 4, 6), (1, 3, 5), (4, 3, 5)):
        x = torch.rand((2, c, h, w), generator=g) * 1.2 - 0.1
        flat = x.view(-1)
        specials = torch.tensor([0.0, 1.0, 0.5, 1.0 / 255, 254.0 / 255, 1.5 / 255, 2.5 / 255, 0.5 / 255, 253.5 / 255])
        flat[: len(specials)] = specials
        up = torch.randn(x.shape, generator=g)
        model = ICRFModelDirect(n_points=256, channels=c, interpolation_mode=InterpMode.CATMULL)
        with torch.no_grad():
            for k, p in enumerate(model.direct_params):
                p.copy_(curve(c)[k])
        model.update_icrf()
        xi = x.clone().requires_grad_(True)
        y = model(xi)
        (dydx,) = torch.autograd.grad(y, xi, torch.ones_like(y), retain_graph=True)
        (y * up).sum().backward()
        gtheta = torch.stack([p.grad for p in model.direct_params])
        save(f"forward_catmull_c{c}_w{w}", x=x, theta=curve(c), y=y, dydx=dydx, upstream=up, grad_theta=gtheta)


def gen_frame_stats():
    """WBOMeanVar (common/statistics.py:112-259) in fp32, weighted and unweighted, 3 batches; and
    compute_video_mean_and_std (inference/inferential_statistics.py:19-49) with and without a model."""
    from clair_torch.common.enums import VarianceMode
    from clair_torch.common.statistics import WBOMeanVar
    from clair_torch.inference.inferential_statistics import compute_video_mean_and_std
    g = torch.Generator().manual_seed(5)
    vals = torch.rand((11, 3, 6, 8), generator=g)
    wts = torch.rand((11, 3, 6, 8), generator=g) * 2.0
    out = dict(val=vals, weights=wts, bounds=np.array([0, 4, 9, 11]))
    for weighted in (True, False):
        h = WBOMeanVar(dim=0, variance_mode=VarianceMode.RELIABILITY_WEIGHTS)
        for a, b in ((0, 4), (4, 9), (9, 11)):
            h.update_values(vals[a:b], wts[a:b] if weighted else None)
        tag = "w" if weighted else "u"
        out[f"mean_{tag}"] = h.mean
        out[f"m2_{tag}"] = h.m2
        out[f"wsum_{tag}"] = h.sum_of_weights
        out[f"wsq_{tag}"] = h.sum_of_squared_weights
        out[f"var_rel_{tag}"] = h.variance()
    save("framestats_wbomeanvar", **out)
    vals2, _, t = make_stack(seed=61, n=9, c=3, h=8, w=12, bits=8, std_mult=None)
    tv = [torch.from_numpy(v) for v in vals2]
    theta = curve(3)
    for with_model in (False, True):
        model = ICRFModelDirect(icrf=theta.clone()) if with_model else None
        mean, sem = compute_video_mean_and_std(loader(tv, None, t, 4), "cpu", model)
        save(f"framestats_video_model{int(with_model)}", val=vals2, exposure=t, theta=theta if with_model else None,
             batch_size=4, mean=mean, sem=sem)


def _artefact_stub(base_cls, vals, stds, exposures):
    """In-memory stand-in for the file-matching artefact datasets: frame i of the main dataset matches artefact i."""
    class Stub(base_cls):
        def __init__(self):
            self.vals, self.stds, self.exposures = vals, stds, exposures

        def get_matching_artefact_images(self, reference_frame_settings_list):
            items = []
            for ref in reference_frame_settings_list:
                i = int(ref) % len(self.vals)
                items.append((i, self.vals[i].clone(), None if self.stds is None else self.stds[i].clone(),
                              {"exposure_time": float(self.exposures[i])}))
            return custom_collate(items)
    return Stub()


def gen_artefacts():
    """Dark-field (conditional blur of hot pixels) and flat-field corrections with their variance terms:
    inference/hdr_merge.py:76-92,117-126,131-153 and inference/linearization.py:50-57,73-91,108-130."""
    from clair_torch.datasets.image_dataset import DarkFieldArtefactMapDataset, FlatFieldArtefactMapDataset
    rng = np.random.default_rng(71)
    for name, kw, bs, use_dark, use_flat in (("dark", dict(seed=71, n=4, c=3, h=9, w=12, bits=8), 4, True, False),
                                             ("flat", dict(seed=72, n=4, c=3, h=9, w=12, bits=8), 4, False, True),
                                             ("both_2batches", dict(seed=73, n=5, c=3, h=8, w=11, bits=16), 3, True, True)):
        vals, stds, t = make_stack(**kw)
        n, c, h, w = vals.shape
        theta = curve(c)
        # dark frames: mostly ~0.01, a few hot pixels well above the 0.05 threshold, some near it
        dark = rng.uniform(0.0, 0.02, size=vals.shape).astype(np.float32)
        hot = rng.random(vals.shape) < 0.08
        dark[hot] = rng.uniform(0.03, 0.3, size=int(hot.sum())).astype(np.float32)
        dark_std = (0.1 * dark + 0.001).astype(np.float32)
        flat = rng.uniform(0.6, 1.0, size=(1, c, h, w)).astype(np.float32)
        flat_std = (0.02 * flat).astype(np.float32)
        tv = [torch.from_numpy(v) for v in vals]
        ts = [torch.from_numpy(x) for x in stds]
        dark_ds = _artefact_stub(DarkFieldArtefactMapDataset, [torch.from_numpy(d) for d in dark],
                                 [torch.from_numpy(d) for d in dark_std], t) if use_dark else None
        flat_ds = _artefact_stub(FlatFieldArtefactMapDataset, [torch.from_numpy(flat[0])], [torch.from_numpy(flat_std[0])],
                                 [1.0]) if use_flat else None
        model = ICRFModelDirect(icrf=theta.clone())
        rad, sig = compute_hdr_image(loader(tv, ts, t, bs), "cpu", model, gaussian_value_weights, flat_ds, None, dark_ds)
        lin, lsig = [], []
        for a, b, _ in linearize_dataset_generator(loader(tv, ts, t, 1), "cpu", model, flat_ds, None, dark_ds):
            lin.append(a.numpy())
            lsig.append(b.numpy())
        save(f"artefact_{name}", val=vals, std=stds, exposure=t, theta=theta, batch_size=bs,
             dark=dark if use_dark else None, dark_std=dark_std if use_dark else None,
             flat=flat if use_flat else None, flat_std=flat_std if use_flat else None,
             radiance=rad, sigma=sig, linearized=np.stack(lin), lin_sigma=np.stack(lsig))


def gen_hdr():
    cases = [
        # name, stack kwargs, batch_size, with model, weight_fn, with std
        ("c1small", dict(seed=1234, n=5, c=3, h=18, w=24, bits=8), 5, True, True, True),
        ("c1small_2batches", dict(seed=1234, n=5, c=3, h=18, w=24, bits=8), 3, True, True, True),
        ("c1small_3batches", dict(seed=77, n=7, c=3, h=10, w=14, bits=8), 3, True, True, True),
        ("w_not_div3", dict(seed=5, n=4, c=3, h=7, w=13, bits=8), 4, True, True, True),
        ("u16", dict(seed=6, n=6, c=3, h=9, w=12, bits=16), 6, True, True, True),
        ("nostd", dict(seed=7, n=5, c=3, h=8, w=12, bits=8, std_mult=None), 5, True, True, False),
        ("nomodel", dict(seed=8, n=5, c=3, h=8, w=12, bits=8), 5, False, True, True),
        ("unitweights", dict(seed=9, n=5, c=3, h=8, w=12, bits=8), 5, True, False, True),
        ("unitweights_2batches", dict(seed=9, n=5, c=3, h=8, w=12, bits=8), 2, True, False, True),
        ("mono", dict(seed=10, n=3, c=1, h=6, w=10, bits=8), 3, True, True, True),
    ]
    for name, kw, bs, with_model, with_w, with_std in cases:
        vals, stds, t = make_stack(**kw)
        theta = curve(kw["c"])
        model = ICRFModelDirect(icrf=theta.clone()) if with_model else None
        tv = [torch.from_numpy(v) for v in vals]
        ts = None if stds is None else [torch.from_numpy(s) for s in stds]
        mean, sigma = compute_hdr_image(loader(tv, ts, t, bs), "cpu", model,
                                        gaussian_value_weights if with_w else None)
        save(f"hdr_{name}", val=vals, std=stds, exposure=t, theta=theta if with_model else None,
             batch_size=bs, gaussian=int(with_w), radiance=mean, sigma=sigma)


def gen_linearize():
    for name, kw in (("u8", dict(seed=21, n=3, c=3, h=6, w=9, bits=8)),
                     ("u16_w7", dict(seed=22, n=2, c=3, h=5, w=7, bits=16)),
                     ("nostd", dict(seed=23, n=2, c=3, h=5, w=6, bits=8, std_mult=None))):
        vals, stds, t = make_stack(**kw)
        theta = curve(kw["c"])
        model = ICRFModelDirect(icrf=theta.clone())
        tv = [torch.from_numpy(v) for v in vals]
        ts = None if stds is None else [torch.from_numpy(s) for s in stds]
        lin, sig = [], []
        for a, b, _ in linearize_dataset_generator(loader(tv, ts, t, 1), "cpu", model):
            lin.append(a.numpy())
            sig.append(b.numpy())
        save(f"linearize_{name}", val=vals, std=stds, exposure=t, theta=theta,
             linearized=np.stack(lin), sigma=np.stack(sig))


def gen_linearity():
    base = dict(seed=31, n=6, c=3, h=12, w=16, bits=8)
    combos = []
    for rel in (True, False):
        for unc in (True, False):
            for with_std in (True, False):
                for with_model in (True, False):
                    combos.append((rel, unc, with_std, with_model))
    for rel, unc, with_std, with_model in combos:
        kw = dict(base)
        if not with_std:
            kw["std_mult"] = None
        vals, stds, t = make_stack(**kw)
        theta = curve(3)
        model = ICRFModelDirect(icrf=theta.clone()) if with_model else None
        tv = [torch.from_numpy(v) for v in vals]
        ts = None if stds is None else [torch.from_numpy(s) for s in stds]
        ratio, m, sd, err = measure_linearity(loader(tv, ts, t, len(tv)), "cpu", unc, rel, model)
        save(f"linearity_rel{int(rel)}_unc{int(unc)}_std{int(with_std)}_model{int(with_model)}",
             val=vals, std=stds, exposure=t, theta=theta if with_model else None,
             ratio=ratio, mean=m, stddev=sd, errmean=err)
    # 16-bit, W not divisible by C, uneven exposure spacing
    vals, stds, t = make_stack(seed=32, n=5, c=3, h=9, w=11, bits=16)
    t = t * np.array([1.0, 1.1, 0.9, 1.3, 1.0])
    theta = curve(3)
    tv = [torch.from_numpy(v) for v in vals]
    ts = [torch.from_numpy(s) for s in stds]
    ratio, m, sd, err = measure_linearity(loader(tv, ts, t, len(tv)), "cpu", True, True,
                                          ICRFModelDirect(icrf=theta.clone()))
    save("linearity_u16_w11", val=vals, std=stds, exposure=t, theta=theta, ratio=ratio, mean=m, stddev=sd,
         errmean=err)


def train_step_reference(model, optimizers, images, stds, exposures, *, rel, unc, alpha, beta, gamma, delta,
                         lo=1 / 255, hi=254 / 255, thr=0.1):
    """Body of clair_torch/training/icrf_training.py:105-156, using the reference's own functions."""
    i_idx, j_idx, ratio_pairs = get_valid_exposure_pairs(increasing_exposure_values=exposures,
                                                         exposure_ratio_threshold=thr)
    valid_mask = get_pairwise_valid_pixel_mask(images, i_idx, j_idx, stds, val_lower=lo, val_upper=hi)
    gaussian_weight = combined_gaussian_pair_weights(images, i_idx, j_idx)
    for optimizer in optimizers:
        optimizer.zero_grad()
    images.requires_grad_(True)
    linearized = model(images)
    if stds is not None:
        grads = torch.autograd.grad(outputs=linearized, inputs=images, grad_outputs=torch.ones_like(linearized),
                                    retain_graph=True)[0]
        linearized_stds = (grads * stds).abs()
    else:
        linearized_stds = None
    icrf_curve = model.icrf
    pixelwise_loss, pixelwise_errors = pixelwise_linearity_loss(linearized, i_idx, j_idx, ratio_pairs,
                                                                linearized_stds, rel)
    spatial, _, _ = compute_spatial_linearity_loss(pixelwise_loss, pixelwise_errors, gaussian_weight, valid_mask,
                                                   unc)
    linearity_loss = torch.sqrt((spatial ** 2).sum(dim=0))
    loss = (linearity_loss + alpha * compute_monotonicity_penalty(icrf_curve, per_channel=True)
            + beta * compute_range_penalty(icrf_curve, per_channel=True)
            + gamma * compute_endpoint_penalty(icrf_curve, per_channel=True)
            + delta * compute_smoothness_penalty(icrf_curve, per_channel=True))
    if len(optimizers) == 1:
        loss = torch.sum(loss)
    connected = model.icrf.requires_grad
    if connected:
        for c, _ in enumerate(optimizers):
            loss[c].backward(retain_graph=True)
    grads_theta = None
    if connected:
        grads_theta = torch.stack([p.grad.clone() for p in model.direct_params])
    for optimizer in optimizers:
        optimizer.step()
    model.update_icrf()
    return loss.detach(), spatial.detach(), linearity_loss.detach(), grads_theta


def gen_trainstep():
    for name, kw, rel, unc, thr, coeffs, with_std, power in (
            ("script", dict(seed=2345, n=6, c=3, h=16, w=24, bits=8), True, False, 0.25, (10.0, 1.0, 1.0, 1.0),
             True, 2.5),
            ("defaults", dict(seed=41, n=5, c=3, h=12, w=18, bits=8), True, True, 0.1, (1.0, 1.0, 1.0, 1.0),
             True, 2.5),
            ("absolute", dict(seed=42, n=5, c=3, h=12, w=18, bits=8), False, True, 0.1, (1.0, 1.0, 1.0, 1.0),
             True, 2.5),
            ("absolute_nounc", dict(seed=43, n=4, c=3, h=10, w=13, bits=16), False, False, 0.1,
             (1.0, 1.0, 1.0, 1.0), True, 1.7),
            ("nostd", dict(seed=44, n=5, c=3, h=12, w=18, bits=8, std_mult=None), True, True, 0.1,
             (1.0, 1.0, 1.0, 1.0), False, 2.5),
            ("u16", dict(seed=45, n=5, c=3, h=11, w=14, bits=16), True, True, 0.1, (1.0, 1.0, 1.0, 1.0), True,
             1.9)):
        vals, stds, t = make_stack(**kw)
        alpha, beta, gamma, delta = coeffs
        model = ICRFModelDirect(n_points=256, channels=kw["c"], initial_power=power)
        # distinct rows so the k-mod-C channel striping is visible in the gradient
        with torch.no_grad():
            for c, p in enumerate(model.direct_params):
                p.copy_(torch.linspace(0, 1, 256) ** (power + 0.15 * c))
        model.update_icrf()
        theta0 = model.icrf.detach().clone()
        optimizers = [torch.optim.Adam(model.channel_params(c), lr=1e-3, amsgrad=False) for c in range(kw["c"])]
        images = torch.from_numpy(vals)
        tstd = None if stds is None else torch.from_numpy(stds)
        exposures = torch.from_numpy(t)
        out = {}
        n_steps = 4
        for step in range(n_steps):
            loss, spatial, lin, gtheta = train_step_reference(
                model, optimizers, images.clone(), tstd, exposures, rel=rel, unc=unc, alpha=alpha, beta=beta,
                gamma=gamma, delta=delta, thr=thr)
            out[f"loss_{step}"] = loss
            out[f"spatial_{step}"] = spatial
            out[f"linloss_{step}"] = lin
            out[f"grad_theta_{step}"] = gtheta
            out[f"theta_after_{step}"] = model.icrf.detach().clone()
        save(f"trainstep_{name}", val=vals, std=stds, exposure=t, theta0=theta0, rel=int(rel), unc=int(unc),
             thr=thr, coeffs=np.array(coeffs), n_steps=n_steps, **out)


def gen_modes():
    """The drivers with LOOKUP / CATMULL models (clair_torch/models/base.py:138-158, :184-226).  LOOKUP has no autograd
    edge to the image: with std images the linearisation / linearity drivers raise, so those cases run without std; the
    HDR merge still differentiates through the Gaussian weights."""
    modes = {"lookup": InterpMode.LOOKUP, "catmull": InterpMode.CATMULL}
    # compute_hdr_image
    for mname, kw, bs, with_w, with_std in (
            ("lookup", dict(seed=101, n=5, c=3, h=10, w=14, bits=8), 5, True, True),
            ("lookup", dict(seed=102, n=5, c=3, h=8, w=11, bits=8), 2, True, True),
            ("lookup", dict(seed=103, n=4, c=3, h=8, w=12, bits=8, std_mult=None), 4, False, False),
            ("catmull", dict(seed=104, n=5, c=3, h=10, w=14, bits=8), 5, True, True),
            ("catmull", dict(seed=105, n=6, c=3, h=9, w=13, bits=16), 4, True, True),
            ("catmull", dict(seed=106, n=4, c=3, h=8, w=12, bits=8), 4, False, True),
            ("catmull", dict(seed=107, n=3, c=1, h=6, w=10, bits=8), 3, True, True)):
        vals, stds, t = make_stack(**kw)
        theta = curve(kw["c"])
        model = ICRFModelDirect(icrf=theta.clone(), interpolation_mode=modes[mname])
        tv = [torch.from_numpy(v) for v in vals]
        ts = None if stds is None else [torch.from_numpy(s) for s in stds]
        mean, sigma = compute_hdr_image(loader(tv, ts, t, bs), "cpu", model, gaussian_value_weights if with_w else None)
        save(f"hdr_{mname}_s{kw['seed']}", val=vals, std=stds, exposure=t, theta=theta, batch_size=bs, gaussian=int(with_w),
             radiance=mean, sigma=sigma, mode=mname)
    # linearize_dataset_generator
    for mname, kw in (("lookup", dict(seed=111, n=2, c=3, h=5, w=7, bits=8, std_mult=None)),
                      ("catmull", dict(seed=112, n=3, c=3, h=6, w=9, bits=8)),
                      ("catmull", dict(seed=113, n=2, c=3, h=5, w=7, bits=16)),
                      ("catmull", dict(seed=114, n=2, c=3, h=5, w=6, bits=8, std_mult=None))):
        vals, stds, t = make_stack(**kw)
        theta = curve(kw["c"])
        model = ICRFModelDirect(icrf=theta.clone(), interpolation_mode=modes[mname])
        tv = [torch.from_numpy(v) for v in vals]
        ts = None if stds is None else [torch.from_numpy(s) for s in stds]
        lin, sig = [], []
        for a, b, _ in linearize_dataset_generator(loader(tv, ts, t, 1), "cpu", model):
            lin.append(a.numpy())
            sig.append(b.numpy())
        save(f"linearize_{mname}_s{kw['seed']}", val=vals, std=stds, exposure=t, theta=theta, linearized=np.stack(lin),
             sigma=np.stack(sig), mode=mname)
    # measure_linearity
    for mname, rel, unc, with_std in (("lookup", True, False, False), ("lookup", False, True, False),
                                      ("catmull", True, True, True), ("catmull", True, False, True),
                                      ("catmull", False, True, True), ("catmull", False, False, False)):
        kw = dict(seed=121, n=6, c=3, h=12, w=16, bits=8)
        if not with_std:
            kw["std_mult"] = None
        vals, stds, t = make_stack(**kw)
        theta = curve(3)
        model = ICRFModelDirect(icrf=theta.clone(), interpolation_mode=modes[mname])
        tv = [torch.from_numpy(v) for v in vals]
        ts = None if stds is None else [torch.from_numpy(s) for s in stds]
        ratio, m, sd, err = measure_linearity(loader(tv, ts, t, len(tv)), "cpu", unc, rel, model)
        save(f"linearity_{mname}_rel{int(rel)}_unc{int(unc)}_std{int(with_std)}", val=vals, std=stds, exposure=t, theta=theta,
             ratio=ratio, mean=m, stddev=sd, errmean=err, mode=mname)
    # train step, CATMULL (LOOKUP has no gradient to the table parameters' image edge and is not trainable with std)
    for name, kw, rel, unc, thr, coeffs, with_std, power in (
            ("catmull_script", dict(seed=131, n=6, c=3, h=14, w=20, bits=8), True, False, 0.25, (10.0, 1.0, 1.0, 1.0), True, 2.5),
            ("catmull_defaults", dict(seed=132, n=5, c=3, h=12, w=17, bits=8), True, True, 0.1, (1.0, 1.0, 1.0, 1.0), True, 2.2)):
        vals, stds, t = make_stack(**kw)
        alpha, beta, gamma, delta = coeffs
        model = ICRFModelDirect(n_points=256, channels=kw["c"], initial_power=power, interpolation_mode=InterpMode.CATMULL)
        with torch.no_grad():
            for c, p in enumerate(model.direct_params):
                p.copy_(torch.linspace(0, 1, 256) ** (power + 0.15 * c))
        model.update_icrf()
        theta0 = model.icrf.detach().clone()
        optimizers = [torch.optim.Adam(model.channel_params(c), lr=1e-3, amsgrad=False) for c in range(kw["c"])]
        images, tstd, exposures = torch.from_numpy(vals), torch.from_numpy(stds), torch.from_numpy(t)
        out = {}
        n_steps = 3
        for step in range(n_steps):
            loss, spatial, lin, gtheta = train_step_reference(model, optimizers, images.clone(), tstd, exposures, rel=rel, unc=unc,
                                                              alpha=alpha, beta=beta, gamma=gamma, delta=delta, thr=thr)
            out[f"loss_{step}"] = loss
            out[f"spatial_{step}"] = spatial
            out[f"linloss_{step}"] = lin
            out[f"grad_theta_{step}"] = gtheta
            out[f"theta_after_{step}"] = model.icrf.detach().clone()
        save(f"trainstep_{name}", val=vals, std=stds, exposure=t, theta0=theta0, rel=int(rel), unc=int(unc), thr=thr,
             coeffs=np.array(coeffs), n_steps=n_steps, mode="catmull", **out)


def gen_known_answers():
    out = {}
    # tests/unit/common/test_general_functions.py:292-319
    e = torch.tensor([1.0, 2.0, 4.0])
    i, j, r = get_valid_exposure_pairs(e)
    out.update(pairs_exposure=e, pairs_i=i, pairs_j=j, pairs_r=r)
    i, j, r = get_valid_exposure_pairs(e, exposure_ratio_threshold=0.4)
    out.update(pairs_thr=0.4, pairs_thr_i=i, pairs_thr_j=j, pairs_thr_r=r)
    e6 = torch.tensor([1e-3 * 2 ** k for k in range(6)], dtype=torch.float64)
    i, j, r = get_valid_exposure_pairs(e6, exposure_ratio_threshold=0.2)
    out.update(pairs6_exposure=e6, pairs6_thr=0.2, pairs6_i=i, pairs6_j=j, pairs6_r=r)
    # tests/unit/common/test_general_functions.py:332-391
    stack = torch.tensor([[[[0.1, 0.5], [0.9, 1.0]]], [[[0.2, 0.6], [0.4, 0.8]]], [[[0.05, 0.95], [1.1, 0.0]]]])
    ii, jj = torch.tensor([0, 0, 1]), torch.tensor([1, 2, 2])
    out.update(mask_stack=stack, mask_i=ii, mask_j=jj, mask_lo=0.1, mask_hi=1.0,
               mask_expected=get_pairwise_valid_pixel_mask(stack, ii, jj, val_lower=0.1, val_upper=1.0))
    # 16-bit validity under the training thresholds [1/255, 254/255]
    k16 = (torch.arange(65536, dtype=torch.float32) / 65535.0).view(1, 1, 1, -1)
    two = torch.cat([k16, torch.full_like(k16, 0.5)])
    m16 = get_pairwise_valid_pixel_mask(two, torch.tensor([0]), torch.tensor([1]), val_lower=1 / 255,
                                        val_upper=254 / 255)
    out.update(valid16=m16.view(-1))
    k8 = (torch.arange(256, dtype=torch.float32) / 255.0).view(1, 1, 1, -1)
    two = torch.cat([k8, torch.full_like(k8, 0.5)])
    m8 = get_pairwise_valid_pixel_mask(two, torch.tensor([0]), torch.tensor([1]), val_lower=1 / 255,
                                       val_upper=254 / 255)
    out.update(valid8=m8.view(-1))
    # tests/unit/common/test_general_functions.py:109-174
    v = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    w = torch.tensor([[1.0, 1.0], [0.25, 0.75]])
    out.update(wm_values=v, wm_weights=w, wm_weighted=weighted_mean_and_std(v, w, None, dim=1)[0],
               wm_masked=weighted_mean_and_std(v, None, torch.tensor([[True, False], [False, True]]), dim=1)[0])
    # tests/unit/common/test_statistics.py:31-78 : WBOMean equals the direct weighted mean
    g = torch.Generator().manual_seed(3)
    vals = torch.rand((12, 4), generator=g, dtype=torch.float64)
    wts = torch.rand((12, 4), generator=g, dtype=torch.float64)
    h = WBOMean(dim=0)
    for a in range(0, 12, 5):
        h.update_values(vals[a:a + 5], wts[a:a + 5])
    out.update(wbo_values=vals, wbo_weights=wts, wbo_mean=h.mean, wbo_wsum=h.sum_of_weights)
    # tests/unit/training/test_losses.py:12-89
    x = torch.linspace(-0.2, 1.2, 57)
    out.update(gw_x=x, gw_30=gaussian_value_weights(x), gw_10=gaussian_value_weights(x, 10.0))
    save("known_answers", **out)


def gen_ingest():
    """The transform chain in front of the hot path: a camera buffer (H, W, 3) BGR of integer codes through the
    reference's CvToTorch -> CastTo(float32) -> Normalize(max_val, min_val=0) (common/transforms.py), and the std image its datasets
    synthesise when no std file exists (datasets/base.py:35,128-133 — the two expressions are repeated here because the dataset
    class only runs on files)."""
    from clair_torch.common.transforms import CastTo, CvToTorch, Normalize
    rng = np.random.default_rng(2024)
    for bits, dtype, tdtype in ((8, np.uint8, torch.uint8), (16, np.uint16, torch.uint16)):
        for max_val in ((255,) if bits == 8 else (65535, 4095)):
            n_codes = max_val + 1                         # every code of the range, in each channel, in a different order
            side = {256: 32, 4096: 64, 65536: 256}[n_codes]
            planes = [rng.permutation(np.arange(side * side) % n_codes).reshape(side, side) for _ in range(3)]
            camera = np.stack(planes, axis=-1).astype(dtype)  # (H, W, 3), OpenCV's B, G, R
            x = torch.from_numpy(camera.astype(np.int32)).to(tdtype) if bits == 16 else torch.from_numpy(camera)
            for tf in (CvToTorch(), CastTo(data_type=torch.float32), Normalize(max_val=max_val, min_val=0)):
                x = tf(x)
            shared_std_tensor = torch.tensor(0.05)        # datasets/base.py:35
            std_mult = x * shared_std_tensor              # :133
            std_const = torch.tensor(0.01).expand_as(x)   # :131
            save(f"ingest_u{bits}_max{max_val}", camera=camera, max_val=np.float64(max_val), val=x.contiguous(),
                 std_multiplier=std_mult.contiguous(), std_constant=std_const.contiguous(), multiplier=np.float64(0.05),
                 constant=np.float64(0.01))


if __name__ == "__main__":
    generators = {"known_answers": gen_known_answers, "forward": gen_forward, "catmull": gen_catmull, "frame_stats": gen_frame_stats,
                  "artefacts": gen_artefacts, "hdr": gen_hdr, "linearize": gen_linearize, "linearity": gen_linearity,
                  "trainstep": gen_trainstep, "modes": gen_modes, "ingest": gen_ingest}
    for key in (sys.argv[1:] or list(generators)):       # e.g. `make_golden.py modes` regenerates one family only
        generators[key]()
