"""Small invocation of every kernel (for compute-sanitizer): odd sizes, every template family."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from clair_torch_b200.datasets import StdSpec
from clair_torch_b200.common.statistics import WBOMeanVar
from clair_torch_b200.training import linearity_loss_and_table_grad
dev = "cuda:0"
theta = ct.synthetic.reference_curve(3).to(dev)
only_wide = len(sys.argv) > 1 and sys.argv[1] == "wide"      # just the 9..16-frame integer ingest
for n, h, w in () if only_wide else ((5, 17, 23), (3, 16, 20), (9, 9, 14), (41, 6, 8)):
    val, std, _ = ct.synthetic.make_stack(n, 3, h, w, bits=16, seed=n)
    t = 1e-3 * 1.2 ** np.arange(n)
    v, s = val.to(dev), std.to(dev)
    kernels.hdr_merge_update(kernels.HdrMergeState(), v, s, t, theta, True, True)
    kernels.hdr_merge_update(kernels.HdrMergeState(), v, None, t, None, False, True)
    st = kernels.HdrMergeState()
    kernels.hdr_merge_update(st, v[: n // 2 + 1].contiguous(), s[: n // 2 + 1].contiguous(), t[: n // 2 + 1], theta, True, False)
    kernels.hdr_merge_update(st, v[n // 2 + 1:].contiguous(), s[n // 2 + 1:].contiguous(), t[n // 2 + 1:], theta, True, True) if n // 2 + 1 < n else None
    if (h * w) % 4 == 0:
        codes = torch.round(val * 65535).to(torch.uint16).to(dev)
        kernels.hdr_merge_update(kernels.HdrMergeState(), codes, StdSpec("multiplier", 0.05), t, theta, True, True)
        kernels.hdr_merge_update(kernels.HdrMergeState(), torch.round(val * 255).to(torch.uint8).to(dev), s, t, theta, True, True)
    kernels.linearize(v, s, theta)
    kernels.icrf_forward(v, theta, ct._native.INTERP_LOOKUP)
    y, d = kernels.icrf_forward(v, theta, ct._native.INTERP_CATMULL, want_derivative=True)
    kernels.icrf_backward_theta(v, y, 3, 256)
    kernels.icrf_backward_theta(v, y, 3, 256, interp_mode=ct._native.INTERP_CATMULL)
    if n <= 16:
        i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.0)
        for rel in (True, False):
            for unc in (True, False):
                kernels.pair_stats(v, s, i, j, r, theta, 1 / 255, 254 / 255, rel, unc)
                linearity_loss_and_table_grad(v, s, i, j, r, theta, 1 / 255, 254 / 255, rel, unc)
        kernels.pair_stats(v, None, i, j, r, None, 1 / 255, 254 / 255, True, True)
    dark = torch.rand_like(v) * 0.1
    kernels.dark_field_mix(v, s, dark, dark * 0.1)
    rad, sig = kernels.hdr_merge_update(kernels.HdrMergeState(), v, s, t, theta, True, True)
    kernels.flat_field_correct_(rad, sig, torch.rand(3, h, w) + 0.5, torch.rand(3, h, w) * 0.01, True)
    hm = WBOMeanVar()
    hm.update_values(v, None, table=theta); hm.update_values(v, torch.rand_like(v))
# integer ingest of 9..16 frames (the 2-code register kernel), planar and camera layout, exact-fit planes (no slack behind them)
for n, h, w in ((9, 10, 14), (12, 6, 6), (16, 2, 2), (13, 30, 44)):
    val, std, _ = ct.synthetic.make_stack(n, 3, h, w, bits=16, seed=n)
    t = 1e-3 * 1.2 ** np.arange(n)
    for scale, dt in ((65535, torch.uint16), (255, torch.uint8)):
        codes = torch.round(val * scale).to(dt)
        camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous().to(dev)
        for sp in (StdSpec("multiplier", 0.05), std.to(dev)):
            kernels.hdr_merge_update(kernels.HdrMergeState(), codes.to(dev), sp, t, theta, True, True)
            kernels.hdr_merge_update(kernels.HdrMergeState(), camera, sp, t, theta, True, True, code_layout="hwc_bgr")
pen_grad = torch.zeros((3, 256), dtype=torch.float64, device=dev)
kernels.curve_penalties(theta, 1.0, 1.0, 1.0, 1.0, pen_grad)
torch.cuda.synchronize()
print("all kernels ran")
