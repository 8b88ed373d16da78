"""Profiling driver: c2 training step and c3 linearity, few iterations (for ncu launch lists / full captures)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "train"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if mode == "train":
    val, std, t = ct.synthetic.make_stack(10, 3, 1080, 1920, bits=8, seed=2345, device=dev)
    exposures = torch.from_numpy(t)
    model = ct.ICRFModelDirect(256, 3, ct.InterpMode.LINEAR, 2.5).to(dev)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3) for c in range(3)]
    unc = len(sys.argv) > 3 and sys.argv[3] == "unc"
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=unc, alpha=10.0, beta=1.0, gamma=1.0, delta=1.0,
              exposure_ratio_threshold=0.1 if unc else 0.25)
    for _ in range(iters):
        loss = ct.train_icrf_step(model, opts, val, std, exposures, **kw)
    torch.cuda.synchronize()
    print("loss", loss.tolist())
else:
    val, std, t = ct.synthetic.make_stack(16, 3, 2160, 3840, bits=16, seed=3456, device=dev)
    theta = ct.synthetic.reference_curve(3).to(dev)
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.2)
    for _ in range(iters):
        sums = kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True)
    torch.cuda.synchronize()
    print("sums", sums[0, 0].tolist())
