"""ncu driver: c3 statistics (FULL), c2 means + gradient, c5 single-pass step — one launch each after a warm-up."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
theta = ct.synthetic.reference_curve(3).to(dev)
val, std, t = ct.synthetic.make_stack(16, 3, 2160, 3840, bits=16, seed=3456, device=dev)
i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.2)
for _ in range(2):
    kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True)
del val, std
val, std, t = ct.synthetic.make_stack(10, 3, 1080, 1920, bits=8, seed=2345, device=dev)
i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.25)
for _ in range(2):
    sums = kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, means_only=True)
    _, _, up, mg = kernels.pair_upstream(sums)
    kernels.pair_grad(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, up, mg)
del val, std
val, std, _ = ct.synthetic.make_stack(2, 3, 8192, 12288, bits=16, seed=5678, device=dev)
i, j, r = ct.common.get_valid_exposure_pairs(torch.tensor([0.01, 0.02], dtype=torch.float64), 0.25)
for _ in range(2):
    kernels.pair_fused(val, i, j, r, theta, 1 / 255, 254 / 255, True)
torch.cuda.synchronize()
print("ok")
