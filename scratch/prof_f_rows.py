"""Profiling driver for the 8(f) kernels at c1 size: frame statistics, flat field, fused dark merge (few launches each)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from clair_torch_b200.common.statistics import WBOMeanVar
dev = torch.device("cuda", 0)
N, C, H, W = 5, 3, 1080, 1920
val, std, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=99, device=dev)
theta = ct.synthetic.reference_curve(C).to(dev)
dark = torch.rand_like(val) * 0.06
dark_std = dark * 0.1 + 1e-3
h = WBOMeanVar(dim=0)
for _ in range(3):
    h.update_values(val, None, table=theta)
rad, sig = kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True, radiance_dtype=torch.float32)
flat = torch.rand_like(rad) * 0.4 + 0.6
fstd = flat * 0.02
for _ in range(2):
    kernels.flat_field_correct_(rad, sig, flat, fstd, True)
for _ in range(2):
    kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True, radiance_dtype=torch.float32, dark=(dark, dark_std))
torch.cuda.synchronize()
print("ok")
