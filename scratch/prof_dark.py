import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
N, C, H, W = 5, 3, 1080, 1920
val, std, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=10, device=dev)
dark = torch.rand_like(val) * 0.1
dstd = torch.rand_like(val) * 0.01
theta = ct.synthetic.reference_curve(C).to(dev)
t_host = np.ascontiguousarray(1e-3 * 2.0 ** np.arange(N))
for _ in range(3):
    kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t_host, theta, True, True, radiance_dtype=torch.float32, dark=(dark, dstd))
    kernels.dark_field_mix(val, std, dark, dstd)
torch.cuda.synchronize()
print("ok")
