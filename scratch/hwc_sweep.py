import time, numpy as np, torch
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from clair_torch_b200.datasets import StdSpec
dev = torch.device("cuda", 0)
N, C, H, W = 5, 3, 1080, 1920
val, _, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=1, device=dev)
c8 = torch.round(val * 255).to(torch.uint8)
cam_h = torch.stack([c8[:, 2], c8[:, 1], c8[:, 0]], dim=-1).contiguous().cpu().pin_memory()
pl_h = c8.cpu().pin_memory()
theta = ct.synthetic.reference_curve(C).to(dev)
rad_h = torch.empty((C, H, W), dtype=torch.float32).pin_memory(); sig_h = torch.empty_like(rad_h).pin_memory()
spec = StdSpec("multiplier", 0.05)
def run(x, layout, staged, bands):
    def once():
        kernels.hdr_merge_update(kernels.HdrMergeState(), x, spec, t, theta, True, True, radiance_dtype=torch.float32, device=dev,
                                 host_out=(rad_h, sig_h), code_layout=layout, staged=staged, bands=bands)
    for _ in range(3): once()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): once()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 20 * 1e3
print("planar zero-copy", run(pl_h, "planar", False, 1))
for b in (1, 2, 3, 4, 6, 8, 16):
    print("planar staged", b, run(pl_h, "planar", True, b), "  hwc staged", b, run(cam_h, "hwc_bgr", True, b))
print("hwc zero-copy", run(cam_h, "hwc_bgr", False, 1))
