"""End-to-end c4 from pinned uint16 codes: zero-copy vs staged (bands), results written by the kernel to pinned host memory
or to device memory and copied back by the second copy engine; plus the bare copy-engine ceiling of the same traffic."""
import os, time, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from clair_torch_b200.datasets import StdSpec
dev = torch.device("cuda", 0)
N, C, H, W = 9, 3, 4000, 6000
val, _, t = ct.synthetic.make_stack(N, C, H, W, bits=16, seed=4567, device=dev)
c16 = torch.round(val * 65535).to(torch.int32).to(torch.uint16)
del val
pl_h = c16.cpu().pin_memory()
cam_h = torch.stack([c16[:, 2], c16[:, 1], c16[:, 0]], dim=-1).contiguous().cpu().pin_memory()
del c16
theta = ct.synthetic.reference_curve(C).to(dev)
rad_h = torch.empty((C, H, W), dtype=torch.float32).pin_memory(); sig_h = torch.empty_like(rad_h).pin_memory()
spec = StdSpec("multiplier", 0.05)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def run(x, layout, staged, bands):
    return timed(lambda: kernels.hdr_merge_update(kernels.HdrMergeState(), x, spec, t, theta, True, True, radiance_dtype=torch.float32,
                 device=dev, host_out=(rad_h, sig_h), code_layout=layout, staged=staged, bands=bands))
# ceiling: the two copy engines alone, H2D of the codes and D2H of the results at the same time
d_in = torch.empty_like(pl_h, device=dev); d_out = torch.empty((2, C, H, W), dtype=torch.float32, device=dev)
h_out = torch.empty((2, C, H, W), dtype=torch.float32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_in.copy_(pl_h, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(pl_h, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
print("copy engines: H2D alone %.2f ms, D2H alone %.2f ms, both at once %.2f ms" % (timed(h2d), timed(d2h), timed(both)))
del d_in, d_out, h_out
print("planar zero-copy          %.2f ms" % run(pl_h, "planar", False, 1))
for b in (2, 4, 8, 16, 32):
    print("bands %2d: planar staged %.2f ms   hwc staged %.2f ms" % (b, run(pl_h, "planar", True, b), run(cam_h, "hwc_bgr", True, b)))
