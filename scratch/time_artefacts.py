"""Artefact kernels at c1 size: dark mix pre-pass, flat field, frame stats (CUDA events, rotated inputs)."""
import numpy as np, torch
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
N, C, H, W = 5, 3, 1080, 1920
sets = []
for k in range(3):
    val, std, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=10 + k, device=dev)
    dark = torch.rand_like(val) * 0.1
    dstd = torch.rand_like(val) * 0.01
    sets.append((val, std, dark, dstd))
stream = torch.cuda.current_stream(dev)
def timed(fn, warm=3, reps=30):
    for k in range(warm): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(reps): fn(k)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = timed(lambda k: kernels.dark_field_mix(*sets[k % 3]))
print(f"dark mix: {ms*1e3:.1f} us  ({N*C*H*W*24/ms/1e6:.0f} GB/s)")
theta = ct.synthetic.reference_curve(C).to(dev)
for dt in (torch.float32, torch.float64):
    rads = [torch.rand((C, H, W), device=dev, dtype=dt) for _ in range(6)]
    sigs = [torch.rand((C, H, W), device=dev) * 0.01 for _ in range(6)]
    flat = torch.rand((C, H, W), device=dev) * 0.5 + 0.5
    fstd = torch.rand((C, H, W), device=dev) * 0.01
    ms = timed(lambda k: kernels.flat_field_correct_(rads[k % 6], sigs[k % 6], flat, fstd, True))
    b = C*H*W*(2*rads[0].element_size() + 8 + 8 + rads[0].element_size() + 4)
    print(f"flat field {dt}: {ms*1e3:.1f} us  ({b/ms/1e6:.0f} GB/s incl. the reduce pass reads)")
# fused dark merge vs pre-pass + merge
t_host = np.ascontiguousarray(1e-3 * 2.0 ** np.arange(N))
def fused(k):
    v, s, d, ds = sets[k % 3]
    kernels.hdr_merge_update(kernels.HdrMergeState(), v, s, t_host, theta, True, True, radiance_dtype=torch.float32, dark=(d, ds))
def two_pass(k):
    v, s, d, ds = sets[k % 3]
    m, se = kernels.dark_field_mix(v, s, d, ds)
    kernels.hdr_merge_update(kernels.HdrMergeState(), m, se, t_host, theta, True, True, radiance_dtype=torch.float32)
print(f"dark-field merge fused: {timed(fused)*1e3:.1f} us;  pre-pass + merge: {timed(two_pass)*1e3:.1f} us")
