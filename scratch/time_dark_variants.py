import numpy as np, torch
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
lib = ct._native.load()
N, C, H, W = 5, 3, 1080, 1920
sets = []
for k in range(3):
    val, std, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=10 + k, device=dev)
    sets.append((val, std, torch.rand_like(val) * 0.1, torch.rand_like(val) * 0.01))
theta = ct.synthetic.reference_curve(C).to(dev)
t_host = np.ascontiguousarray(1e-3 * 2.0 ** np.arange(N))
stream = torch.cuda.current_stream(dev)
def timed(fn, warm=3, reps=30):
    for k in range(warm): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(reps): fn(k)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def fused(k):
    v, s, d, ds = sets[k % 3]
    kernels.hdr_merge_update(kernels.HdrMergeState(), v, s, t_host, theta, True, True, radiance_dtype=torch.float32, dark=(d, ds))
def two_pass(k):
    v, s, d, ds = sets[k % 3]
    m, se = kernels.dark_field_mix(v, s, d, ds)
    kernels.hdr_merge_update(kernels.HdrMergeState(), m, se, t_host, theta, True, True, radiance_dtype=torch.float32)
for ch in (1, 2, 3):
    lib.clair_set_tuning(b"grad_pix", ch)
    for waves in (1, 2, 3):
        lib.clair_set_tuning(b"hdr_waves", waves)
        print(f"chunk {ch} waves {waves}: fused {timed(fused)*1e3:.1f} us", flush=True)
lib.clair_set_tuning(b"hdr_waves", 0); lib.clair_set_tuning(b"grad_pix", 0)
print(f"pre-pass + merge: {timed(two_pass)*1e3:.1f} us")
