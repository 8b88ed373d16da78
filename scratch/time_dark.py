"""Fused dark merge, dark pre-pass and pre-pass + merge at c1 size (CUDA events, three rotating input sets)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
lib = ct._native.load()
for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):
    k, _, v = item.partition("=")
    ct._native.check(lib.clair_set_tuning(k.encode(), int(v)), "tune")
N, C, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "5x3x1080x1920").split("x"))
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
def prepass(k):
    v, s, d, ds = sets[k % 3]
    kernels.dark_field_mix(v, s, d, ds)
for r in range(2):
    print(f"{N}x{C}x{H}x{W}: fused dark merge {timed(fused)*1e3:.1f} us   dark pre-pass {timed(prepass)*1e3:.1f} us", flush=True)
