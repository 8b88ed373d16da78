"""Dark pre-pass, flat field, frame statistics, code expansion at c1 / c3 size under the aux_waves knob (CUDA events)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
from clair_torch_b200.common.statistics import WBOMeanVar
lib = ct._native.load()
dev = torch.device("cuda", 0)
N, C, H, W = 5, 3, 1080, 1920
val, std, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=99, device=dev)
dark = torch.rand_like(val) * 0.06
dark_std = dark * 0.1 + 1e-3
theta = ct.synthetic.reference_curve(C).to(dev)
radiance, sigma = kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True, radiance_dtype=torch.float32)
flat = torch.rand_like(radiance) * 0.4 + 0.6
flat_std = flat * 0.02
codes = torch.randint(0, 65536, (16, 3, 2160, 3840), dtype=torch.int32, device=dev).to(torch.uint16)
from clair_torch_b200.datasets import StdSpec
def timed(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for rnd in range(2):
    for knob in (1, 2, 4, 8, 16, 32):
        ct._native.check(lib.clair_set_tuning(b"aux_waves", knob), "tune")
        handler = WBOMeanVar(dim=0)
        d = timed(lambda: kernels.dark_field_mix(val, std, dark, dark_std))
        f = timed(lambda: kernels.flat_field_correct_(radiance, sigma, flat, flat_std, True))
        s = timed(lambda: handler.update_values(val, None, table=theta))
        e = timed(lambda: kernels.expand_codes(codes, StdSpec("multiplier", 0.05), 65535.0), 10)
        print(f"aux_waves={knob:3d}: dark pre-pass {d:6.1f} us  flat {f:5.1f} us  frame stats {s:5.1f} us  expand c3 codes {e:6.1f} us", flush=True)
