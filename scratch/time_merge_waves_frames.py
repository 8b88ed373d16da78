import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
lib = ct._native.load()
dev = torch.device("cuda", 0)
theta = ct.synthetic.reference_curve(3).to(dev)
def timed(fn, reps):
    for k in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for n in (9, 10, 12, 14, 16):
    v, s, t = ct.synthetic.make_stack(n, 3, 4000, 6000, bits=16, seed=1, device=dev)
    res = {}
    knobs = [2, 3, 6]
    for rnd in range(6):
        for knob in knobs[rnd % 3:] + knobs[:rnd % 3]:       # order rotated: the board's power state drifts within a round
            ct._native.check(lib.clair_set_tuning(b"hdr_waves", knob), "tune")
            res.setdefault(knob, []).append(timed(lambda: kernels.hdr_merge_update(kernels.HdrMergeState(), v, s, t, theta, True, True, radiance_dtype=torch.float32), 8))
    print(n, {k: round(sorted(x)[len(x) // 2], 1) for k, x in res.items()}, flush=True)
    del v, s
    torch.cuda.empty_cache()
