"""HDR merge c1 / c4 under the hdr_waves knob, knob values interleaved over several rounds (the board's power state drifts)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
lib = ct._native.load()
dev = torch.device("cuda", 0)
theta = ct.synthetic.reference_curve(3).to(dev)
def timed(fn, reps):
    for k in range(2): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps): fn(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c1":
    sets = [ct.synthetic.make_stack(5, 3, 1080, 1920, bits=8, seed=10 + k, device=dev) for k in range(6)]
    knobs, reps = (1, 2, 3), 120
    fn = lambda k: kernels.hdr_merge_update(kernels.HdrMergeState(), sets[k % 6][0], sets[k % 6][1], sets[0][2], theta, True, True, radiance_dtype=torch.float32)
else:
    v4, s4, t4 = ct.synthetic.make_stack(9, 3, 4000, 6000, bits=16, seed=4567, device=dev)
    knobs, reps = (2, 3, 6, 12, 24, 48), 20
    fn = lambda k: kernels.hdr_merge_update(kernels.HdrMergeState(), v4, s4, t4, theta, True, True, radiance_dtype=torch.float32)
res = {k: [] for k in knobs}
for rnd in range(5):
    for knob in knobs:
        ct._native.check(lib.clair_set_tuning(b"hdr_waves", knob), "tune")
        res[knob].append(timed(fn, reps))
for knob in knobs:
    print(f"{which} hdr_waves={knob:3d}: " + "  ".join(f"{v:7.1f}" for v in res[knob]) + f"   median {sorted(res[knob])[2]:7.1f} us", flush=True)
