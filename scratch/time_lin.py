"""Device-resident lineariser (3 frames of 24 MP, f(x) and sigma) and forward pass under the fwd_blocks knob."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
lib = ct._native.load()
dev = torch.device("cuda", 0)
shape = (3, 3, 4000, 6000) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
val, std, t = ct.synthetic.make_stack(*shape, bits=16, seed=4567, device=dev)
frames, mp = shape[0], shape[2] * shape[3] / 1e6
theta = ct.synthetic.reference_curve(3).to(dev)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ref = None
for rnd in range(2):
    for knob in (-1, 2, 3, 4, 6, 8, 12, 16):
        if lib.clair_set_tuning(b"fwd_blocks", knob) != 0:      # a library without the knob: one measurement
            if knob != -1: continue
        ms = timed(lambda: kernels.linearize(val, std, theta), 30) / frames
        out = kernels.linearize(val, std, theta)
        if ref is None: ref = [o.clone() for o in out]
        same = all(torch.equal(a, b) for a, b in zip(out, ref))
        frac = 3 * mp * 1e6 * 16 / (ms * 1e-3) / 1e9 / 6551.7
        print(f"fwd_blocks={knob:3d}: {ms:.4f} ms per frame, {frac:.3f} of the copy peak, identical to the one-tile form: {same}", flush=True)
