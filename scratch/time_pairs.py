"""Kernel-only timing of the pair kernels (CUDA events) on c2 / c3 / c5 shapes under CLAIR_TUNE knobs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
lib = ct._native.load()
for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):
    k, _, v = item.partition("=")
    ct._native.check(lib.clair_set_tuning(k.encode(), int(v)), "tune")
dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["c2", "c3", "c5"]

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

theta = ct.synthetic.reference_curve(3).to(dev)
tag = os.environ.get("CLAIR_TUNE", "default")
for w in which:
    if w == "c2":
        val, std, t = ct.synthetic.make_stack(10, 3, 1080, 1920, bits=8, seed=2345, device=dev); thr = 0.25
    elif w == "c3":
        val, std, t = ct.synthetic.make_stack(16, 3, 2160, 3840, bits=16, seed=3456, device=dev); thr = 0.2
    else:
        val, std, t = ct.synthetic.make_stack(2, 3, 8192, 12288, bits=16, seed=5678, device=dev); thr = 0.25
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), thr)
    full = timed(lambda: kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True), 10)
    means = timed(lambda: kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, means_only=True))
    sums = kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, means_only=True)
    _, _, up, mg = kernels.pair_upstream(sums)
    grad = timed(lambda: kernels.pair_grad(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, up, mg))
    gradu = timed(lambda: kernels.pair_grad(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True, up, mg), 10)
    print(f"{tag:40s} {w} P={len(i):2d}: stats(full,unc) {full:7.3f} ms  means {means:7.3f} ms  grad {grad:7.3f} ms  grad(unc) {gradu:7.3f} ms")
    del val, std
    torch.cuda.empty_cache()
