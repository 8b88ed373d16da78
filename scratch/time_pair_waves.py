"""Pair kernels under stats_waves / grad_waves, knob values interleaved over rounds (one process, one box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
lib = ct._native.load()
dev = torch.device("cuda", 0)
theta = ct.synthetic.reference_curve(3).to(dev)
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
shapes = {"c2": (10, 1080, 1920, 8, 2345, 0.25), "c3": (16, 2160, 3840, 16, 3456, 0.2), "c5": (2, 8192, 12288, 16, 5678, 0.25),
          "c5/8": (2, 1024, 12288, 16, 5678, 0.25), "c5/2": (2, 4096, 12288, 16, 5678, 0.25), "c5/4": (2, 2048, 12288, 16, 5678, 0.25),
          "8k": (2, 4320, 7680, 16, 99, 0.25), "4k8": (2, 2160, 3840, 8, 98, 0.25)}
knobs = (1, 2, 3, 4, 6, 8)[:int(os.environ.get('N_KNOBS', '6'))]
for name in sys.argv[1:] or list(shapes):
    n, h, w, bits, seed, thr = shapes[name]
    val, std, t = ct.synthetic.make_stack(n, 3, h, w, bits=bits, seed=seed, device=dev)
    i, j, r = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), thr)
    sums = kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, means_only=True)
    _, _, up, mg = kernels.pair_upstream(sums)
    res = {}
    for rnd in range(6):
        for k in knobs[rnd % len(knobs):] + knobs[:rnd % len(knobs)]:      # order rotated: the power state drifts within a round
            ct._native.check(lib.clair_set_tuning(b"stats_waves", k), "tune")
            ct._native.check(lib.clair_set_tuning(b"grad_waves", k), "tune")
            full = timed(lambda: kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, True))
            means = timed(lambda: kernels.pair_stats(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, means_only=True), 20)
            grad = timed(lambda: kernels.pair_grad(val, std, i, j, r, theta, 1 / 255, 254 / 255, True, False, up, mg), 20)
            fused = timed(lambda: kernels.pair_fused(val, i, j, r, theta, 1 / 255, 254 / 255, True), 20) if len(i) == 1 else float("nan")
            res.setdefault(k, []).append((full, means, grad, fused))
    for k in knobs:
        med = [sorted(v[q] for v in res[k])[len(res[k]) // 2] for q in range(4)]
        print(f"{name:5s} waves={k}: stats {med[0]:.3f}  means {med[1]:.3f}  grad {med[2]:.3f}  fused {med[3]:.3f} ms", flush=True)
    del val, std
    torch.cuda.empty_cache()
