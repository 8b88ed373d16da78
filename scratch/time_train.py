"""Time the c2 training step, eager and as a replayed CUDA graph, under tuning knobs (CLAIR_TUNE env)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
lib = ct._native.load()
for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):
    k, _, v = item.partition("=")
    ct._native.check(lib.clair_set_tuning(k.encode(), int(v)), "tune")
dev = torch.device("cuda", 0)
val, std, t = ct.synthetic.make_stack(10, 3, 1080, 1920, bits=8, seed=2345, device=dev)
exposures = torch.from_numpy(t)

def timed(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for unc, thr in ((False, 0.25), (True, 0.1)):
    model = ct.ICRFModelDirect(256, 3, ct.InterpMode.LINEAR, 2.5).to(dev)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3, capturable=True, fused=True) for c in range(3)]
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=unc, alpha=10.0, beta=1.0, gamma=1.0, delta=1.0,
              exposure_ratio_threshold=thr)
    eager = timed(lambda: ct.train_icrf_step(model, opts, val, std, exposures, **kw))
    step = ct.GraphedTrainStep(model, opts, val, std, exposures, **kw)
    graphed = timed(step)
    print(os.environ.get("CLAIR_TUNE", ""), "unc" if unc else "script", f"eager {eager:.3f} ms/step, graph replay {graphed:.3f} ms/step")
