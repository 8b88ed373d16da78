"""Profiling driver: c5 data-parallel training step on one GPU (100.7 MP 16-bit exposure pair), few iterations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import distributed as cd
dev = torch.device("cuda", 0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
height, width = 8192, 12288
val, std, _ = ct.synthetic.make_stack(n_frames, 3, height, width, bits=16, seed=5678, device=dev)
exposures = torch.tensor([0.01 * 2 ** k for k in range(n_frames)], dtype=torch.float64)
rb = cd.band_row_base(3, height, width, 0)
model = ct.ICRFModelDirect(256, 3, ct.InterpMode.LINEAR, 2.5).to(dev)
opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3) for c in range(3)]
kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=False, alpha=10.0, beta=1.0, gamma=1.0, delta=1.0,
          exposure_ratio_threshold=0.25)
for _ in range(iters):
    loss = cd.train_icrf_step_data_parallel(model, opts, val, std, exposures, rb, **kw)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    loss = cd.train_icrf_step_data_parallel(model, opts, val, std, exposures, rb, **kw)
b.record(); torch.cuda.synchronize()
print("ms/step", a.elapsed_time(b) / iters, "loss", loss.tolist())
