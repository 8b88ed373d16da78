"""Launch list driver: the c5 training step on ONE rank's band of an 8-GPU run (1024 of 8192 rows), eager, capturable fused Adam."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import distributed as cd
dev = torch.device("cuda", 0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
val, std, _ = ct.synthetic.make_stack(2, 3, rows, 12288, bits=16, seed=5678, device=dev)
exposures = torch.tensor([0.01, 0.02], dtype=torch.float64)
rb = cd.band_row_base(3, 8192, 12288, 0)
model = ct.ICRFModelDirect(256, 3, ct.InterpMode.LINEAR, 2.5).to(dev)
opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3, capturable=True, fused=True) for c in range(3)]
kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=False, alpha=10.0, beta=1.0, gamma=1.0, delta=1.0,
          exposure_ratio_threshold=0.25)
for _ in range(4):
    cd.train_icrf_step_data_parallel(model, opts, val, std, exposures, rb, **kw)
torch.cuda.synchronize()
g = cd.graphed_train_step_data_parallel(model, opts, val, std, exposures, rb, **kw)
for _ in range(5): g()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): g()
b.record(); torch.cuda.synchronize()
print("graph replay ms/step", a.elapsed_time(b) / 50)
