"""Timing of the 8(f) kernels at c1 size the way bench.py reports them."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, clair_torch_b200 as ct
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ct._native.load()
x = bench.secondary_metrics(dev)
for k in ("artefacts_c1", "frame_stats_c1", "linearity_c3"):
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in x[k].items() if a not in ("note", "config")})
print("icrf_train_c2", x["icrf_train_c2"]["ms_per_step"])
