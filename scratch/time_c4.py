import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
lib = ct._native.load()
for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):
    k, _, v = item.partition("="); lib.clair_set_tuning(k.encode(), int(v))
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9
val, std, t = ct.synthetic.make_stack(n, 3, 4000, 6000, bits=16, seed=4567, device=dev)
theta = ct.synthetic.reference_curve(3).to(dev)
def run(): return kernels.hdr_merge_update(kernels.HdrMergeState(), val, std, t, theta, True, True, radiance_dtype=torch.float32)
for _ in range(2): run()
torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
bytes_ = n * 3 * 24e6 * 8 + 3 * 24e6 * 8
print(os.environ.get("CLAIR_TUNE", ""), "N", n, ms, "ms", bytes_ / ms / 1e6 / 6551.7, "of roofline")
