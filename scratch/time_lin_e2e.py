import time, numpy as np, torch
import clair_torch_b200 as ct
from clair_torch_b200 import kernels
dev = torch.device("cuda", 0)
val, std, t = ct.synthetic.make_stack(1, 3, 4000, 6000, bits=16, seed=4567, device=dev)
theta = ct.synthetic.reference_curve(3).to(dev)
hv, hs = val.cpu().pin_memory(), std.cpu().pin_memory()
for staged, bands in ((False, 1), (True, 4), (True, 8), (True, 16), (True, 32), (True, 64)):
    for _ in range(2):
        kernels.linearize(hv, hs, theta, device=dev, pinned_out=True, staged=staged, bands=bands)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        out = kernels.linearize(hv, hs, theta, device=dev, pinned_out=True, staged=staged, bands=bands)
    torch.cuda.synchronize()
    print(f"staged={staged} bands={bands}: {(time.perf_counter()-t0)/5*1e3:.2f} ms per 24 MP frame", flush=True)
