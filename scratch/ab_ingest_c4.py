"""Kernel time of clair_hdr_merge_codes on a c4 stack handed over as uint16 codes (planar and BGR camera layout), for the
library selected by CLAIR_B200_LIB.  Used to A/B the 2-code register kernel (9..16 frames) against the parked kernel."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import clair_torch_b200 as ct  # noqa: E402

dev = torch.device("cuda:0")
lib = ct._native.load()
n, c, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 9, 3, 4000, 6000
val, _, t = ct.synthetic.make_stack(n, c, h, w, bits=16, seed=4321, device=dev)
codes = torch.round(val * 65535.0).to(torch.int32).to(torch.uint16)
del val
camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous()
theta = ct.synthetic.reference_curve(3).to(dev)
t_host = np.asarray(t, dtype=np.float64)
rad = torch.empty((c, h, w), dtype=torch.float32, device=dev)
sig = torch.empty_like(rad)
stream = torch.cuda.current_stream(dev)
out = {}
for name, buf, fn in (("planar", codes, lib.clair_hdr_merge_codes), ("camera", camera, None)):
    def launch():
        if fn is not None:
            rc = fn(buf.data_ptr(), 2, 65535.0, None, 2, 0.05, t_host.ctypes.data_as(ctypes.c_void_p), n, theta.data_ptr(), c, 256,
                    h * w, None, 1, None, None, None, 1, 1, rad.data_ptr(), 0, sig.data_ptr(), stream.cuda_stream)
            ct._native.check(rc, "clair_hdr_merge_codes")
        else:
            from clair_torch_b200.datasets import StdSpec
            ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), buf, StdSpec("multiplier", 0.05), t, theta, True, True,
                                        code_layout="hwc_bgr", radiance_dtype=torch.float32)
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(30):
        launch()
    b.record(stream)
    torch.cuda.synchronize()
    out[name] = round(a.elapsed_time(b) / 30, 4)
    out[name + "_sum"] = float(rad.double().sum().item()), float(sig.double().sum().item())
print(n, out)
