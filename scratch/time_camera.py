"""Kernel time of the HDR merge on device-resident integer codes, planar against the (N, H, W, 3) BGR camera layout, for the
library selected by CLAIR_B200_LIB.  argv: configs as n:h:w:bits."""
import sys

import torch

sys.path.insert(0, ".")
import clair_torch_b200 as ct  # noqa: E402
from clair_torch_b200.datasets import StdSpec  # noqa: E402

dev = torch.device("cuda:0")
lib = ct._native.load()
import os  # noqa: E402
lib.clair_set_tuning(b"hdr_prefetch", int(os.environ.get("CLAIR_PREFETCH", "0")))     # camera layout: 1 = L1, 2 = L2
theta = ct.synthetic.reference_curve(3).to(dev)
for cfg in sys.argv[1:] or ["5:1080:1920:8"]:
    n, h, w, bits = (int(a) for a in cfg.split(":"))
    val, _, t = ct.synthetic.make_stack(n, 3, h, w, bits=bits, seed=7, device=dev)
    maxval = float(2 ** bits - 1)
    codes = torch.round(val * maxval).to(torch.int32).to(torch.uint8 if bits == 8 else torch.uint16)
    del val
    camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous()
    spec = StdSpec("multiplier", 0.05)
    out, sums = {}, {}
    for name, buf, kw in (("planar", codes, {}), ("camera", camera, {"code_layout": "hwc_bgr"})):
        def once():
            return ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), buf, spec, t, theta, True, True, radiance_dtype=torch.float32, **kw)
        reps = 30 if h * w > 4e6 else 300
        # replayed as a CUDA graph so that the Python call overhead stays out of the small cases
        for _ in range(3):
            r = once()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                once()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(10):
                        once()
            torch.cuda.synchronize()
            g.replay()
            a.record()
            for _ in range(reps // 10):
                g.replay()
            b.record()
        except Exception as e:      # not capturable: plain back-to-back launches
            print("no graph:", type(e).__name__, flush=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps // 10 * 10):
                once()
            b.record()
        torch.cuda.synchronize()
        out[name] = round(a.elapsed_time(b) / (reps // 10 * 10), 4)
        sums[name] = (float(r[0].double().sum()), float(r[1].double().sum()))
    assert sums["planar"] == sums["camera"], sums
    print(cfg, out, flush=True)
