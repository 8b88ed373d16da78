"""End-to-end time of the HDR merge from pinned host integer codes (zero-copy: the kernel reads host memory over PCIe) to pinned
host radiance + sigma, planar against camera layout, with the next-trip prefetch of the camera kernels off (-1) and at its default."""
import sys
import time

import torch

sys.path.insert(0, ".")
import clair_torch_b200 as ct  # noqa: E402
from clair_torch_b200.datasets import StdSpec  # noqa: E402

dev = torch.device("cuda:0")
lib = ct._native.load()
theta = ct.synthetic.reference_curve(3).to(dev)
for cfg in sys.argv[1:] or ["9:4000:6000:16"]:
    n, h, w, bits = (int(a) for a in cfg.split(":"))
    val, _, t = ct.synthetic.make_stack(n, 3, h, w, bits=bits, seed=7, device=dev)
    codes = torch.round(val * float(2 ** bits - 1)).to(torch.int32).to(torch.uint8 if bits == 8 else torch.uint16)
    del val
    camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous().cpu().pin_memory()
    planar = codes.cpu().pin_memory()
    del codes
    rad_h = torch.empty((3, h, w), dtype=torch.float32).pin_memory()
    sig_h = torch.empty_like(rad_h).pin_memory()
    spec = StdSpec("multiplier", 0.05)
    out = {}
    for name, buf, kw, knob in (("planar", planar, {}, 0), ("camera_noprefetch", camera, {"code_layout": "hwc_bgr"}, -1),
                                ("camera_default", camera, {"code_layout": "hwc_bgr"}, 0)):
        lib.clair_set_tuning(b"hdr_prefetch", knob)

        def once():
            ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), buf, spec, t, theta, True, True, radiance_dtype=torch.float32,
                                        device=dev, host_out=(rad_h, sig_h), **kw)
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        reps = 5 if h * w > 4e6 else 50
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        out[name] = round((time.perf_counter() - t0) / reps * 1e3, 3)
        out[name + "_sum"] = float(rad_h.double().sum())
    print(cfg, out, flush=True)
