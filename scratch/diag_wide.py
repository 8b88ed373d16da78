"""Which HDR-merge kernels agree to the bit on a full-size stack of 13..16 frames: fp32 register kernel, fp32 parked kernel
(hdr_fixed_max=8), uint16 codes (planar, camera BGR).  Library selected by CLAIR_B200_LIB.  Prints mismatch counts and times."""
import sys

import torch

sys.path.insert(0, ".")
import clair_torch_b200 as ct  # noqa: E402
from clair_torch_b200.datasets import StdSpec  # noqa: E402

dev = torch.device("cuda:0")
lib = ct._native.load()
h, w = 4000, 6000
theta = ct.synthetic.reference_curve(3).to(dev)
for n in [int(a) for a in sys.argv[1:]] or [16]:
    val, _, t = ct.synthetic.make_stack(n, 3, h, w, bits=16, seed=4321, device=dev)
    codes = torch.round(val * 65535.0).to(torch.int32).to(torch.uint16)
    assert torch.equal(codes.to(torch.float32) / torch.tensor(65535.0, device=dev), val)      # IEEE quotients (a tensor divisor)
    std = val * torch.tensor(0.05, device=dev)

    def merge(v, s, **kw):
        out = ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), v, s, t, theta, True, True, radiance_dtype=torch.float32, **kw)
        return out[0].clone(), out[1].clone()

    def timed(v, s, **kw):
        for _ in range(3):
            merge(v, s, **kw)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), v, s, t, theta, True, True, radiance_dtype=torch.float32, **kw)
        b.record()
        torch.cuda.synchronize()
        return round(a.elapsed_time(b) / 10, 4)

    res = {"f32_reg": merge(val, std)}
    times = {"f32_reg": timed(val, std)}
    lib.clair_set_tuning(b"hdr_fixed_max", 8)
    res["f32_parked"] = merge(val, std)
    res["codes_parked"] = merge(codes, StdSpec("multiplier", 0.05))
    times["codes_parked"] = timed(codes, StdSpec("multiplier", 0.05))
    lib.clair_set_tuning(b"hdr_fixed_max", 0)
    res["codes"] = merge(codes, StdSpec("multiplier", 0.05))
    times["codes"] = timed(codes, StdSpec("multiplier", 0.05))
    del val, std
    camera = torch.stack([codes[:, 2], codes[:, 1], codes[:, 0]], dim=-1).contiguous()
    res["camera"] = merge(camera, StdSpec("multiplier", 0.05), code_layout="hwc_bgr")
    times["camera"] = timed(camera, StdSpec("multiplier", 0.05), code_layout="hwc_bgr")
    ref = res["f32_reg"]
    diff = {k: (int((v[0] != ref[0]).sum()), int((v[1] != ref[1]).sum())) for k, v in res.items() if k != "f32_reg"}
    worst = {k: float(((v[1] - ref[1]).abs() / ref[1].abs().clamp_min(1e-30)).max()) for k, v in res.items() if k != "f32_reg"}
    print(n, "mismatch(rad,sig) vs f32_reg", diff, "max rel sig", worst, "ms", times, flush=True)
    del res, codes, camera
    torch.cuda.empty_cache()
