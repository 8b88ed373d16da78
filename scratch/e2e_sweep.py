"""Zero-copy c1 merge (pinned host in / pinned host out) under different launch shapes: ms per stack."""
import ctypes, itertools, sys
import numpy as np, torch
import clair_torch_b200 as ct

dev = torch.device("cuda", 0)
lib = ct._native.load()
N, C, H, W, L = 5, 3, 1080, 1920, 256
theta = ct.synthetic.reference_curve(C, L).to(dev)
val, std, t = ct.synthetic.make_stack(N, C, H, W, bits=8, seed=1, device=dev)
val_h, std_h = val.cpu().pin_memory(), std.cpu().pin_memory()
rad_h = torch.empty((C, H, W), dtype=torch.float32).pin_memory()
sig_h = torch.empty_like(rad_h).pin_memory()
rad_d = torch.empty((C, H, W), dtype=torch.float32, device=dev)
sig_d = torch.empty_like(rad_d)
t_host = np.ascontiguousarray(t)
stream = torch.cuda.current_stream(dev)


def run(vp, sp, rp, gp, reps=8):
    def once():
        ct._native.check(lib.clair_hdr_merge_update(vp, sp, t_host.ctypes.data_as(ctypes.c_void_p), N, theta.data_ptr(), C, L, H * W,
                                                    None, 1, None, None, None, 1, 1, rp, 0, gp, stream.cuda_stream), "merge")
    once(); once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        once()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for vec, waves in itertools.product((1, 2, 4), (1, 2, 4, 8)):
    lib.clair_set_tuning(b"hdr_vec", vec); lib.clair_set_tuning(b"hdr_waves", waves)
    both = run(val_h.data_ptr(), std_h.data_ptr(), rad_h.data_ptr(), sig_h.data_ptr())
    inp = run(val_h.data_ptr(), std_h.data_ptr(), rad_d.data_ptr(), sig_d.data_ptr())
    print(f"vec {vec} waves {waves}: host->host {both:.3f} ms   host->device {inp:.3f} ms  ({(val_h.numel()*8)/inp/1e6:.1f} GB/s in)", flush=True)
# plain copies for scale
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(stream)
for _ in range(8):
    val.copy_(val_h, non_blocking=True); std.copy_(std_h, non_blocking=True)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 8
print(f"copy-engine H2D of the stack: {ms:.3f} ms ({val_h.numel()*8/ms/1e6:.1f} GB/s)")

# staged (copy engine + band kernels) through the Python wrapper, host -> host
lib.clair_set_tuning(b"hdr_vec", 0); lib.clair_set_tuning(b"hdr_waves", 0)
for bands in (1, 2, 4, 8, 16, 32, 64):
    def once():
        ct.kernels.hdr_merge_update(ct.kernels.HdrMergeState(), val_h, std_h, t_host, theta, True, True, radiance_dtype=torch.float32,
                                    device=dev, host_out=(rad_h, sig_h), staged=True, bands=bands)
    once(); once(); torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(8):
        once()
    e1.record(stream); torch.cuda.synchronize()
    print(f"staged bands {bands}: host->host {e0.elapsed_time(e1)/8:.3f} ms")
