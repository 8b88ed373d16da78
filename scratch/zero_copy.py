"""Experiment: HDR merge kernel reading the stack straight from pinned host memory and writing results straight back."""
import ctypes, os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
lib = ct._native.load()
for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):
    k, _, v = item.partition("="); lib.clair_set_tuning(k.encode(), int(v))
dev = torch.device("cuda", 0)
N, C, H, W = 5, 3, 1080, 1920
val, std, t = ct.synthetic.make_stack(N, C, H, W, device=dev)
theta = ct.synthetic.reference_curve(3).to(dev)
val_h, std_h = val.cpu().pin_memory(), std.cpu().pin_memory()
rad_h = torch.empty((C, H, W), dtype=torch.float32).pin_memory(); sig_h = torch.empty_like(rad_h).pin_memory()
rad_d = torch.empty((C, H, W), dtype=torch.float32, device=dev); sig_d = torch.empty_like(rad_d)
tt = np.ascontiguousarray(t)
st = torch.cuda.current_stream(dev)
def run(v, s, r, g):
    rc = lib.clair_hdr_merge_update(v.data_ptr(), s.data_ptr(), tt.ctypes.data_as(ctypes.c_void_p), N, theta.data_ptr(), C, 256, H * W,
                                    None, 1, None, None, None, 1, 1, r.data_ptr(), 0, g.data_ptr(), st.cuda_stream)
    ct._native.check(rc, "merge")
def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("device->device      ", timeit(lambda: run(val, std, rad_d, sig_d), 50), "ms")
print("host in, device out ", timeit(lambda: run(val_h, std_h, rad_d, sig_d)), "ms")
print("host in, host out   ", timeit(lambda: run(val_h, std_h, rad_h, sig_h)), "ms")
def staged():
    v = val_h.to(dev, non_blocking=True); s = std_h.to(dev, non_blocking=True)
    run(v, s, rad_d, sig_d); rad_h.copy_(rad_d, non_blocking=True); sig_h.copy_(sig_d, non_blocking=True)
print("staged copies       ", timeit(staged), "ms")
run(val, std, rad_d, sig_d); run(val_h, std_h, rad_h, sig_h); torch.cuda.synchronize()
print("equal:", torch.equal(rad_h, rad_d.cpu()), torch.equal(sig_h, sig_d.cpu()))
