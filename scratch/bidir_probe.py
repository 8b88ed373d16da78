"""How fast can a 24 MP frame go host -> device and device -> host at the same time (two copy engines)?"""
import torch
dev = torch.device("cuda", 0)
n = 3 * 4000 * 6000
hin = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
hout = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
din = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
dout = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2)]
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(both, reps=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        with torch.cuda.stream(s1):
            for k in range(2): din[k].copy_(hin[k], non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                for k in range(2): hout[k].copy_(dout[k], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for both in (False, True):
    run(both, 2)
    ms = run(both)
    print(f"{'H2D + D2H' if both else 'H2D only'} of 576 MB each: {ms:.2f} ms  ({2*n*4/ms/1e6:.1f} GB/s per direction)")
