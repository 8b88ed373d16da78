"""Per-kernel timing of the c5 data-parallel step on this rank's band (torchrun, N ranks): means pass, gradient pass, whole eager
step, graph replay."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clair_torch_b200 as ct
from clair_torch_b200 import distributed as cd, kernels
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
if world > 1: dist.init_process_group("nccl", device_id=dev)
H, W = 8192, 12288
r0, r1 = cd.row_band(H, rank, world)
mode = sys.argv[1] if len(sys.argv) > 1 else "slice"
if mode == "slice":
    fv, fs, _ = ct.synthetic.make_stack(2, 3, H, W, bits=16, seed=5678, device=dev)
    val, std = cd.take_band(fv, r0, r1), cd.take_band(fs, r0, r1)
    del fv, fs; torch.cuda.empty_cache()
else:
    val, std, _ = ct.synthetic.make_stack(2, 3, r1 - r0, W, bits=16, seed=5678 + rank, device=dev)
exposures = torch.tensor([0.01, 0.02], dtype=torch.float64)
rb = cd.band_row_base(3, H, W, r0)
theta = torch.stack([torch.linspace(0, 1, 256) ** (2.5 + 0.15 * c) for c in range(3)]).to(dev)
i, j, r = ct.common.get_valid_exposure_pairs(exposures, 0.25)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
means = timed(lambda: kernels.pair_stats(val, std, i, j, r, theta, 1/255, 254/255, True, False, row_base=rb, means_only=True))
sums = kernels.pair_stats(val, std, i, j, r, theta, 1/255, 254/255, True, False, row_base=rb, means_only=True)
_, _, up, mg = kernels.pair_upstream(sums)
grad = timed(lambda: kernels.pair_grad(val, std, i, j, r, theta, 1/255, 254/255, True, False, up, mg, row_base=rb))
print(f"rank {rank}/{world} mode {mode} rows {r1-r0}: means {means:.3f} ms grad {grad:.3f} ms", flush=True)
if world > 1: dist.destroy_process_group()
