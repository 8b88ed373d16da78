#!/usr/bin/env python
"""Benchmark of the headline metric: HDR-merge throughput in Mpixel*frames/s (BASELINE.json).

Workload (SURVEY.md §8(d), config c4 = BASELINE.json configs[3], the configuration the "at 1/2/4/8 B200" metric is quoted
on): every rank owns its share of the 64 stacks of 9 frames x 24 MP (4000x6000) 16-bit RGB, handed over as fp32 value +
fp32 std images; fixed 256-entry ICRF with distinct rows (LINEAR), Gaussian weights, whole stack in one batch, radiance +
uncertainty written as fp32.  One "step" = one merge of one stack (216 Mpixel*frames, 5.76 GB of algorithmic traffic).
`--workload c1` selects the reference's own CPU-sized case instead (configs[0], 5 x 1080p 8-bit); the default run
reports it under `extra`.

  value     kernel-only: stacks resident in HBM, K back-to-back launches through the C ABI, CUDA events.
            Distinct stacks are rotated (each is 5.2 GB, far beyond the 126 MB L2).
  e2e       the same merge through the public API (compute_hdr_image) from PINNED HOST buffers, with the
            host->device copy of the stack and the device->host read of radiance + uncertainty inside the
            timed region.
  roofline  algorithmic bytes (28.8 B per pixel*frame) / measured kernel time vs the measured HBM copy peak.
  cpu_baseline  the CPU oracle port of the reference algorithm timed on this box's host cores on a row-cropped
            sample of the same stack (rank 0, N=1 only).

`--impl reference` times the CPU implementation of the path (the oracle port: the reference is pure Python/torch
and cannot travel to the GPU box) with all host threads on the same config and prints the same JSON line.
Multi-GPU (torchrun): stacks shard by stack across ranks (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_FRAMES, CHANNELS, HEIGHT, WIDTH, BITS, LUT = 5, 3, 1080, 1920, 8, 256       # c1 geometry (secondary metrics)
METRIC = "hdr_merge_mpixel_frames_per_s"
WORKLOADS = {
    "c4": {"n": 9, "c": 3, "h": 4000, "w": 6000, "bits": 16, "seed": 4567, "n_sets": 3,
           "kernel": "clair::hdr_merge_fixed_kernel<2,9,1,true,0> (2 px/thread, N=9 in registers, single batch)",
           "label": ("c4: HDR merge of 9-frame 24 MP (4000x6000) 16-bit RGB synthetic exposure stacks as fp32 val+std, sharded "
                     "by stack (each rank merges its share of the 64 stacks), 256-entry LINEAR ICRF (distinct rows), gaussian "
                     "weights, first-order uncertainty, fp32 radiance+sigma")},
    "c1": {"n": 5, "c": 3, "h": 1080, "w": 1920, "bits": 8, "seed": 1234, "n_sets": 4,
           "kernel": "clair::hdr_merge_fixed_kernel<2,5,1,true,0> (2 px/thread, N=5 in registers, single batch)",
           "label": ("c1: HDR merge, 5x 8-bit RGB 1920x1080 synthetic exposure stack as fp32 val+std, 256-entry LINEAR ICRF "
                     "(distinct rows), gaussian weights, first-order uncertainty, fp32 radiance+sigma")},
}


def algo_bytes_per_pixel_frame(cfg):
    """C*(b_in(1+sigma) + 8/N) with fp32 val + std in and fp32 radiance + sigma out."""
    return cfg["c"] * (4 * 2 + 8 / cfg["n"])


def ncu_traffic(key):
    """DRAM bytes per launch of the headline kernel from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r1_hdr_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        d = json.load(fh).get(key)
    return None if d is None else d["dram_bytes_read_per_launch"] + d["dram_bytes_write_per_launch"]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)            # nvidia-smi holds driver locks while it runs: be sure it is gone
        except Exception:                        # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(cfg, rows, repeats=1):
    """Mpixel*frames/s of the CPU oracle (C + OpenMP port of the reference algorithm, all host threads) on the
    first `rows` rows of one stack of the workload.  Returns (rate, seconds per merge, threads)."""
    import clair_torch_b200.synthetic as syn
    from oracle import c_oracle as corc
    threads = corc.use_all_host_threads()                             # torchrun exports OMP_NUM_THREADS=1
    val, std, t = syn.make_stack(cfg["n"], cfg["c"], rows, cfg["w"], bits=cfg["bits"], seed=cfg["seed"])
    theta = syn.reference_curve(cfg["c"], LUT).numpy()
    v, s = val.numpy(), std.numpy()
    corc.hdr_merge(v[:, :, :8], s[:, :, :8], t, theta, True)          # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    for _ in range(repeats):
        corc.hdr_merge(v, s, t, theta, True)
    dt = (time.perf_counter() - t0) / repeats
    return cfg["n"] * rows * cfg["w"] / dt / 1e6, dt, threads


def shared_config(cfg):
    """The `config` object both arms print (identical, so the driver can pair the lines)."""
    return {"workload": cfg["label"],
            "stacks_per_rank_rotated": cfg["n_sets"],
            "l2": f"{cfg['n_sets']} distinct stacks of {2 * cfg['n'] * cfg['c'] * cfg['h'] * cfg['w'] * 4 / 1e6:.0f} MB rotated "
                  "per step, each far larger than the 126 MB L2",
            "sharding": "by stack, one rank per GPU, no data-path collective",
            "cpu_arm": "the reference arm / cpu_baseline run the reference's own compute_hdr_image (torch CPU, all host threads) "
                       "on a band of rows of one stack per step; the metric is a throughput, so the band size cancels"}


class ReferenceHdr:
    """The reference's own compute_hdr_image (oracle/_ref, torch CPU, every host thread) on rows 0..rows of one stack of
    the workload; falls back to the C/OpenMP port only when oracle/_ref did not travel."""

    def __init__(self, cfg, rows):
        import clair_torch_b200.synthetic as syn
        from oracle import reference_runner as rr
        self.cfg, self.rows = cfg, rows
        val, std, self.t = syn.make_stack(cfg["n"], cfg["c"], rows, cfg["w"], bits=cfg["bits"], seed=cfg["seed"])
        self.v, self.s = val.numpy(), std.numpy()
        self.theta = syn.reference_curve(cfg["c"], LUT).numpy()
        self.kind = "reference" if rr.available() else "port"
        if self.kind == "reference":
            self.threads = rr.use_all_host_threads()
            self.rr = rr
        else:
            from oracle import c_oracle as corc
            self.threads = corc.use_all_host_threads()
            self.corc = corc
        self.units = cfg["n"] * rows * cfg["w"] / 1e6

    def step(self):
        """One merge of the band; returns seconds."""
        if self.kind == "reference":
            return self.rr.hdr_merge(self.v, self.s, self.t, self.theta, True)[2]
        t0 = time.perf_counter()
        self.corc.hdr_merge(self.v, self.s, self.t, self.theta, True)
        return time.perf_counter() - t0

    def describe(self, steps):
        what = ("clair_torch.inference.hdr_merge.compute_hdr_image of the unmodified reference (oracle/_ref), torch CPU"
                if self.kind == "reference" else "C/OpenMP oracle port of the reference algorithm (oracle/_ref absent)")
        return (f"{steps} merges of rows 0..{self.rows} of {self.cfg['h']} of one stack ({self.units:.2f} Mpixel*frames each, "
                f"{self.rows / self.cfg['h']:.4f} of a stack); {what}, {self.threads} threads")


def reference_rows(cfg, budget_s, n_steps):
    """Rows of one stack per step so that n_steps steps of the reference take about budget_s (calibrated on a thin band)."""
    probe_rows = 16
    probe = ReferenceHdr(cfg, probe_rows)
    probe.step()
    dt = min(probe.step(), probe.step())
    rows = int(probe_rows * budget_s / max(dt * max(n_steps, 1), 1e-9))
    # >= 32 rows keeps torch's per-op overhead small; <= 400 rows of a 24 MP frame keeps the reference's temporaries < 30 GB
    cap = 400 if cfg["w"] > 2000 else cfg["h"]
    return max(32, min(cfg["h"], cap, rows))


def cpu_baseline_hdr(cfg, budget_s=12.0, steps=3):
    """cpu_baseline object for the HDR merge: the reference itself, with the C/OpenMP port as a second stated number."""
    rows = reference_rows(cfg, budget_s, steps + 1)
    ref = ReferenceHdr(cfg, rows)
    ref.step()
    dts = sorted(ref.step() for _ in range(steps))
    dt = dts[len(dts) // 2]
    port_rate, port_dt, port_threads = cpu_oracle_rate(cfg, min(cfg["h"], 400), repeats=3)
    return {"value": ref.units / dt, "unit": "Mpixel*frames/s", "cores": ref.threads, "kind": ref.kind,
            "sample": ref.describe(steps) + f", median {dt:.3f} s per merge",
            "port": {"value": port_rate, "unit": "Mpixel*frames/s", "cores": port_threads,
                     "what": "oracle/clair_oracle.c (C + OpenMP restatement, closed-form variance), the checker the parity "
                             "tests use at full size; reported for scale only"}}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path, K timed steps on rank 0."""
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    rows = reference_rows(cfg, 75.0, args.steps + args.warmup)
    ref = ReferenceHdr(cfg, rows)
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = ref.units / dt
    sample = ref.describe(args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel*frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(cfg),
            "cpu_baseline": {"value": value, "unit": "Mpixel*frames/s", "cores": ref.threads, "kind": ref.kind, "sample": sample},
            "e2e": {"value": value, "unit": "Mpixel*frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline_train_c2(budget_rows=120):
    """The reference's train-step body (icrf_training.py:105-156, oracle/reference_runner.TrainStep) on a band of rows of
    the c2 stack; a step's cost is linear in the pixel count, so the full-frame figure is the stated extrapolation."""
    import clair_torch_b200.synthetic as syn
    from oracle import reference_runner as rr
    if not rr.available():
        return None
    threads = rr.use_all_host_threads()
    val, std, t = syn.make_stack(10, CHANNELS, budget_rows, WIDTH, bits=8, seed=2345)
    step = rr.TrainStep(val.numpy(), std.numpy(), t, channels=CHANNELS, relative=True, unc_weighting=False, alpha=10.0, beta=1.0,
                        gamma=1.0, delta=1.0, threshold=0.25)
    step()                                   # step 0 only connects the table to the parameters
    dts = sorted(step()[3] for _ in range(2))
    dt = dts[0]
    full = dt * HEIGHT / budget_rows
    return {"value": 1.0 / full, "unit": "steps/s", "cores": threads, "kind": "reference",
            "measured_s_per_step_on_sample": dt, "extrapolated_s_per_full_step": full,
            "sample": f"rows 0..{budget_rows} of {HEIGHT} of the c2 stack (10x3x{budget_rows}x{WIDTH}, P=17), step body of "
                      "clair_torch/training/icrf_training.py:105-156 from the reference's own functions, torch CPU; the full "
                      "frame needs ~29 GB of (P,C,H,W) temporaries, so the per-step time is scaled by H/rows (labelled "
                      "extrapolation, BASELINE.md section 3)"}


def cpu_baseline_linearity_c3(rows=96):
    """The reference's measure_linearity on a band of rows of the c3 stack (the full 4K stack needs ~250 GB of temporaries)."""
    import clair_torch_b200.synthetic as syn
    from oracle import reference_runner as rr
    if not rr.available():
        return None
    threads = rr.use_all_host_threads()
    val, std, t = syn.make_stack(16, CHANNELS, rows, 3840, bits=16, seed=3456)
    theta = syn.reference_curve(CHANNELS, LUT).numpy()
    v, s = val.numpy(), std.numpy()
    rr.measure_linearity(v[:, :, :8], s[:, :, :8], t, theta, True, True)
    dt = rr.measure_linearity(v, s, t, theta, True, True)[4]
    full = dt * 2160 / rows
    return {"value": full * 1e3, "unit": "ms", "cores": threads, "kind": "reference", "measured_s_on_sample": dt,
            "extrapolated_ms_full": full * 1e3,
            "sample": f"rows 0..{rows} of 2160 of the c3 stack (16x3x{rows}x3840, P=29), clair_torch.inference.measure_linearity "
                      "of the unmodified reference, torch CPU; scaled by H/rows (labelled extrapolation: the full stack would "
                      "need ~250 GB)"}


def native_ingest_metrics(dev, lib, theta, t_host, cfg=None, reps=400, e2e_reps=20, camera=True):
    """The workload again (c1 by default), but handing over what the camera produced: uint8 / uint16 codes +
    MissingStdMode.MULTIPLIER(0.05) evaluated in the kernel (reported separately from the
    fp32-boundary headline)."""
    import ctypes
    import torch
    import clair_torch_b200 as ct
    from clair_torch_b200.datasets import StdSpec
    cfg = cfg or WORKLOADS["c1"]
    N_FRAMES, CHANNELS, HEIGHT, WIDTH, BITS = cfg["n"], cfg["c"], cfg["h"], cfg["w"], cfg["bits"]     # shadow the c1 globals
    maxval = float(2 ** BITS - 1)
    code_dtype, code_kind, code_bytes = (torch.uint8, 1, 1) if BITS == 8 else (torch.uint16, 2, 2)
    n_sets = 4 if BITS == 8 else 2
    stream = torch.cuda.current_stream(dev)
    sets = []
    for k in range(n_sets):
        val, _, _ = ct.synthetic.make_stack(N_FRAMES, CHANNELS, HEIGHT, WIDTH, bits=BITS, seed=4321 + k, device=dev)
        sets.append(torch.round(val * maxval).to(torch.int32).to(code_dtype))
        del val
    radiance = torch.empty((CHANNELS, HEIGHT, WIDTH), dtype=torch.float32, device=dev)
    sigma = torch.empty_like(radiance)

    def launch(k):
        rc = lib.clair_hdr_merge_codes(sets[k % n_sets].data_ptr(), code_kind, maxval, None, 2, 0.05, t_host.ctypes.data_as(ctypes.c_void_p),
                                       N_FRAMES, theta.data_ptr(), CHANNELS, LUT, HEIGHT * WIDTH, None, 1, None, None, None, 1, 1,
                                       radiance.data_ptr(), 0, sigma.data_ptr(), stream.cuda_stream)
        ct._native.check(rc, "clair_hdr_merge_codes")

    for k in range(10):
        launch(k)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(reps):
        launch(k)
    b.record(stream)
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / reps
    units = N_FRAMES * HEIGHT * WIDTH / 1e6
    # end to end: pinned uint8 stack in, pinned fp32 radiance + sigma out, both over PCIe from inside the kernel (zero-copy)
    codes_h = sets[0].cpu().pin_memory()
    rad_h = torch.empty((CHANNELS, HEIGHT, WIDTH), dtype=torch.float32).pin_memory()
    sig_h = torch.empty_like(rad_h).pin_memory()
    batch = (torch.arange(N_FRAMES), codes_h, StdSpec("multiplier", 0.05), {"exposure_time": torch.from_numpy(t_host)})

    class OneBatch(torch.utils.data.Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return batch

    loader = torch.utils.data.DataLoader(OneBatch(), batch_size=None, shuffle=False)
    model = ct.ICRFModelDirect(icrf=theta.clone()).to(dev)

    def e2e():
        ct.compute_hdr_image(loader, dev, model, max, radiance_dtype=torch.float32, host_out=(rad_h, sig_h))

    for _ in range(3):
        e2e()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        e2e()
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) / e2e_reps * 1e3
    algo = N_FRAMES * CHANNELS * HEIGHT * WIDTH * code_bytes + CHANNELS * HEIGHT * WIDTH * 8
    out = {"config": f"{cfg['label'].split(':')[0]} as uint{8 * code_bytes} codes, CastTo+Normalize({int(maxval)}) and std = 0.05*value fused "
                     "into the load",
           "kernel_ms": ms, "mpixel_frames_per_s": units / (ms * 1e-3), "dram_bytes_per_pixel_frame": algo / (N_FRAMES * HEIGHT * WIDTH),
           "hbm_frac": algo / (ms * 1e-3) / 1e9 / peaks()[0],
           "e2e_ms": e2e_ms, "e2e_mpixel_frames_per_s": units / (e2e_ms * 1e-3), "e2e_h2d_bytes": codes_h.numel() * code_bytes,
           "e2e_d2h_bytes": 2 * rad_h.numel() * 4}
    if not camera:
        return out
    # the same stack as the camera delivers it: (N, H, W, 3) BGR (cv2.imread layout), CvToTorch fused into the load as well
    from clair_torch_b200 import kernels
    cam = [torch.stack([c8[:, 2], c8[:, 1], c8[:, 0]], dim=-1).contiguous() for c8 in sets[:2]]
    spec = StdSpec("multiplier", 0.05)

    def launch_cam(k):
        kernels.hdr_merge_update(kernels.HdrMergeState(), cam[k % 2], spec, t_host, theta, True, True, radiance_dtype=torch.float32,
                                 code_layout="hwc_bgr")

    for k in range(5):
        launch_cam(k)
    torch.cuda.synchronize(dev)
    a.record(stream)
    cam_reps = min(100, reps)
    for k in range(cam_reps):
        launch_cam(k)
    b.record(stream)
    torch.cuda.synchronize(dev)
    ms_cam = a.elapsed_time(b) / cam_reps
    cam_h = cam[0].cpu().pin_memory()
    batch_cam = (torch.arange(N_FRAMES), cam_h, spec, {"exposure_time": torch.from_numpy(t_host)})

    class OneCamBatch(torch.utils.data.Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return batch_cam

    cam_loader = torch.utils.data.DataLoader(OneCamBatch(), batch_size=None, shuffle=False)

    def e2e_cam():
        ct.compute_hdr_image(cam_loader, dev, model, max, radiance_dtype=torch.float32, host_out=(rad_h, sig_h), code_layout="hwc_bgr")

    for _ in range(3):
        e2e_cam()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        e2e_cam()
    torch.cuda.synchronize(dev)
    e2e_cam_ms = (time.perf_counter() - t0) / e2e_reps * 1e3
    out["camera_layout"] = {"config": "same codes as (N, H, W, 3) BGR camera buffers, CvToTorch fused into the load too",
                            "kernel_ms": ms_cam, "e2e_ms": e2e_cam_ms, "e2e_mpixel_frames_per_s": units / (e2e_cam_ms * 1e-3)}
    return out


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and hence its first-touch pinned buffers) to the CPUs NVML reports as local to
    the GPU: the end-to-end path streams the stack over PCIe straight from host memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def dp_training_metrics(dev, rank, world):
    """c5: data-parallel ICRF training on a 100 MP 16-bit exposure pair, row bands across the ranks; one pass over the band
    (clair_pair_fused: statistics + un-normalised table gradient) and ONE NCCL sum all-reduce of its 18.6 KB float64 buffer
    per step (strong scaling: the image is fixed).

    Every rank synthesises the SAME full image (same device generator seed) and keeps its own band, so N = 1 and N = 8
    train on the same pixels.  Before the timed steps the banded, all-reduced pass is checked against rank 0's pass over
    the whole image at the same table: `parity` = relative error of the loss and error of the gradient over its maximum."""
    import torch
    import torch.distributed as dist
    import clair_torch_b200 as ct
    from clair_torch_b200 import distributed as cd
    from clair_torch_b200.training import linearity_loss_and_table_grad
    height, width, n_frames = 8192, 12288, 2
    r0, r1 = cd.row_band(height, rank, world)
    full_val, full_std, _ = ct.synthetic.make_stack(n_frames, CHANNELS, height, width, bits=16, seed=5678, device=dev)
    val, std = cd.take_band(full_val, r0, r1), cd.take_band(full_std, r0, r1)
    exposures = torch.tensor([0.01, 0.02], dtype=torch.float64)
    rb = cd.band_row_base(CHANNELS, height, width, r0)
    # parity of the sharded pass (all ranks) against the whole image (rank 0) at a table with distinct rows
    table = torch.stack([torch.linspace(0, 1, LUT) ** (2.5 + 0.15 * c) for c in range(CHANNELS)]).to(dev)
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(exposures, 0.25)
    lin_b, _, g_b = linearity_loss_and_table_grad(val, std, i_idx, j_idx, ratio, table, 1 / 255, 254 / 255, True, False, row_base=rb,
                                                  reduce_fn=lambda t: cd.all_reduce_sum_(t))
    parity = None
    if rank == 0:
        lin_w, _, g_w = linearity_loss_and_table_grad(full_val, full_std, i_idx, j_idx, ratio, table, 1 / 255, 254 / 255, True, False)
        parity = {"loss_rel": float(((lin_b - lin_w).abs() / lin_w.abs()).max()),
                  "grad_rel_of_max": float((g_b - g_w).abs().max() / g_w.abs().max()),
                  "what": "banded + all-reduced pass on all ranks vs rank 0's pass over the whole 100.7 MP image, same table"}
    if world > 1:
        del full_val, full_std
        torch.cuda.empty_cache()
    model = ct.ICRFModelDirect(256, CHANNELS, ct.InterpMode.LINEAR, 2.5).to(dev)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3, capturable=True, fused=True) for c in range(CHANNELS)]
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=False, alpha=10.0, beta=1.0, gamma=1.0, delta=1.0,
              exposure_ratio_threshold=0.25)
    step = lambda: cd.train_icrf_step_data_parallel(model, opts, val, std, exposures, rb, **kw)
    for _ in range(3):
        step()
    mode, graphed = "eager", None
    try:      # kernels + both all-reduces + optimisers as one replayed CUDA graph (capture executes nothing)
        graphed = cd.graphed_train_step_data_parallel(model, opts, val, std, exposures, rb, **kw)
    except Exception as exc:      # noqa: BLE001 - a capture the collective library refuses falls back to the eager step
        mode = f"eager (graph capture failed: {type(exc).__name__})"
    flag = torch.tensor([0.0 if graphed is None else 1.0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # every rank must take the same path
    if flag.item() == 1.0:
        step, mode = graphed, "cuda graph (kernels, NCCL all-reduces and optimisers captured)"
    for _ in range(10):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    # (no nvidia-smi sampling here: the step is ~35 short launches per replay, and a polling nvidia-smi on a many-GPU box
    # holds driver locks long enough to show in it)
    a.record()
    for _ in range(reps):
        step()
    b.record()
    torch.cuda.synchronize(dev)
    own_ms = a.elapsed_time(b) / reps
    ms = torch.tensor([own_ms], dtype=torch.float64, device=dev)
    ms_min = ms.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_min, op=dist.ReduceOp.MIN)
    ms, ms_min = float(ms.item()), float(ms_min.item())
    # this rank's kernels alone (no collective, no optimiser): what the band costs on the device
    def local_pass():
        linearity_loss_and_table_grad(val, std, i_idx, j_idx, ratio, table, 1 / 255, 254 / 255, True, False, row_base=rb)
    for _ in range(3):
        local_pass()
    torch.cuda.synchronize(dev)
    a.record()
    for _ in range(10):
        local_pass()
    b.record()
    torch.cuda.synchronize(dev)
    local_ms = torch.tensor([a.elapsed_time(b) / 10], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(local_ms, op=dist.ReduceOp.MAX)
    local_ms = float(local_ms.item())
    # what does not shrink with N: the all-reduce alone (eager launches back to back, outside the graph)
    collective_us = None
    fused = torch.zeros(CHANNELS * 5 + CHANNELS * CHANNELS * LUT, dtype=torch.float64, device=dev)
    if world > 1:
        for _ in range(5):
            dist.all_reduce(fused)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(50):
            dist.all_reduce(fused)
        b.record()
        torch.cuda.synchronize(dev)
        collective_us = a.elapsed_time(b) / 50 * 1e3
    # ... and everything else of a step that does not shrink either: the step replayed on an 8-row band (launch overheads of
    # the ~25 graph nodes, finalize / combine / penalties, the autograd edge, three fused Adam steps, update_icrf)
    fixed_us = None
    try:
        tiny_val, tiny_std = val[:, :, :8].contiguous(), std[:, :, :8].contiguous()
        tiny_model = ct.ICRFModelDirect(256, CHANNELS, ct.InterpMode.LINEAR, 2.5).to(dev)
        tiny_opts = [torch.optim.Adam(tiny_model.channel_params(c), lr=1e-3, capturable=True, fused=True) for c in range(CHANNELS)]
        for _ in range(3):
            ct.train_icrf_step(tiny_model, tiny_opts, tiny_val, tiny_std, exposures, row_base=rb, **kw)
        tiny = ct.GraphedTrainStep(tiny_model, tiny_opts, tiny_val, tiny_std, exposures, row_base=rb, **kw)
        for _ in range(5):
            tiny()
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(50):
            tiny()
        b.record()
        torch.cuda.synchronize(dev)
        fixed_us = a.elapsed_time(b) / 50 * 1e3
    except Exception:      # noqa: BLE001 - a diagnostic, never a reason to lose the line
        fixed_us = None
    return {"config": "100.7 MP (8192x12288) 16-bit RGB exposure pair (one image, same seed at every N), row bands, one pass "
                      "over the band and one NCCL all-reduce (sums + un-normalised gradient tables, 18.6 KB) per step",
            "ms_per_step": ms, "steps_per_s": 1e3 / ms, "scaling": "strong", "n_gpus": world, "step": mode, "parity": parity,
            "ms_per_step_fastest_rank": ms_min, "local_kernels_ms": local_ms,
            "collective_us": collective_us, "fixed_us": fixed_us,
            "residual": "collective_us = the all-reduce alone (eager, back to back); fixed_us = the captured step on an 8-row band "
                        "without the collective: what a step costs however small the band"}


def secondary_metrics(dev):
    """BASELINE.json's other single-GPU configs, device-resident, CUDA-event timed: c2 ICRF train steps/s and
    c3 linearity measurement.  Reported under "extra"; the headline metric stays the HDR merge."""
    import torch
    import clair_torch_b200 as ct
    from clair_torch_b200 import kernels
    from clair_torch_b200.inference.measure_linearity import spatial_statistics
    out = {}
    stream = torch.cuda.current_stream(dev)

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    # c2: ICRF training, 10 exposures of 1080p 8-bit, 256-bin curve, script settings
    val, std, t = ct.synthetic.make_stack(10, CHANNELS, HEIGHT, WIDTH, bits=8, seed=2345, device=dev)
    exposures = torch.from_numpy(t)
    model = ct.ICRFModelDirect(256, CHANNELS, ct.InterpMode.LINEAR, 2.5).to(dev)
    # capturable, fused Adam = train_icrf's default optimisers: the step is replayed as one CUDA graph (GraphedTrainStep)
    opts = [torch.optim.Adam(model.channel_params(c), lr=1e-3, capturable=True, fused=True) for c in range(CHANNELS)]
    kw = dict(use_relative_linearity_loss=True, use_uncertainty_weighting=False, alpha=10.0, beta=1.0, gamma=1.0,
              delta=1.0, exposure_ratio_threshold=0.25)
    ms_eager = timed(lambda: ct.train_icrf_step(model, opts, val, std, exposures, **kw), 3, 20)
    ms = timed(ct.GraphedTrainStep(model, opts, val, std, exposures, **kw), 3, 50)
    kw2 = dict(kw, use_uncertainty_weighting=True, exposure_ratio_threshold=0.1)
    ms2_eager = timed(lambda: ct.train_icrf_step(model, opts, val, std, exposures, **kw2), 2, 10)
    ms2 = timed(ct.GraphedTrainStep(model, opts, val, std, exposures, **kw2), 3, 20)
    elems = 10 * CHANNELS * HEIGHT * WIDTH
    out["icrf_train_c2"] = {"steps_per_s": 1e3 / ms, "ms_per_step": ms, "pairs": 17,
                            "config": "10x3x1080x1920 8-bit, L=256, thr 0.25, relative, no uncertainty weighting, Adam x3 "
                                      "(capturable), step replayed as a CUDA graph (GraphedTrainStep, what train_icrf does for "
                                      "device-resident batches)",
                            "eager_ms_per_step": ms_eager, "eager_steps_per_s": 1e3 / ms_eager,
                            "hbm_frac_one_pass": elems * 8 / (ms * 1e-3) / 1e9 / peaks()[0],
                            "defaults_variant": {"steps_per_s": 1e3 / ms2, "ms_per_step": ms2, "eager_ms_per_step": ms2_eager,
                                                 "pairs": 24, "config": "thr 0.1, uncertainty weighting on"}}
    del val, std, model, opts
    # c3: linearity measurement, 16 exposures of 4K 16-bit
    val, std, t = ct.synthetic.make_stack(16, CHANNELS, 2160, 3840, bits=16, seed=3456, device=dev)
    theta = ct.synthetic.reference_curve(CHANNELS, LUT).to(dev)
    i_idx, j_idx, ratio = ct.common.get_valid_exposure_pairs(torch.from_numpy(t), 0.2)

    def lin():
        sums = kernels.pair_stats(val, std, i_idx, j_idx, ratio, theta, 1 / 255, 254 / 255, True, True)
        return spatial_statistics(sums, True)

    ms3 = timed(lin, 2, 10)
    # end to end through measure_linearity from pinned host memory: fp32 val + std (3.19 GB over PCIe) against the camera's
    # uint16 codes + StdSpec (0.80 GB, normalised on the device by clair_expand_codes)
    from torch.utils.data import DataLoader as _DL
    from clair_torch_b200.datasets import StdSpec as _Spec
    lin_model = ct.ICRFModelDirect(icrf=theta.clone()).to(dev)

    def one_batch_loader(batch):
        class One(torch.utils.data.Dataset):
            def __len__(self):
                return 1

            def __getitem__(self, i):
                return batch

        return _DL(One(), batch_size=None, shuffle=False)

    meta = {"exposure_time": torch.from_numpy(t)}
    codes_h = torch.round(val * 65535.0).to(torch.int32).to(torch.uint16).cpu().pin_memory()
    l_codes = one_batch_loader((torch.arange(16), codes_h, _Spec("multiplier", 0.05), meta))
    l_f32 = one_batch_loader((torch.arange(16), val.cpu().pin_memory(), std.cpu().pin_memory(), meta))

    def wall(loader, reps=3):
        ct.measure_linearity(loader, dev, True, True, lin_model)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            out_ = ct.measure_linearity(loader, dev, True, True, lin_model)
            out_[1].cpu()
        return (time.perf_counter() - t0) / reps * 1e3

    e2e3_f32, e2e3_codes = wall(l_f32), wall(l_codes)
    del val, std, codes_h, l_codes, l_f32
    torch.cuda.empty_cache()
    peak = peaks()[0]
    # per-frame linearisation of the c4 frames (24 MP, 16-bit as fp32 val + std)
    val, std, t = ct.synthetic.make_stack(3, CHANNELS, 4000, 6000, bits=16, seed=4567, device=dev)
    rad = None
    st = kernels.HdrMergeState
    one_val, one_std = val[:3].contiguous(), std[:3].contiguous()
    ms5 = timed(lambda: kernels.linearize(one_val, one_std, theta), 2, 10)
    out["linearize_c4_frames"] = {"ms_per_frame": ms5 / 3, "config": "3 frames of 3x4000x6000, f(x) and sigma",
                                  "mpixel_per_s": 3 * 24.0 / (ms5 * 1e-3),
                                  "hbm_frac": 3 * CHANNELS * 24e6 * 16 / (ms5 * 1e-3) / 1e9 / peak}
    # the same through linearize_dataset_generator from pinned host frames to pinned host results
    from torch.utils.data import DataLoader
    from clair_torch_b200.datasets import ExposureStackDataset
    hv, hs = one_val.cpu().pin_memory(), one_std.cpu().pin_memory()
    ds = ExposureStackDataset(list(hv), list(hs), list(t[:3]))
    view_collate = lambda b: (torch.tensor([b[0][0]]), b[0][1].unsqueeze(0), b[0][2].unsqueeze(0),
                              {"exposure_time": torch.tensor([b[0][3]["exposure_time"]], dtype=torch.float64)})
    lin_loader = DataLoader(ds, batch_size=1, shuffle=False, collate_fn=view_collate)
    lin_model = ct.ICRFModelDirect(icrf=theta.clone()).to(dev)
    for _ in ct.linearize_dataset_generator(lin_loader, dev, lin_model):
        pass
    t0 = time.perf_counter()
    for _ in range(3):
        for _ in ct.linearize_dataset_generator(lin_loader, dev, lin_model):
            pass
    e2e_lin = (time.perf_counter() - t0) / 9 * 1e3
    out["linearize_c4_frames"]["e2e_ms_per_frame"] = e2e_lin
    out["linearize_c4_frames"]["e2e_mpixel_per_s"] = 24.0 / (e2e_lin * 1e-3)
    out["linearize_c4_frames"]["e2e_note"] = "pinned host frame (val+std, 576 MB) in, pinned host lin+sigma (576 MB) out; clair_linearize_staged: H2D copy, kernel and D2H copy of successive bands overlapped on three streams"
    # the same frames handed over as the camera's uint16 codes: CastTo + Normalize + std = 0.05*value in the kernel's load
    from clair_torch_b200.datasets import StdSpec
    hc = torch.round(one_val * 65535.0).to(torch.int32).to(torch.uint16).cpu().pin_memory()
    ds_c = ExposureStackDataset(list(hc), StdSpec("multiplier", 0.05), list(t[:3]))
    code_collate = lambda b: (torch.tensor([b[0][0]]), b[0][1].unsqueeze(0), b[0][2],
                              {"exposure_time": torch.tensor([b[0][3]["exposure_time"]], dtype=torch.float64)})
    code_loader = DataLoader(ds_c, batch_size=1, shuffle=False, collate_fn=code_collate)
    for _ in ct.linearize_dataset_generator(code_loader, dev, lin_model):
        pass
    t0 = time.perf_counter()
    for _ in range(3):
        for _ in ct.linearize_dataset_generator(code_loader, dev, lin_model):
            pass
    e2e_codes = (time.perf_counter() - t0) / 9 * 1e3
    out["linearize_c4_frames"]["e2e_codes_ms_per_frame"] = e2e_codes
    out["linearize_c4_frames"]["e2e_codes_note"] = ("pinned host uint16 codes (144 MB) in, pinned host lin+sigma (576 MB) out, both "
                                                    "over PCIe from inside clair_linearize_codes")
    del val, std, one_val, one_std, rad, hv, hs, hc
    torch.cuda.empty_cache()
    # 8(f) rows at c1 size: dark-field mix pre-pass, flat-field correction, streaming frame statistics
    val, std, t = ct.synthetic.make_stack(N_FRAMES, CHANNELS, HEIGHT, WIDTH, bits=8, seed=99, device=dev)
    dark = torch.rand_like(val) * 0.06
    dark_std = dark * 0.1 + 1e-3
    ms6 = timed(lambda: kernels.dark_field_mix(val, std, dark, dark_std), 2, 20)
    ms6f = timed(lambda: kernels.hdr_merge_update(st(), val, std, t, theta, True, True, radiance_dtype=torch.float32,
                                                  dark=(dark, dark_std)), 2, 20)
    radiance, sigma = kernels.hdr_merge_update(st(), val, std, t, theta, True, True, radiance_dtype=torch.float32)
    flat = torch.rand_like(radiance) * 0.4 + 0.6
    flat_std = flat * 0.02
    ms7 = timed(lambda: kernels.flat_field_correct_(radiance, sigma, flat, flat_std, True), 2, 20)
    from clair_torch_b200.common.statistics import WBOMeanVar
    handler = WBOMeanVar(dim=0)
    ms8 = timed(lambda: handler.update_values(val, None, table=theta), 2, 20)
    ec1 = N_FRAMES * CHANNELS * HEIGHT * WIDTH
    out["artefacts_c1"] = {"dark_mix_ms": ms6, "dark_mix_hbm_frac": ec1 * 24 / (ms6 * 1e-3) / 1e9 / peak,
                           "dark_merge_fused_ms": ms6f,
                           "dark_merge_fused_hbm_frac": (ec1 * 16 + CHANNELS * HEIGHT * WIDTH * 8) / (ms6f * 1e-3) / 1e9 / peak,
                           "flat_field_ms": ms7, "flat_field_hbm_frac": CHANNELS * HEIGHT * WIDTH * 24 / (ms7 * 1e-3) / 1e9 / peak,
                           "note": "dark mix pre-pass: 4 input + 2 output fp32 stacks; dark_merge_fused: the c1 merge with the mix "
                                   "done in its load (4 input stacks, radiance + sigma out); flat field: value, sigma, flat, "
                                   "flat_std, in place"}
    out["frame_stats_c1"] = {"ms": ms8, "config": "5x3x1080x1920 batch, ICRF + running mean/M2 merge",
                             "hbm_frac": (ec1 * 4 + CHANNELS * HEIGHT * WIDTH * 32) / (ms8 * 1e-3) / 1e9 / peak}
    elems3 = 16 * CHANNELS * 2160 * 3840
    out["linearity_c3"] = {"ms": ms3, "pairs": int(len(i_idx)), "config": "16x3x2160x3840 16-bit, thr 0.2, relative, unc. weighting",
                           "pair_elements_per_s": len(i_idx) * CHANNELS * 2160 * 3840 / (ms3 * 1e-3),
                           "hbm_frac": elems3 * 8 / (ms3 * 1e-3) / 1e9 / peaks()[0],
                           "e2e_ms_fp32_host_batch": e2e3_f32, "e2e_ms_uint16_codes_host_batch": e2e3_codes,
                           "e2e_note": "measure_linearity(pinned host batch) incl. H2D and the D2H read of the result: fp32 val + std "
                                       "(3.19 GB) vs uint16 codes + StdSpec (0.80 GB, CastTo + Normalize + std on the device)"}
    return out


def hdr_merge_bench(cfg, key, dev, lib, rank, world, steps, warmup, e2e_steps, sample_clocks=True, burst=False, codes_e2e=False):
    """Kernel-only and end-to-end timing of the HDR merge on workload `cfg` (one rank's view).  Returns a dict; the
    timed regions are bracketed by a barrier + synchronize on both sides and reduced with MAX over ranks."""
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from torch.utils.data import DataLoader

    import clair_torch_b200 as ct

    n, c, h, w = cfg["n"], cfg["c"], cfg["h"], cfg["w"]
    n_sets = cfg["n_sets"]
    theta = ct.synthetic.reference_curve(c, LUT).to(dev)
    # device-resident stacks: each rank owns its own stacks (sharding by stack, no data-path collective)
    stacks = []
    for k in range(n_sets):
        val, std, t = ct.synthetic.make_stack(n, c, h, w, bits=cfg["bits"], seed=cfg["seed"] + 97 * rank + k, device=dev)
        stacks.append((val, std))
    radiance = torch.empty((c, h, w), dtype=torch.float32, device=dev)
    sigma = torch.empty_like(radiance)
    t_host = np.ascontiguousarray(t)
    stream = torch.cuda.current_stream(dev)

    def launch(k):
        val, std = stacks[k % n_sets]
        rc = lib.clair_hdr_merge_update(val.data_ptr(), std.data_ptr(), t_host.ctypes.data_as(ctypes.c_void_p), n,
                                        theta.data_ptr(), c, LUT, h * w, None, 1, None, None, None, 1, 1,
                                        radiance.data_ptr(), 0, sigma.data_ptr(), stream.cuda_stream)
        ct._native.check(rc, "clair_hdr_merge_update")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world > 1:
            tt = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return x

    for k in range(warmup):
        launch(k)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0 and sample_clocks:
        sampler.start()
    launches_before = ct._native.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for k in range(steps):
        launch(k)
    stop.record(stream)
    barrier()
    kernel_ms = max_over_ranks(start.elapsed_time(stop))
    launches = ct._native.launch_count() - launches_before
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    ms_per_step = kernel_ms / steps
    units_per_step = n * h * w / 1e6                                        # Mpixel*frames per stack
    burst_ms = sustained_ms = sustained_clocks = None
    sustained_launches = 0
    if burst:
        # burst: 8 launches (what one rank's share of c4, 8 stacks, looks like), started from a synchronised device that
        # has just done the K timed steps — after an idle gap the first launches run while the SM clock ramps back up;
        # sustained: >= 1 s of back-to-back launches with the clocks sampled (the board reaches its power cap)
        barrier()
        start.record(stream)
        for k in range(8):
            launch(k)
        stop.record(stream)
        barrier()
        burst_ms = max_over_ranks(start.elapsed_time(stop)) / 8
        sustained_launches = max(50, int(1100.0 / max(ms_per_step, 1e-3)))
        sampler2 = ClockSampler(dev.index or 0)
        if rank == 0:
            sampler2.start()
        barrier()
        start.record(stream)
        for k in range(sustained_launches):
            launch(k)
        stop.record(stream)
        barrier()
        sustained_ms = max_over_ranks(start.elapsed_time(stop)) / sustained_launches
        sustained_clocks = sampler2.stop() if rank == 0 else None

    # ---- end to end through the public API from pinned host memory ----
    val_h = stacks[0][0].cpu().pin_memory()
    std_h = stacks[0][1].cpu().pin_memory()
    del stacks[1:]                                                          # make room for the staging buffers
    rad_h = torch.empty((c, h, w), dtype=torch.float32).pin_memory()
    sig_h = torch.empty_like(rad_h).pin_memory()
    batch = (torch.arange(n), val_h, std_h, {"exposure_time": torch.from_numpy(t_host)})

    class OneBatch(torch.utils.data.Dataset):
        def __len__(self):
            return 1

        def __getitem__(self, i):
            return batch

    loader = DataLoader(OneBatch(), batch_size=None, shuffle=False)
    model = ct.ICRFModelDirect(icrf=theta.clone()).to(dev)

    def e2e_step():
        # pinned host batch in, pinned host results out: the copy engine streams bands of the stack to the device while
        # the kernel merges the previous band and writes radiance and sigma straight back to the pinned buffers (the bytes
        # below cross PCIe inside the timed region)
        ct.compute_hdr_image(loader, dev, model, max, radiance_dtype=torch.float32, host_out=(rad_h, sig_h))

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms) / e2e_steps)
    # what the host side can deliver: the same bytes moved by the two copy engines alone, H2D and D2H at once, on every
    # rank at the same time (pinned buffers; no kernel) -- the ceiling the end-to-end numbers are judged against
    side_in, side_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def host_ceiling_ms(src_host, n_out_bytes, reps=3):
        d_in = torch.empty(src_host.shape, dtype=src_host.dtype, device=dev)
        d_out = torch.empty(n_out_bytes, dtype=torch.uint8, device=dev)
        h_out = torch.empty(n_out_bytes, dtype=torch.uint8).pin_memory()

        def both():
            with torch.cuda.stream(side_in):
                d_in.copy_(src_host, non_blocking=True)
            with torch.cuda.stream(side_out):
                h_out.copy_(d_out, non_blocking=True)

        both()
        barrier()
        t1 = time.perf_counter()
        for _ in range(reps):
            both()
        barrier()
        return max_over_ranks((time.perf_counter() - t1) * 1e3 / reps)

    out_bytes = 2 * rad_h.numel() * 4
    both_h = torch.empty(2 * val_h.numel(), dtype=torch.float32).pin_memory()
    ceiling_fp32 = host_ceiling_ms(both_h, out_bytes)
    del both_h
    e2e_codes = None
    if codes_e2e:
        # the same stack as the camera's uint16 codes: 4x fewer bytes over PCIe, bit-identical output
        from clair_torch_b200.datasets import StdSpec
        codes_h = torch.round(stacks[0][0] * float(2 ** cfg["bits"] - 1)).to(torch.int32).to(torch.uint16 if cfg["bits"] > 8 else torch.uint8).cpu().pin_memory()
        cbatch = (torch.arange(n), codes_h, StdSpec("multiplier", 0.05), {"exposure_time": torch.from_numpy(t_host)})

        class OneCodeBatch(torch.utils.data.Dataset):
            def __len__(self):
                return 1

            def __getitem__(self, i):
                return cbatch

        cloader = DataLoader(OneCodeBatch(), batch_size=None, shuffle=False)

        def codes_step():
            ct.compute_hdr_image(cloader, dev, model, max, radiance_dtype=torch.float32, host_out=(rad_h, sig_h))

        for _ in range(3):
            codes_step()
        barrier()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            codes_step()
        barrier()
        codes_ms = max_over_ranks((time.perf_counter() - t1) * 1e3 / e2e_steps)
        ceiling_codes = host_ceiling_ms(codes_h, out_bytes)
        e2e_codes = {"value": world * units_per_step / (codes_ms * 1e-3), "unit": "Mpixel*frames/s", "ms_per_step": codes_ms,
                     "h2d_bytes_per_step": codes_h.numel() * codes_h.element_size(), "d2h_bytes_per_step": out_bytes,
                     "host_ceiling_ms": ceiling_codes, "frac_of_host_ceiling": ceiling_codes / codes_ms,
                     "api": "clair_torch_b200.compute_hdr_image(pinned uint16 code batch + StdSpec('multiplier', 0.05), host_out=pinned "
                            "buffers): CastTo + Normalize + std synthesis in the kernel's load, clair_hdr_merge_staged, 48 bands"}
        del codes_h
    peak, peak_src = peaks()
    algo_bytes = algo_bytes_per_pixel_frame(cfg) * n * h * w                  # per launch = per stack
    achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(key), "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
            "kernel": cfg["kernel"],
            # the measured peak is a 1:1 read:write copy; this kernel reads 9 bytes for every byte it writes
            # and HBM reads pay no write-turnaround, so on an unthrottled box frac can land a little above 1
            "peak_note": "peak = device copy (1 read : 1 write); the merge is 2N reads : 2 writes, so frac "
                         "slightly above 1 is a read-heavier mix than the copy, not a measurement error"}
    if burst_ms is not None:
        roof["frac_burst"] = algo_bytes / (burst_ms * 1e-3) / 1e9 / peak
        roof["burst"] = {"launches": 8, "ms_per_launch": burst_ms, "note": "8 back-to-back launches right after the timed steps"}
        roof["frac_sustained"] = algo_bytes / (sustained_ms * 1e-3) / 1e9 / peak
        roof["sustained"] = {"launches": sustained_launches, "ms_per_launch": sustained_ms, "clocks": sustained_clocks,
                             "sw_power_cap": bool(sustained_clocks and "sw_power_cap" in sustained_clocks.get("reasons", []))}
    return {
        "ms_per_step": ms_per_step, "value": world * units_per_step / (ms_per_step * 1e-3), "launches": int(launches),
        "clocks": clocks, "n_sets": n_sets, "stack_bytes": 2 * n * c * h * w * 4,
        "roofline": roof,
        "e2e": {"value": world * units_per_step / (e2e_ms * 1e-3), "unit": "Mpixel*frames/s",
                "h2d_bytes_per_step": val_h.numel() * 4 + std_h.numel() * 4, "d2h_bytes_per_step": 2 * rad_h.numel() * 4,
                "ms_per_step": e2e_ms, "host_ceiling_ms": ceiling_fp32, "frac_of_host_ceiling": ceiling_fp32 / e2e_ms,
                "host_ceiling": "the same H2D + D2H bytes moved by the two copy engines alone, both directions at once, on all "
                                "ranks at the same time (pinned buffers, no kernel)",
                "api": "clair_torch_b200.compute_hdr_image(pinned host batch, host_out=pinned buffers): clair_hdr_merge_staged, "
                       "16 bands, H2D copy overlapped with the band kernels, results stored to pinned host memory by the kernel"},
        "e2e_codes": e2e_codes,
        "theta": theta, "t_host": t_host,
    }


_RESULT_FD = None


def keep_stdout_for_the_result():
    """stdout carries ONE line, the JSON result: anything libraries print on file descriptor 1 meanwhile (NCCL's version banner
    at communicator creation, for instance) goes to stderr instead."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = json.dumps(line) + "\n"
    if _RESULT_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, text.encode())


def main():
    keep_stdout_for_the_result()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # default K: one rank's share of c4 is 8 stacks, so a job is a burst of a few milliseconds; 200 back-to-back merges
    # (~0.2 s) already run into the board's power cap (SM clock 1965 -> ~1550 MHz), longer runs only more so
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary timings (c1 merge, c2 training, c3 linearity, ...)")
    ap.add_argument("--only-dp", action="store_true", help="diagnosis: run the data-parallel c5 training section alone")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import clair_torch_b200 as ct

    local_cpus = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ct._native.load()
    for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):     # kernel tuning experiments only
        key, _, val = item.partition("=")
        ct._native.check(lib.clair_set_tuning(key.encode(), int(val)), "clair_set_tuning")

    if args.only_dp:
        dp = dp_training_metrics(dev, rank, world)
        if rank == 0:
            emit({"dp_train_c5": dp})
        if world > 1:
            dist.destroy_process_group()
        return
    cfg = WORKLOADS[args.workload]
    res = hdr_merge_bench(cfg, args.workload, dev, lib, rank, world, args.steps, args.warmup, args.e2e_steps, burst=True,
                          codes_e2e=True)
    torch.cuda.empty_cache()

    dp = None
    if world > 1 and not args.no_extras:
        dp = dp_training_metrics(dev, rank, world)       # collective: every rank takes part
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            # bounded sample on rank 0 (the other ranks idle at the final barrier); printed at every N
            cpu = cpu_baseline_hdr(cfg, budget_s=12.0 if world == 1 else 6.0)
        extra = None
        if world > 1 and dp is not None:
            extra = {"dp_train_c5": dp}
        if world == 1 and not args.no_extras:
            extra = {}
            other = "c1" if args.workload == "c4" else "c4"
            r2 = hdr_merge_bench(WORKLOADS[other], other, dev, lib, 0, 1, 300 if other == "c1" else 50, 5, 3, sample_clocks=False)
            torch.cuda.empty_cache()
            extra[f"hdr_merge_{other}"] = {"workload": WORKLOADS[other]["label"], "ms_per_step": r2["ms_per_step"],
                                           "mpixel_frames_per_s": r2["value"], "roofline": r2["roofline"],
                                           "e2e": {k: v for k, v in r2["e2e"].items() if k != "api"}}
            c1 = res if args.workload == "c1" else r2
            extra.update(secondary_metrics(dev))
            if not args.no_cpu_baseline:
                extra["icrf_train_c2"]["cpu_baseline"] = cpu_baseline_train_c2()
                extra["linearity_c3"]["cpu_baseline"] = cpu_baseline_linearity_c3()
                for key, unit in (("icrf_train_c2", "steps_per_s"), ("linearity_c3", "ms")):
                    cb = extra[key]["cpu_baseline"]
                    if cb is not None:
                        extra[key]["speedup_vs_cpu_reference"] = (extra[key]["steps_per_s"] / cb["value"] if unit == "steps_per_s"
                                                                  else cb["value"] / extra[key]["ms"])
            extra["native_ingest_c1"] = native_ingest_metrics(dev, lib, c1["theta"], c1["t_host"])
            torch.cuda.empty_cache()
            c4 = res if args.workload == "c4" else r2
            extra["native_ingest_c4"] = native_ingest_metrics(dev, lib, c4["theta"], c4["t_host"], WORKLOADS["c4"], reps=30,
                                                              e2e_reps=5, camera=True)
            torch.cuda.empty_cache()
            extra["dp_train_c5"] = dp_training_metrics(dev, 0, 1)
        res["e2e"]["host_cpus_bound"] = local_cpus
        line = {
            "metric": METRIC, "value": res["value"], "unit": "Mpixel*frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(cfg),
            "roofline": res["roofline"],
            "cpu_baseline": cpu,
            "e2e": res["e2e"],
            "e2e_codes": res["e2e_codes"],
            "gpu_launches": res["launches"], "clocks": res["clocks"], "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
