"""Tensor-level launchers for the sm_100a kernels (thin: argument checks, output allocation, ctypes call).

Everything here takes CUDA tensors; torch supplies memory and the current stream, nothing else.  The
functions mirror the C ABI of include/clair_b200.h one to one.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native
from .common.errors import ArgumentTypeError

_F32 = torch.float32
_F64 = torch.float64


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_copy_streams: dict = {}


def _copy_stream(device: torch.device, which: str = "in") -> int:
    """Per-device side streams for the copy legs of the staged (copy / compute overlapped) entry points."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), which)
    st = _copy_streams.get(key)
    if st is None:
        st = _copy_streams[key] = torch.cuda.Stream(device=device)
    return st.cuda_stream


def _stack(t: torch.Tensor, name: str, allow_pinned: bool = False) -> torch.Tensor:
    """An (N, C, H, W) fp32 CUDA stack, made contiguous (MissingStdMode.CONSTANT hands over stride-0 views).
    With `allow_pinned`, a contiguous page-locked HOST tensor is accepted as well: the kernel then reads it over PCIe
    directly (every input element is read exactly once, so staging it in HBM first only adds a copy)."""
    if not isinstance(t, torch.Tensor):
        raise ArgumentTypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        if allow_pinned and t.is_pinned() and t.is_contiguous() and t.dtype == _F32 and t.dim() == 4:
            return t.detach()
        raise RuntimeError(f"{name} must live on a CUDA device: clair_torch_b200 has no CPU path")
    if t.dtype != _F32:
        raise ArgumentTypeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 4:
        raise ValueError(f"{name} must have shape (N, C, H, W), got {tuple(t.shape)}")
    return t.detach().contiguous()


def _table(theta: Optional[torch.Tensor], device, channels: int) -> Optional[torch.Tensor]:
    if theta is None:
        return None
    th = theta.detach().to(device=device, dtype=_F32).contiguous()
    if th.dim() != 2 or th.shape[0] != channels:
        raise ValueError(f"ICRF table must have shape (C={channels}, L), got {tuple(th.shape)}")
    return th


def _rows(row_base: Optional[Sequence[int]], channels: int):
    if row_base is None:
        return None, None
    arr = np.ascontiguousarray(np.asarray(row_base, dtype=np.int32))
    if arr.shape != (channels,):
        raise ValueError(f"curve_row_base must have {channels} entries")
    return arr, arr.ctypes.data_as(ctypes.c_void_p)


def shard_row_base(channels: int, full_height: int, width: int, first_row: int) -> np.ndarray:
    """curve_row_base for a row band [first_row, ...) of a full (C, full_height, width) frame."""
    c = np.arange(channels, dtype=np.int64)
    return ((c * full_height * width + first_row * width) % channels).astype(np.int32)


# ----------------------------------------------------------------------------------------------------------
def icrf_forward(x: torch.Tensor, theta: torch.Tensor, interp_mode: int = _native.INTERP_LINEAR,
                 want_derivative: bool = False, row_base=None):
    """f(x) [and autograd's df/dx] for an (N, C, H, W) stack.  models/base.py:135-182."""
    lib = _native.load()
    x = _stack(x, "image")
    n, c, h, w = x.shape
    th = _table(theta, x.device, c)
    y = torch.empty_like(x)
    dydx = torch.empty_like(x) if want_derivative else None
    keep, rows = _rows(row_base, c)
    with torch.cuda.device(x.device):
        rc = lib.clair_icrf_forward(_ptr(x), _ptr(th), _ptr(y), _ptr(dydx), n, c, h * w, th.shape[1], interp_mode,
                                    rows, _stream(x.device))
    _native.check(rc, "clair_icrf_forward")
    return (y, dydx) if want_derivative else y


def _workspace(device, channels: int, lut: int) -> torch.Tensor:
    nbytes = _native.load().clair_grad_workspace_bytes(channels, lut)
    return torch.empty(nbytes // 4, dtype=_F32, device=device)


def icrf_backward_theta(x: torch.Tensor, grad_y: torch.Tensor, channels: int, lut: int, row_base=None,
                        interp_mode: int = _native.INTERP_LINEAR) -> torch.Tensor:
    """(C, L) float64 gradient of sum(grad_y * f(x)) with respect to the table.  models/base.py:173-182 backward."""
    lib = _native.load()
    x = _stack(x, "image")
    gy = _stack(grad_y, "grad_output")
    n, c, h, w = x.shape
    grad = torch.zeros((channels, lut), dtype=_F64, device=x.device)
    ws = _workspace(x.device, channels, lut)
    keep, rows = _rows(row_base, c)
    with torch.cuda.device(x.device):
        rc = lib.clair_icrf_backward_theta(_ptr(x), _ptr(gy), _ptr(grad), n, c, h * w, lut, interp_mode, rows, _ptr(ws),
                                           ws.numel() * 4, _stream(x.device))
    _native.check(rc, "clair_icrf_backward_theta")
    return grad


def linearize(val: torch.Tensor, std, theta: torch.Tensor, row_base=None, device=None,
              pinned_out: bool = False, interp_mode: int = _native.INTERP_LINEAR, staged: Optional[bool] = None,
              bands: int = 16, code_max=None):
    """(f(x), sqrt((f'(x) std)^2)) per image of an (N, C, H, W) stack.  inference/linearization.py:94-106.

    With `device`, `val` / `std` may be pinned host tensors (read over PCIe by the kernel); with `pinned_out` the two
    results are written straight into page-locked host tensors (torch's caching host allocator), which is what
    linearize_dataset_generator hands to its consumer.  Host in AND host out is `staged` by default: three streams move
    band b+1 in, linearise band b and move band b-1 out at the same time (clair_linearize_staged).

    Integer ingest: `val` may hold raw uint8 / uint16 codes (device or pinned host memory); the kernel applies the
    reference's CastTo(float32) + Normalize(max_val=code_max) itself (code_max defaults to 255 / 65535) and `std` may be a
    `datasets.StdSpec` instead of a tensor (clair_linearize_codes; models in any InterpMode)."""
    lib = _native.load()
    if isinstance(val, torch.Tensor) and val.dtype in _CODE_DTYPES:
        return _linearize_codes(lib, val, std, theta, row_base, device, pinned_out, interp_mode, code_max)
    if code_max is not None:
        raise ValueError("code_max only applies to uint8 / uint16 value codes")
    if std is not None and not torch.is_tensor(std):
        raise ValueError("a StdSpec is evaluated by the integer-ingest kernel: pass uint8 / uint16 value codes with it")
    if interp_mode == _native.INTERP_LOOKUP and std is not None:
        # linearization.py:100-105: autograd.grad of an output that does not depend on the image
        raise RuntimeError("a LOOKUP model has no derivative with respect to the image: std images cannot be propagated "
                           "(the reference raises here too)")
    val = _stack(val, "val_batch", allow_pinned=device is not None)
    std = None if std is None else _stack(std, "std_batch", allow_pinned=device is not None)
    dev = torch.device(device) if device is not None else val.device
    n, c, h, w = val.shape
    th = _table(theta, dev, c)
    if pinned_out:
        lin = torch.empty(tuple(val.shape), dtype=_F32, pin_memory=True)
        sigma = torch.empty(tuple(val.shape), dtype=_F32, pin_memory=True)
    else:
        lin = torch.empty(tuple(val.shape), dtype=_F32, device=dev)
        sigma = torch.empty(tuple(val.shape), dtype=_F32, device=dev)
    keep, rows = _rows(row_base, c)
    host_to_host = pinned_out and not val.is_cuda and (std is None or not std.is_cuda)
    if staged is None:
        staged = host_to_host
    if staged and not host_to_host:
        raise ValueError("staged=True is for pinned host inputs with pinned_out=True")
    with torch.cuda.device(dev):
        if staged:
            stage = [torch.empty(tuple(val.shape), dtype=_F32, device=dev) if need else None for need in (True, std is not None, True, True)]
            rc = lib.clair_linearize_staged(_ptr(val), _ptr(std), _ptr(lin), _ptr(sigma), *(_ptr(b) for b in stage), _ptr(th), n, c,
                                            h * w, th.shape[1], int(interp_mode), rows, int(bands), _copy_stream(dev, "in"),
                                            _copy_stream(dev, "out"), _stream(dev))
            _native.check(rc, "clair_linearize_staged")
            return lin, sigma
        rc = lib.clair_linearize(_ptr(val), _ptr(std), _ptr(th), _ptr(lin), _ptr(sigma), n, c, h * w, th.shape[1],
                                 int(interp_mode), rows, _stream(dev))
    _native.check(rc, "clair_linearize")
    return lin, sigma


def _linearize_codes(lib, val, std, theta, row_base, device, pinned_out, interp_mode, code_max):
    """linearize() for uint8 / uint16 codes: CastTo + Normalize and the std synthesis happen in the kernel's load."""
    val = _code_stack(val, device is not None)
    if interp_mode == _native.INTERP_LOOKUP and std is not None:
        raise RuntimeError("a LOOKUP model has no derivative with respect to the image: std images cannot be propagated "
                           "(the reference raises here too)")
    std_mode, std_value = 0, 0.0
    if std is not None and not torch.is_tensor(std):
        if not hasattr(std, "mode") or std.mode not in _STD_MODES:
            raise ArgumentTypeError(f"std_batch must be a tensor, None or a StdSpec, got {type(std)}")
        std_mode, std_value, std = _STD_MODES[std.mode], float(np.float32(std.value)), None
    elif std is not None:
        std = _stack(std, "std_batch", allow_pinned=device is not None)
        if std.shape != val.shape:
            raise ValueError("std_batch must have the same shape as val_batch")
        std_mode = 1
    dev = torch.device(device) if device is not None else val.device
    n, c, h, w = val.shape
    if (h * w) % 4 != 0:
        raise ValueError("integer ingest needs H*W to be a multiple of 4")
    th = _table(theta, dev, c)
    code_bytes, default_max = _CODE_DTYPES[val.dtype]
    kw = dict(dtype=_F32, pin_memory=True) if pinned_out else dict(dtype=_F32, device=dev)
    lin, sigma = torch.empty(tuple(val.shape), **kw), torch.empty(tuple(val.shape), **kw)
    keep, rows = _rows(row_base, c)
    with torch.cuda.device(dev):
        rc = lib.clair_linearize_codes(_ptr(val), code_bytes, float(default_max if code_max is None else code_max), _ptr(std),
                                       std_mode, std_value, _ptr(th), _ptr(lin), _ptr(sigma), n, c, h * w, th.shape[1],
                                       int(interp_mode), rows, _stream(dev))
    _native.check(rc, "clair_linearize_codes")
    return lin, sigma


def copy_band_to_device(dst: torch.Tensor, src_host: torch.Tensor, first_row: int, stream: torch.cuda.Stream) -> None:
    """Rows [first_row, first_row + dst.shape[-2]) of every (frame, channel) slab of the page-locked (N, C, H, W) host stack into
    the dense device buffer `dst` (N, C, rows, W), as one strided copy on `stream` (clair_copy_band_h2d).  The caller keeps
    `src_host` alive until the stream has passed the copy."""
    lib = _native.load()
    n, c, h, w = src_host.shape
    rows = dst.shape[-2]
    if not (src_host.is_pinned() and src_host.is_contiguous() and dst.is_cuda and dst.is_contiguous() and dst.dtype == src_host.dtype
            and tuple(dst.shape) == (n, c, rows, w) and 0 <= first_row and first_row + rows <= h):
        raise ValueError("copy_band_to_device: a contiguous page-locked (N, C, H, W) source and a dense (N, C, rows, W) device band")
    size = src_host.element_size()
    with torch.cuda.device(dst.device):
        rc = lib.clair_copy_band_h2d(_ptr(dst), _ptr(src_host), n * c, h * w * size, first_row * w * size, rows * w * size,
                                     ctypes.c_void_p(stream.cuda_stream))
    _native.check(rc, "clair_copy_band_h2d")


def expand_codes(codes: torch.Tensor, std=None, code_max=None, device=None):
    """(fp32 value stack, fp32 std stack | None) from uint8 / uint16 camera codes: CastTo(float32) + Normalize(max_val=code_max)
    and, for a `datasets.StdSpec`, the synthesised std, evaluated on the device (clair_expand_codes; bit-identical to the
    reference's CPU transforms).  `codes` may sit in (pinned) host memory: it crosses PCIe as 1-2 bytes per sample."""
    lib = _native.load()
    if not torch.is_tensor(codes) or codes.dtype not in _CODE_DTYPES:
        raise TypeError("codes must be a uint8 / uint16 tensor")
    dev = torch.device(device) if device is not None else codes.device
    if dev.type != "cuda":
        raise RuntimeError("expand_codes runs on a CUDA device: clair_torch_b200 has no CPU path")
    codes = codes.detach().contiguous().to(device=dev, non_blocking=True)
    if codes.numel() % 4 != 0:
        raise ValueError("integer ingest needs the element count to be a multiple of 4")
    code_bytes, default_max = _CODE_DTYPES[codes.dtype]
    std_mode, std_value = 0, 0.0
    if std is not None:
        if not hasattr(std, "mode") or std.mode not in _STD_MODES:
            raise TypeError(f"std must be None or a StdSpec, got {type(std)}")
        std_mode, std_value = _STD_MODES[std.mode], float(np.float32(std.value))
    val = torch.empty(tuple(codes.shape), dtype=_F32, device=dev)
    sd = torch.empty(tuple(codes.shape), dtype=_F32, device=dev) if std_mode else None
    with torch.cuda.device(dev):
        rc = lib.clair_expand_codes(_ptr(codes), code_bytes, float(default_max if code_max is None else code_max), std_mode, std_value,
                                    codes.numel(), _ptr(val), _ptr(sd), _stream(dev))
    _native.check(rc, "clair_expand_codes")
    return val, sd


class HdrMergeState:
    """Running (mean, sum of weights, variance) of the merge — WBOMean's state (common/statistics.py:27-29)
    plus hdr_merge.py's running_variance, kept as device buffers between DataLoader batches."""

    def __init__(self):
        self.mean: Optional[torch.Tensor] = None     # (C, H, W) float64
        self.wsum: Optional[torch.Tensor] = None     # (C, H, W) float32
        self.var: Optional[torch.Tensor] = None      # (C, H, W) float32
        self.batches = 0

    def _ensure(self, shape, device, with_var: bool):
        if self.mean is None:
            self.mean = torch.empty(shape, dtype=_F64, device=device)
            self.wsum = torch.empty(shape, dtype=_F32, device=device)
        if with_var and self.var is None:
            self.var = torch.empty(shape, dtype=_F32, device=device)


_CODE_DTYPES = {torch.uint8: (1, 255.0), torch.uint16: (2, 65535.0)}
_STD_MODES = {"multiplier": 2, "constant": 3}


def _code_stack(t: torch.Tensor, allow_pinned: bool) -> torch.Tensor:
    if t.dim() != 4:
        raise ValueError(f"val_batch must have shape (N, C, H, W), got {tuple(t.shape)}")
    if not t.is_cuda and not (allow_pinned and t.is_pinned() and t.is_contiguous()):
        raise RuntimeError("val_batch must live on a CUDA device (or be pinned host memory): clair_torch_b200 has no CPU path")
    return t.detach().contiguous()


def hdr_merge_update(state: HdrMergeState, val: torch.Tensor, std, exposure,
                     theta: Optional[torch.Tensor], gaussian_weights: bool, is_final: bool,
                     radiance_dtype: torch.dtype = _F64, row_base=None, device=None, host_out=None, code_max=None,
                     interp_mode: int = _native.INTERP_LINEAR, staged: Optional[bool] = None, bands: Optional[int] = None, dark=None,
                     code_layout: str = "planar"):
    """One batch of compute_hdr_image (inference/hdr_merge.py:95-128).  Returns (radiance, sigma) when
    `is_final`, else None.  `exposure` is the collated float64 'exposure_time' (host tensor, array or list).

    `val` / `std` may be pinned host tensors (`device` then names the GPU that runs the kernel), and
    `host_out=(radiance, sigma)` pinned host buffers make the kernel write its results straight to host memory.
    fp32 host stacks are `staged` by default: the copy engine streams bands of the planes into device buffers while the
    kernel merges the previous band (clair_hdr_merge_staged); `staged=False` (the default for integer codes) lets the
    kernel read the host memory itself (zero-copy).

    `code_layout="hwc_bgr"`: the integer codes are (N, H, W, 3) in OpenCV's BGR order, i.e. what `cv2.imread` returns,
    stacked; the reference's CvToTorch transform (BGR -> RGB, HWC -> CHW) then happens inside the kernel's load.  Outputs,
    std tensors and running state stay planar (3, H, W) RGB.

    `interp_mode`: the model's InterpMode (LINEAR: fused fast kernels; LOOKUP / CATMULL: the all-modes kernel).

    `dark=(dark_val, dark_std)` (stacks shaped like `val`, or with a leading 1) fuses the dark-field correction of
    inference/hdr_merge.py:76-92,117-126 into the kernel's load when `can_fuse_dark` says so; otherwise run
    `dark_field_mix` first and pass its outputs.

    Integer ingest: `val` may hold raw uint8 / uint16 codes; the kernel then applies the reference's
    CastTo(float32) + Normalize(max_val=code_max, min_val=0) itself (code_max defaults to 255 / 65535), and `std` may be
    a `datasets.StdSpec` (std = value * m, or a constant) instead of a tensor."""
    lib = _native.load()
    if not isinstance(val, torch.Tensor):
        raise ArgumentTypeError(f"val_batch must be a torch.Tensor, got {type(val)}")
    codes = val.dtype in _CODE_DTYPES
    if codes:
        val = _code_stack(val, device is not None)
    else:
        val = _stack(val, "val_batch", allow_pinned=device is not None)
    std_mode, std_value = 0, 0.0
    if std is not None and not torch.is_tensor(std):
        if not hasattr(std, "mode") or std.mode not in _STD_MODES:
            raise ArgumentTypeError(f"std_batch must be a tensor, None or a StdSpec, got {type(std)}")
        std_mode, std_value, std = _STD_MODES[std.mode], float(np.float32(std.value)), None
        if not codes:
            raise ValueError("a StdSpec is evaluated by the integer-ingest kernel: pass uint8 / uint16 value codes with it")
    elif std is not None:
        std = _stack(std, "std_batch", allow_pinned=device is not None)
        std_mode = 1
    has_std = std_mode != 0
    dev = torch.device(device) if device is not None else val.device
    if val.is_cuda and val.device != dev:
        raise ValueError("val_batch lives on a different device than the one requested")
    if std is not None and code_layout == "planar" and std.shape != val.shape:
        raise ValueError("std_batch must have the same shape as val_batch")
    if std is not None and code_layout != "planar" and tuple(std.shape) != (val.shape[0], 3, val.shape[1], val.shape[2]):
        raise ValueError("with code_layout='hwc_bgr' a std tensor must be planar (N, 3, H, W)")
    if code_max is not None and not codes:
        raise ValueError("code_max only applies to uint8 / uint16 value codes")
    if code_layout not in ("planar", "hwc_bgr"):
        raise ValueError(f"code_layout must be 'planar' or 'hwc_bgr', got {code_layout!r}")
    hwc = code_layout == "hwc_bgr"
    if hwc:
        if not codes or val.shape[3] != 3:
            raise ValueError("code_layout='hwc_bgr' takes uint8 / uint16 codes of shape (N, H, W, 3)")
        n, h, w, c = val.shape
        if (h * w) % 4 != 0:
            raise ValueError("code_layout='hwc_bgr' needs H*W to be a multiple of 4")
    else:
        n, c, h, w = val.shape
    if torch.is_tensor(exposure):
        exposure = exposure.detach().cpu().numpy()
    t = np.ascontiguousarray(np.asarray(exposure, dtype=np.float64).reshape(-1))
    if t.shape[0] != n:
        raise ValueError(f"{n} frames but {t.shape[0]} exposure times")
    th = _table(theta, dev, c)
    lut = 0 if th is None else th.shape[1]
    is_first = state.batches == 0
    if has_std and not is_first and state.var is None:
        raise ValueError("std images appeared after a batch without them")
    if not (is_first and is_final):
        state._ensure((c, h, w), dev, has_std)
    radiance = sigma = None
    if is_final:
        if radiance_dtype not in (_F32, _F64):
            raise ArgumentTypeError("radiance_dtype must be torch.float32 or torch.float64")
        if host_out is not None:
            radiance, sigma = host_out
            ok = (radiance.is_pinned() and radiance.is_contiguous() and tuple(radiance.shape) == (c, h, w)
                  and radiance.dtype == radiance_dtype)
            if has_std:
                ok = ok and sigma is not None and sigma.is_pinned() and sigma.is_contiguous() and sigma.dtype == _F32 \
                    and tuple(sigma.shape) == (c, h, w)
            if not ok:
                raise ValueError("host_out must be pinned, contiguous (C, H, W) tensors of the output dtypes")
            if not has_std:
                sigma = None
        else:
            radiance = torch.empty((c, h, w), dtype=radiance_dtype, device=dev)
            if has_std:
                sigma = torch.empty((c, h, w), dtype=_F32, device=dev)
    keep, rows = _rows(row_base, c)
    desc = _native.MergeDesc()
    desc.struct_bytes = ctypes.sizeof(_native.MergeDesc)
    desc.code_bytes, desc.code_max = 0, 1.0
    if codes:
        desc.code_bytes, default_max = _CODE_DTYPES[val.dtype]
        desc.code_max = float(default_max if code_max is None else code_max)
    desc.std_mode, desc.std_value = std_mode, std_value
    desc.n_frames, desc.exposure_host = n, t.ctypes.data
    desc.theta_dev, desc.n_channels, desc.lut_size, desc.interp_mode = _ptr(th), c, lut, int(interp_mode)
    desc.gaussian_weights, desc.plane, desc.plane_stride = int(bool(gaussian_weights)), h * w, 0
    desc.curve_row_base_host = keep.ctypes.data if keep is not None else None
    desc.mean_state_dev, desc.wsum_state_dev, desc.var_state_dev = _ptr(state.mean), _ptr(state.wsum), _ptr(state.var)
    desc.is_first, desc.is_final, desc.radiance_f64 = int(is_first), int(is_final), int(radiance_dtype == _F64)
    desc.code_layout = _native.CODES_HWC_BGR if hwc else _native.CODES_PLANAR
    desc.radiance_dev, desc.sigma_dev = _ptr(radiance), _ptr(sigma)
    if dark is not None:
        dark_val, dark_std = dark
        if not can_fuse_dark(val, std, dark_std, interp_mode if th is not None else _native.INTERP_LINEAR):
            raise ValueError("this batch cannot take the fused dark-field path (see can_fuse_dark): use dark_field_mix first")
        dark_val, dark_std = _dark_stack(dark_val, val, "dark_field_val"), _dark_stack(dark_std, val, "dark_field_std")
        desc.dark_dev, desc.dark_std_dev, desc.height, desc.width = _ptr(dark_val), _ptr(dark_std), h, w
        desc.dark_threshold, desc.dark_alpha = DARK_THRESHOLD, DARK_ALPHA
    on_host = not val.is_cuda
    big = val.numel() * val.element_size() >= (128 << 20)
    if staged is None:
        # measured (profiles/README.md): fp32 stacks are input-bound and gain from the copy engine's faster host reads;
        # small integer stacks (1-2 B in, 8 B out per pixel: c1 is 31 MB) are quicker read in place by the kernel, large ones
        # (c4: 1.3 GB of codes) go through the copy engine in many bands — 26.2 ms against 29.6 ms read in place, with the
        # two copy engines alone needing 24.6 ms for the same traffic (scratch/e2e_codes_c4.py);
        # interleaved codes are always staged: in place, each of the three channel blocks would pull all three channels' bytes
        staged = on_host and (not codes or hwc or big)
    if staged and not on_host:
        raise ValueError("staged=True is for pinned host stacks")
    if bands is None:
        # few large bands when the input is small (scratch/hwc_sweep.py), many when it is large: only the first band's copy
        # and the last band's kernel are not overlapped
        bands = (48 if big else 4) if codes else 16
    with torch.cuda.device(dev):
        if staged:
            stage_val = torch.empty(val.shape, dtype=val.dtype, device=dev)
            stage_std = torch.empty(std.shape, dtype=_F32, device=dev) if std is not None else None
            desc.val_dev, desc.std_dev = _ptr(stage_val), _ptr(stage_std)
            rc = lib.clair_hdr_merge_staged(ctypes.byref(desc), _ptr(val), _ptr(std), int(bands), _copy_stream(dev),
                                            _stream(dev))
            what = "clair_hdr_merge_staged"
        else:
            desc.val_dev, desc.std_dev = _ptr(val), _ptr(std)
            rc = lib.clair_hdr_merge(ctypes.byref(desc), _stream(dev))
            what = "clair_hdr_merge"
    _native.check(rc, what)
    state.batches += 1
    return (radiance, sigma) if is_final else None


# ----------------------------------------------------------------------------------------------------------
def _pair_arrays(i_idx, j_idx, ratio):
    def host(a, dt):
        if torch.is_tensor(a):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(a).reshape(-1).astype(dt))
    pi, pj, pr = host(i_idx, np.int32), host(j_idx, np.int32), host(ratio, np.float64)
    if not (pi.shape == pj.shape == pr.shape):
        raise ValueError("i_idx, j_idx and ratio_pairs must have the same length")
    return pi, pj, pr


def _check_lookup_std(theta, std, interp_mode):
    if theta is not None and std is not None and interp_mode == _native.INTERP_LOOKUP:
        # measure_linearity.py:57-63 / icrf_training.py:117-124: autograd.grad of an output that does not depend on the image
        raise RuntimeError("a LOOKUP model has no derivative with respect to the image: std images cannot be propagated "
                           "(the reference raises here too)")


def pair_stats(val: torch.Tensor, std: Optional[torch.Tensor], i_idx, j_idx, ratio,
               theta: Optional[torch.Tensor], valid_lo: float, valid_hi: float, relative: bool,
               unc_weighting: bool, row_base=None, out: Optional[torch.Tensor] = None,
               means_only: bool = False, interp_mode: int = _native.INTERP_LINEAR) -> torch.Tensor:
    """(P, C, 5) float64 sums [sum MWt, sum MWt l, sum MWt l^2, sum M err, sum M] for one batch
    (only the first two when `means_only`, the training-step variant).  `interp_mode` is the InterpMode of the model
    `theta` belongs to (any of the three).
    training/losses.py:13-108, common/general_functions.py:118-178,276-312."""
    lib = _native.load()
    val = _stack(val, "val_batch")
    std = None if std is None else _stack(std, "std_batch")
    _check_lookup_std(theta, std, interp_mode)
    n, c, h, w = val.shape
    pi, pj, pr = _pair_arrays(i_idx, j_idx, ratio)
    p = pi.shape[0]
    th = _table(theta, val.device, c)
    lut = 0 if th is None else th.shape[1]
    sums = torch.zeros((p, c, 5), dtype=_F64, device=val.device) if out is None else out
    keep, rows = _rows(row_base, c)
    with torch.cuda.device(val.device):
        entry = lib.clair_pair_means if means_only else lib.clair_pair_stats
        rc = entry(
            _ptr(val), _ptr(std), n, c, h * w, pi.ctypes.data_as(ctypes.c_void_p),
            pj.ctypes.data_as(ctypes.c_void_p), pr.ctypes.data_as(ctypes.c_void_p), p, _ptr(th), lut, int(interp_mode), rows,
            float(np.float32(valid_lo)), float(np.float32(valid_hi)), int(bool(relative)), int(bool(unc_weighting)),
            _ptr(sums), _stream(val.device))
    _native.check(rc, "clair_pair_stats")
    return sums


def pair_grad(val: torch.Tensor, std: Optional[torch.Tensor], i_idx, j_idx, ratio, theta: torch.Tensor,
              valid_lo: float, valid_hi: float, relative: bool, unc_weighting: bool, upstream: torch.Tensor,
              mean: torch.Tensor, row_base=None, out: Optional[torch.Tensor] = None,
              interp_mode: int = _native.INTERP_LINEAR) -> torch.Tensor:
    """(C, L) float64 gradient of the linearity loss with respect to the table."""
    lib = _native.load()
    val = _stack(val, "val_batch")
    std = None if std is None else _stack(std, "std_batch")
    _check_lookup_std(theta, std, interp_mode)
    n, c, h, w = val.shape
    pi, pj, pr = _pair_arrays(i_idx, j_idx, ratio)
    p = pi.shape[0]
    th = _table(theta, val.device, c)
    lut = th.shape[1]
    up = upstream.detach().to(device=val.device, dtype=_F64).contiguous()
    mn = mean.detach().to(device=val.device, dtype=_F64).contiguous()
    if tuple(up.shape) != (p, c) or tuple(mn.shape) != (p, c):
        raise ValueError("upstream and mean must have shape (P, C)")
    grad = torch.zeros((c, lut), dtype=_F64, device=val.device) if out is None else out
    ws = _workspace(val.device, c, lut)
    keep, rows = _rows(row_base, c)
    with torch.cuda.device(val.device):
        rc = lib.clair_pair_grad(
            _ptr(val), _ptr(std), n, c, h * w, pi.ctypes.data_as(ctypes.c_void_p),
            pj.ctypes.data_as(ctypes.c_void_p), pr.ctypes.data_as(ctypes.c_void_p), p, _ptr(th), lut, int(interp_mode), rows,
            float(np.float32(valid_lo)), float(np.float32(valid_hi)), int(bool(relative)), int(bool(unc_weighting)),
            _ptr(up), _ptr(mn), _ptr(grad), _ptr(ws), ws.numel() * 4, _stream(val.device))
    _native.check(rc, "clair_pair_grad")
    return grad


def pair_fused(val: torch.Tensor, i_idx, j_idx, ratio, theta: torch.Tensor, valid_lo: float, valid_hi: float, relative: bool,
               row_base=None, out: Optional[torch.Tensor] = None, interp_mode: int = _native.INTERP_LINEAR) -> torch.Tensor:
    """Single-pass statistics + un-normalised table gradient of ONE exposure pair without uncertainty weighting
    (clair_pair_fused): float64 buffer [(1, C, 5) sums | (C, C, L) tables], accumulated into `out` when given."""
    lib = _native.load()
    val = _stack(val, "val_batch")
    n, c, h, w = val.shape
    pi, pj, pr = _pair_arrays(i_idx, j_idx, ratio)
    th = _table(theta, val.device, c)
    lut = th.shape[1]
    size = lib.clair_pair_fused_doubles(c, lut)
    buf = torch.zeros(size, dtype=_F64, device=val.device) if out is None else out
    ws = _workspace(val.device, c, lut)
    keep, rows = _rows(row_base, c)
    with torch.cuda.device(val.device):
        rc = lib.clair_pair_fused(_ptr(val), n, c, h * w, pi.ctypes.data_as(ctypes.c_void_p), pj.ctypes.data_as(ctypes.c_void_p),
                                  pr.ctypes.data_as(ctypes.c_void_p), pi.shape[0], _ptr(th), lut, int(interp_mode), rows,
                                  float(np.float32(valid_lo)), float(np.float32(valid_hi)), int(bool(relative)), _ptr(buf), _ptr(ws),
                                  ws.numel() * 4, _stream(val.device))
    _native.check(rc, "clair_pair_fused")
    return buf


def can_fuse_pair(val: torch.Tensor, std, n_pairs: int, unc_weighting: bool) -> bool:
    """Whether a training step can take the single-pass kernel: exactly one exposure pair, weights that do not depend on
    the uncertainty, an even plane and an 8-byte aligned fp32 stack."""
    return (n_pairs == 1 and not (unc_weighting and std is not None) and torch.is_tensor(val) and val.is_cuda and val.dtype == _F32
            and val.dim() == 4 and (val.shape[2] * val.shape[3]) % 2 == 0 and val.is_contiguous() and val.data_ptr() % 8 == 0)


def pair_fused_combine(fused: torch.Tensor, channels: int, lut: int, want_grad: bool = True):
    """(linloss (C,), mean (1, C), grad (C, L) | None), float64, from the (all-reduced) buffer of pair_fused."""
    lib = _native.load()
    dev = fused.device
    linloss = torch.empty((channels,), dtype=_F64, device=dev)
    mean = torch.empty((1, channels), dtype=_F64, device=dev)
    grad = torch.zeros((channels, lut), dtype=_F64, device=dev) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.clair_pair_fused_combine(_ptr(fused), channels, lut, _ptr(linloss), _ptr(mean), _ptr(grad), _stream(dev))
    _native.check(rc, "clair_pair_fused_combine")
    return linloss, mean, grad


def pair_upstream(sums: torch.Tensor):
    """(linloss (C,), mean (P,C), upstream (P,C), mean_for_grad (P,C)) float64 from the (P,C,5) sums, on the device."""
    lib = _native.load()
    p, c, _ = sums.shape
    dev = sums.device
    linloss = torch.empty((c,), dtype=_F64, device=dev)
    mean = torch.empty((p, c), dtype=_F64, device=dev)
    upstream = torch.empty((p, c), dtype=_F64, device=dev)
    mean_for_grad = torch.empty((p, c), dtype=_F64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.clair_pair_upstream(_ptr(sums), p, c, _ptr(linloss), _ptr(mean), _ptr(upstream), _ptr(mean_for_grad),
                                     _stream(dev))
    _native.check(rc, "clair_pair_upstream")
    return linloss, mean, upstream, mean_for_grad


def curve_penalties(theta: torch.Tensor, alpha: float, beta: float, gamma: float, delta: float, grad: torch.Tensor):
    """Per-channel weighted sum of the four curve penalties (C,) float64; their gradient is ADDED to `grad` (C,L) float64."""
    lib = _native.load()
    th = theta.detach().to(dtype=_F32).contiguous()
    if not th.is_cuda:
        raise RuntimeError("the ICRF table must live on a CUDA device")
    c, lut = th.shape
    pen = torch.empty((c,), dtype=_F64, device=th.device)
    with torch.cuda.device(th.device):
        rc = lib.clair_curve_penalties(_ptr(th), c, lut, float(alpha), float(beta), float(gamma), float(delta), _ptr(pen),
                                       _ptr(grad), _stream(th.device))
    _native.check(rc, "clair_curve_penalties")
    return pen


DARK_THRESHOLD, DARK_ALPHA = 0.05, 50.0     # inference/hdr_merge.py:90, common/general_functions.py:442


def _dark_stack(t: torch.Tensor, like: torch.Tensor, name: str) -> torch.Tensor:
    t = _stack(t.to(like.device), name)
    if t.shape[0] == 1 and like.shape[0] > 1:
        t = t.expand(like.shape[0], -1, -1, -1).contiguous()
    if t.shape != like.shape:
        raise ValueError(f"{name} batch dimension must be 1 or {like.shape[0]}, got {t.shape[0]}")
    return t


def can_fuse_dark(val, std, dark_std, interp_mode: int = _native.INTERP_LINEAR) -> bool:
    """Whether hdr_merge_update can apply the dark-field correction inside the merge kernel (clair_merge_desc.dark_dev):
    device-resident fp32 images with std and dark std, LINEAR model (or none), at most 8 frames, even width."""
    return (torch.is_tensor(val) and val.is_cuda and val.dtype == _F32 and val.dim() == 4 and torch.is_tensor(std)
            and dark_std is not None and interp_mode == _native.INTERP_LINEAR and val.shape[0] <= 8
            and val.shape[3] % 2 == 0 and val.shape[2] >= 2)


def dark_field_mix(val: torch.Tensor, std: Optional[torch.Tensor], dark: torch.Tensor, dark_std: Optional[torch.Tensor],
                   threshold: float = DARK_THRESHOLD, alpha: float = DARK_ALPHA):
    """(mixed images, effective std) of the dark-field correction: conditional_gaussian_blur plus the collapse of the
    image-std and dark-std variance terms into one per-element std (see clair_dark_field_mix in the header)."""
    lib = _native.load()
    val = _stack(val, "val_batch")
    dark = _stack(dark, "dark_field_val")
    n, c, h, w = val.shape
    if dark.shape[0] == 1 and n > 1:
        dark = dark.expand(n, -1, -1, -1).contiguous()
    if dark.shape != val.shape:
        raise ValueError(f"mask_map batch dimension must be 1 or {n}, got {dark.shape[0]}")
    want_std = std is not None
    if want_std:
        if dark_std is None:
            raise ValueError("dark-field std images are required when the value images carry std images")
        std = _stack(std, "std_batch")
        dark_std = _stack(dark_std, "dark_field_std")
        if dark_std.shape[0] == 1 and n > 1:
            dark_std = dark_std.expand(n, -1, -1, -1).contiguous()
    out_val = torch.empty_like(val)
    out_std = torch.empty_like(val) if want_std else None
    with torch.cuda.device(val.device):
        rc = lib.clair_dark_field_mix(_ptr(val), _ptr(std) if want_std else None, _ptr(dark), _ptr(dark_std) if want_std else None,
                                      n, c, h, w, float(threshold), float(alpha), _ptr(out_val), _ptr(out_std),
                                      _stream(val.device))
    _native.check(rc, "clair_dark_field_mix")
    return out_val, out_std


def flat_field_correct_(value: torch.Tensor, sigma: Optional[torch.Tensor], flat: torch.Tensor,
                        flat_std: Optional[torch.Tensor], mean_in_graph: bool):
    """In-place flat-field correction of `value` ((C,H,W) or (N,C,H,W), fp32 or fp64) and of its std `sigma` (fp32)."""
    lib = _native.load()
    if not value.is_cuda or not value.is_contiguous():
        raise RuntimeError("flat-field correction works in place on contiguous CUDA tensors")
    shape = value.shape if value.dim() == 4 else (1,) + tuple(value.shape)
    n, c, h, w = shape
    f = flat.detach().to(device=value.device, dtype=_F32).reshape(-1, h, w).contiguous()
    if f.shape[0] != c:
        raise ValueError(f"flat field must have shape (C={c}, H, W) (or (1, C, H, W)), got {tuple(flat.shape)}")
    fs = None if flat_std is None else flat_std.detach().to(device=value.device, dtype=_F32).reshape(-1, h, w).contiguous()
    if sigma is not None and (not sigma.is_cuda or sigma.dtype != _F32 or not sigma.is_contiguous() or sigma.numel() != value.numel()):
        raise ValueError("sigma must be a contiguous fp32 CUDA tensor of the shape of value")
    scratch = torch.empty(2 * c, dtype=_F64, device=value.device)
    with torch.cuda.device(value.device):
        rc = lib.clair_flat_field_correct(_ptr(value), int(value.dtype == _F64), _ptr(sigma), _ptr(f), _ptr(fs), n, c, h * w,
                                          int(bool(mean_in_graph)), _ptr(scratch), _stream(value.device))
    _native.check(rc, "clair_flat_field_correct")
    return value, sigma
