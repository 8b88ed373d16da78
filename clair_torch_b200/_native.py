"""ctypes binding of libclair_b200.so (the C ABI declared in include/clair_b200.h).

PyTorch is used for device memory and streams only; every arithmetic step of the hot path happens inside
the hand-written sm_100a kernels behind this binding.  There is NO fallback: if the shared library is
missing, or no CUDA device is present, the ops raise instead of silently computing something else.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLAIR_B200_LIB") or os.path.join(_PKG_DIR, "lib", "libclair_b200.so")   # env: kernel experiments
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

ABI_VERSION = 4
MAX_FRAMES = 64
MAX_CHANNELS = 8
MAX_LUT = 1024
MAX_PAIRS = 2016
INTERP_LOOKUP = 1
INTERP_LINEAR = 2
INTERP_CATMULL = 3
CODES_PLANAR = 0
CODES_HWC_BGR = 1

_c = ctypes


class MergeDesc(ctypes.Structure):
    """clair_merge_desc of include/clair_b200.h (field order and types must match the header)."""
    _fields_ = [("struct_bytes", _c.c_uint32), ("code_bytes", _c.c_int32), ("val_dev", _c.c_void_p), ("std_dev", _c.c_void_p),
                ("std_mode", _c.c_int32), ("std_value", _c.c_float), ("code_max", _c.c_float), ("n_frames", _c.c_int32),
                ("exposure_host", _c.c_void_p), ("theta_dev", _c.c_void_p), ("n_channels", _c.c_int32), ("lut_size", _c.c_int32),
                ("interp_mode", _c.c_int32), ("gaussian_weights", _c.c_int32), ("plane", _c.c_int64), ("plane_stride", _c.c_int64),
                ("curve_row_base_host", _c.c_void_p), ("mean_state_dev", _c.c_void_p), ("wsum_state_dev", _c.c_void_p),
                ("var_state_dev", _c.c_void_p), ("is_first", _c.c_int32), ("is_final", _c.c_int32), ("radiance_f64", _c.c_int32),
                ("code_layout", _c.c_int32), ("radiance_dev", _c.c_void_p), ("sigma_dev", _c.c_void_p),
                ("dark_dev", _c.c_void_p), ("dark_std_dev", _c.c_void_p), ("height", _c.c_int32), ("width", _c.c_int32),
                ("dark_threshold", _c.c_float), ("dark_alpha", _c.c_float)]


# every symbol include/clair_b200.h declares, with its ctypes prototype
_PROTOTYPES = {
    "clair_abi_version": (_c.c_int, []),
    "clair_last_error": (_c.c_char_p, []),
    "clair_launch_count": (_c.c_uint64, []),
    "clair_set_tuning": (_c.c_int, [_c.c_char_p, _c.c_int]),
    "clair_grad_workspace_bytes": (_c.c_size_t, [_c.c_int, _c.c_int]),
    "clair_icrf_forward": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                      _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clair_icrf_backward_theta": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64,
                                             _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "clair_linearize": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                   _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clair_linearize_codes": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_float, _c.c_void_p, _c.c_int, _c.c_float, _c.c_void_p,
                                         _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_int,
                                         _c.c_void_p, _c.c_void_p]),
    "clair_linearize_staged": (_c.c_int, [_c.c_void_p] * 9 + [_c.c_int, _c.c_int, _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p,
                                          _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "clair_hdr_merge_update": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int,
                                          _c.c_int, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p,
                                          _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p,
                                          _c.c_void_p]),
    "clair_hdr_merge_codes": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_float, _c.c_void_p, _c.c_int, _c.c_float, _c.c_void_p,
                                         _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_int,
                                         _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int,
                                         _c.c_void_p, _c.c_void_p]),
    "clair_hdr_merge": (_c.c_int, [_c.POINTER(MergeDesc), _c.c_void_p]),
    "clair_hdr_merge_staged": (_c.c_int, [_c.POINTER(MergeDesc), _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clair_dark_field_mix": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int,
                                        _c.c_int, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "clair_flat_field_correct": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int,
                                            _c.c_int64, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clair_frame_stats_update": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_int,
                                            _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                            _c.c_void_p]),
    "clair_pair_stats": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p,
                                    _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                    _c.c_float, _c.c_float, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clair_pair_means": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p,
                                    _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                    _c.c_float, _c.c_float, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    "clair_pair_upstream": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                       _c.c_void_p, _c.c_void_p]),
    "clair_curve_penalties": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_float, _c.c_float, _c.c_float,
                                         _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "clair_pair_grad": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p,
                                   _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                   _c.c_float, _c.c_float, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                   _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "clair_expand_codes": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_float, _c.c_int, _c.c_float, _c.c_int64, _c.c_void_p, _c.c_void_p,
                                      _c.c_void_p]),
    "clair_copy_band_h2d": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_void_p]),
    "clair_pair_fused_doubles": (_c.c_size_t, [_c.c_int, _c.c_int]),
    "clair_pair_fused": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int,
                                    _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_float, _c.c_float, _c.c_int, _c.c_void_p,
                                    _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    "clair_pair_fused_combine": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    """libclair_b200.so is missing or does not match the header."""


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into clair_torch_b200/lib/libclair_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j6"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise NativeLibraryError(f"building libclair_b200.so failed (exit {res.returncode})")
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library and attach prototypes.  Raises NativeLibraryError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {CSRC_DIR}`.  clair_torch_b200 has no CPU or eager fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.clair_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"ABI mismatch: library {lib.clair_abi_version()}, binding {ABI_VERSION}")
        # developer knobs for measurement scripts: CLAIR_TUNE="key=value,key=value" (clair_set_tuning; unknown keys raise)
        for item in filter(None, os.environ.get("CLAIR_TUNE", "").split(",")):
            key, _, value = item.partition("=")
            if lib.clair_set_tuning(key.strip().encode(), int(value)) != 0:
                raise NativeLibraryError(f"CLAIR_TUNE: {lib.clair_last_error().decode()}")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    """Turn a non-zero return code into the exception type the reference would raise for the same mistake."""
    if rc == 0:
        return
    msg = load().clair_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise RuntimeError(f"{what}: {msg} (cudaError {rc})")


def launch_count() -> int:
    return int(load().clair_launch_count())
