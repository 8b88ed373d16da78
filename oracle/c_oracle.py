"""ctypes wrapper of oracle/_build/libclair_oracle.so (the OpenMP C twin of clair_oracle.py).

TEST INFRASTRUCTURE ONLY — see the header of clair_oracle.c.  numpy in, numpy out.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libclair_oracle.so")
_lib = None
_c = ctypes
_F = np.float32


def build():
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the C oracle failed:\n" + res.stdout + res.stderr)
    return _PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        _lib = ctypes.CDLL(_PATH)
        _lib.oracle_max_threads.restype = _c.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def use_all_host_threads() -> int:
    """torch.distributed.run exports OMP_NUM_THREADS=1 and launchers may pin a rank to a few cores: a CPU baseline
    that is meant to use the whole host undoes both.  Returns the OpenMP thread count in effect."""
    n = os.cpu_count() or 1
    try:
        os.sched_setaffinity(0, range(n))
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    lib().oracle_set_threads(int(n))
    return max_threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(_c.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=_F)


def _pairs(i_idx, j_idx, ratio):
    return (np.ascontiguousarray(i_idx, dtype=np.int32), np.ascontiguousarray(j_idx, dtype=np.int32),
            np.ascontiguousarray(ratio, dtype=np.float64))


def icrf_linear(x, theta, flat_offset=0):
    x, theta = _f32(x), _f32(theta)
    n, c = x.shape[0], x.shape[1]
    plane = int(np.prod(x.shape[2:]))
    f, fp, x0 = np.empty_like(x), np.empty_like(x), np.empty(x.shape, dtype=np.int32)
    lib().oracle_icrf_linear(_p(x), _p(theta), n, c, _c.c_int64(plane), theta.shape[1], _c.c_int64(flat_offset), _p(f),
                             _p(fp), _p(x0))
    return f, fp, x0


def icrf_lookup(x, theta):
    x, theta = _f32(x), _f32(theta)
    n, c = x.shape[0], x.shape[1]
    plane = int(np.prod(x.shape[2:]))
    y, idx = np.empty_like(x), np.empty(x.shape, dtype=np.int32)
    lib().oracle_icrf_lookup(_p(x), _p(theta), n, c, _c.c_int64(plane), theta.shape[1], _p(y), _p(idx))
    return y, idx


def linearize(val, std, theta, flat_offset=0):
    val, std, theta = _f32(val), _f32(std), _f32(theta)
    n, c = val.shape[0], val.shape[1]
    plane = int(np.prod(val.shape[2:]))
    lin, sig = np.empty_like(val), np.empty_like(val)
    lib().oracle_linearize(_p(val), _p(std), _p(theta), n, c, _c.c_int64(plane), theta.shape[1], _c.c_int64(flat_offset),
                           _p(lin), _p(sig))
    return lin, sig


def hdr_merge(val, std, exposure, theta=None, gaussian=True, batch_size=None, flat_offset=0):
    val, std, theta = _f32(val), _f32(std), _f32(theta)
    n, c = val.shape[0], val.shape[1]
    if n > 64:
        raise ValueError("the C oracle handles at most 64 frames")
    plane = int(np.prod(val.shape[2:]))
    t = np.ascontiguousarray(exposure, dtype=np.float64)
    rad = np.empty(val.shape[1:], dtype=np.float64)
    sig = None if std is None else np.empty(val.shape[1:], dtype=np.float64)
    lib().oracle_hdr_merge(_p(val), _p(std), _p(t), n, c, _c.c_int64(plane), _p(theta),
                           0 if theta is None else theta.shape[1], int(bool(gaussian)),
                           0 if batch_size is None else int(batch_size), _c.c_int64(flat_offset), _p(rad), _p(sig))
    return rad, sig


def pair_stats(val, std, i_idx, j_idx, ratio, theta=None, lo=1 / 255, hi=254 / 255, relative=True, unc_weighting=True,
               flat_offset=0):
    val, std, theta = _f32(val), _f32(std), _f32(theta)
    n, c = val.shape[0], val.shape[1]
    plane = int(np.prod(val.shape[2:]))
    pi, pj, pr = _pairs(i_idx, j_idx, ratio)
    p = pi.shape[0]
    mean, sd = np.zeros((p, c)), np.zeros((p, c))
    err = None if std is None else np.zeros((p, c))
    lib().oracle_pair_stats(_p(val), _p(std), n, c, _c.c_int64(plane), _p(pi), _p(pj), _p(pr), p, _p(theta),
                            0 if theta is None else theta.shape[1], _c.c_int64(flat_offset), _c.c_float(_F(lo)),
                            _c.c_float(_F(hi)), int(bool(relative)), int(bool(unc_weighting)), _p(mean), _p(sd), _p(err))
    return mean, sd, err


def train_grad(val, std, i_idx, j_idx, ratio, theta, lo=1 / 255, hi=254 / 255, relative=True, unc_weighting=True,
               flat_offset=0):
    val, std, theta = _f32(val), _f32(std), _f32(theta)
    n, c = val.shape[0], val.shape[1]
    plane = int(np.prod(val.shape[2:]))
    pi, pj, pr = _pairs(i_idx, j_idx, ratio)
    p = pi.shape[0]
    linloss, mean, grad = np.zeros(c), np.zeros((p, c)), np.zeros(theta.shape)
    lib().oracle_train_grad(_p(val), _p(std), n, c, _c.c_int64(plane), _p(pi), _p(pj), _p(pr), p, _p(theta), theta.shape[1],
                            _c.c_int64(flat_offset), _c.c_float(_F(lo)), _c.c_float(_F(hi)), int(bool(relative)),
                            int(bool(unc_weighting)), _p(linloss), _p(mean), _p(grad))
    return linloss, mean, grad
