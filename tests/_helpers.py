"""Shared helpers for the test-suite (fixture loading, error norms, synthetic stacks)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


# Relative tolerance of the merged uncertainty against the REFERENCE's value, per interpolation mode.  The closed form
# (oracle and kernels, float64 where it cancels) is the exact derivative; the reference's fp32 autograd is what carries
# the noise: LINEAR ~5e-6; CATMULL ~3e-5 (cubic basis polynomials differentiated in fp32); LOOKUP ~3e-5 (f' = 0, so the
# whole uncertainty is the weight-derivative term w'(v_n - mean_B), whose two halves autograd rounds to fp32 separately).
SIGMA_TOL = {"linear": 1e-5, "catmull": 5e-5, "lookup": 1e-4}


def mode_of(z):
    """Interpolation mode a fixture was generated with ('linear' unless it says otherwise)."""
    return str(z["mode"]) if "mode" in z else "linear"


def linear_golden_names(prefix):
    """Fixtures of the reference default (LINEAR) mode only — what the C twin of the oracle and the fused kernels cover."""
    return [n for n in golden_names(prefix) if "_lookup" not in n and "_catmull" not in n]


def max_rel(a, b, floor=1e-30):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def max_abs_over_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
