// Shared device helpers for the clair_b200 kernels (sm_100a).
//
// The ICRF table theta (C, L) fp32 is tiny (3 KB at the reference default C=3, L=256), so every kernel
// stages it in shared memory as (theta[u][k], theta[u][min(k+1, L-1)]) float2 pairs: one 64-bit LDS returns
// both taps of the linear interpolation of models/base.py:160-182.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/clair_b200.h"

namespace clair {

// (flat index of element (c, 0)) mod C for each channel, 4 bits per channel (C <= 8 so every value is < 8).
// Packed into one word so a dynamic channel index never forces the kernel-parameter struct into local memory.
struct CurveRows {
    uint32_t packed;
    __host__ __device__ __forceinline__ int base(int c) const { return static_cast<int>((packed >> (4 * c)) & 15u); }
};

struct FrameScale {
    float inv_t[CLAIR_MAX_FRAMES];      // 1 / exposure time, rounded once from float64
};

// ---- streaming 128-bit / 32-bit global access (every input element is read exactly once) ----
template <int VEC> struct Pack;
template <> struct Pack<4> { float v[4]; };
template <> struct Pack<2> { float v[2]; };
template <> struct Pack<1> { float v[1]; };

template <int VEC>
__device__ __forceinline__ Pack<VEC> load_stream(const float *p) {
    Pack<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else if constexpr (VEC == 2) {
        float2 t = __ldcs(reinterpret_cast<const float2 *>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = __ldcs(p);
    }
    return r;
}

template <int VEC>
__device__ __forceinline__ void store_stream(float *p, const Pack<VEC> &r) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<float2 *>(p), make_float2(r.v[0], r.v[1]));
    } else {
        __stcs(p, r.v[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void store_stream_f64(double *p, const double (&r)[VEC]) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(r[0], r[1]));
        __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(r[2], r[3]));
    } else if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<double2 *>(p), make_double2(r[0], r[1]));
    } else {
        __stcs(p, r[0]);
    }
}

// ---- table staging ----
// smem layout: tab[u * L + k] = (theta[u][k], theta[u][min(k+1, L-1)])
__device__ __forceinline__ void stage_curve_pairs(float2 *tab, const float *__restrict__ theta, int n_rows, int L) {
    const int total = n_rows * L;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i % L;
        const float a = __ldg(theta + i);
        const float b = (k + 1 < L) ? __ldg(theta + i + 1) : a;
        tab[i] = make_float2(a, b);
    }
}

// ---- ICRF evaluation, bit-exact with the reference's fp32 op sequence (no FMA contraction) ----
struct IcrfTap {
    float f;     // g0*(1-w) + g1*w                              models/base.py:182
    float fp;    // autograd d f/dx = (g1-g0)*(L-1) inside the clamp, else 0
    float w;     // fractional LUT position                      models/base.py:171
    int x0;      // lower LUT index                              models/base.py:169
};

__device__ __forceinline__ IcrfTap icrf_linear(float x, const float2 *__restrict__ row, float lm1) {
    IcrfTap t;
    const float xs_raw = __fmul_rn(x, lm1);                       // image * (L - 1)
    const float xs = fminf(fmaxf(xs_raw, 0.0f), lm1);             // .clamp_(0, L - 1)
    const float fl = floorf(xs);
    t.x0 = static_cast<int>(fl);
    t.w = __fsub_rn(xs, fl);
    const float2 g = row[t.x0];
    t.f = __fadd_rn(__fmul_rn(g.x, __fsub_rn(1.0f, t.w)), __fmul_rn(g.y, t.w));
    const bool inside = (xs_raw >= 0.0f) && (xs_raw <= lm1);      // clamp backward: closed interval
    t.fp = inside ? __fmul_rn(__fsub_rn(g.y, g.x), lm1) : 0.0f;
    return t;
}

// LOOKUP mode index: round-half-even, then clamp (models/base.py:145)
__device__ __forceinline__ int icrf_lookup_index(float x, float lm1) {
    const float r = rintf(__fmul_rn(x, lm1));
    return static_cast<int>(fminf(fmaxf(r, 0.0f), lm1));
}

// exp(-scale * (x - 0.5)^2) in the reference's op order (training/losses.py:205)
__device__ __forceinline__ float gaussian_weight(float x, float neg_scale, float &d) {
    d = __fsub_rn(x, 0.5f);
    return expf(__fmul_rn(neg_scale, __fmul_rn(d, d)));
}

__device__ __forceinline__ int wrap_inc(int u, int C) { return (u + 1 == C) ? 0 : u + 1; }

}  // namespace clair
