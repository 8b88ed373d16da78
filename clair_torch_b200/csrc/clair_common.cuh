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

// shared-memory packs (the pointers are VEC*4-byte aligned by construction)
template <int VEC>
__device__ __forceinline__ void store_shared(float *p, const Pack<VEC> &r) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(r.v[0], r.v[1]);
    } else {
        *p = r.v[0];
    }
}

template <int VEC>
__device__ __forceinline__ Pack<VEC> load_shared(const float *p) {
    Pack<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = *p;
    }
    return r;
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

__device__ __forceinline__ float exp2f_approx(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

__device__ __forceinline__ float rcp_approx(float t) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
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

// smem layout for the fused merge kernels: tab[u * L + k] = (theta[u][k], theta[u][min(k+1, L-1)] - theta[u][k])
__device__ __forceinline__ void stage_curve_slopes(float2 *tab, const float *__restrict__ theta, int n_rows, int L) {
    const int total = n_rows * L;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = i % L;
        const float a = __ldg(theta + i);
        const float b = (k + 1 < L) ? __ldg(theta + i + 1) : a;
        tab[i] = make_float2(a, __fsub_rn(b, a));
    }
}

// ---- ICRF evaluation, bit-exact with the reference's fp32 op sequence (no FMA contraction) ----
struct IcrfTap {
    float f;     // g0*(1-w) + g1*w                              models/base.py:182
    float fp;    // autograd d f/dx = (g1-g0)*(L-1) inside the clamp, else 0
    float w;     // fractional LUT position                      models/base.py:171
    int x0;      // lower LUT index                              models/base.py:169
};

// floor of a value in [0, 2^22] without the conversion unit: a round-down add of 2^23 leaves floor(xs) in the low
// mantissa bits (F2I / I2F / FRND all issue on the 16-lane XU pipe, which this path is short of).
__device__ __forceinline__ float floor_small(float xs, int &as_int) {
    const float t = __fadd_rd(xs, 8388608.0f);
    as_int = __float_as_int(t) & 0x007fffff;
    return __fsub_rn(t, 8388608.0f);
}

__device__ __forceinline__ IcrfTap icrf_linear(float x, const float2 *__restrict__ row, float lm1) {
    IcrfTap t;
    const float xs_raw = __fmul_rn(x, lm1);                       // image * (L - 1)
    const float xs = fminf(fmaxf(xs_raw, 0.0f), lm1);             // .clamp_(0, L - 1)
    const float fl = floor_small(xs, t.x0);
    t.w = __fsub_rn(xs, fl);
    const float2 g = row[t.x0];
    t.f = __fadd_rn(__fmul_rn(g.x, __fsub_rn(1.0f, t.w)), __fmul_rn(g.y, t.w));
    // clamp backward passes the gradient on the closed interval, i.e. exactly when the clamp changed nothing
    t.fp = (xs == xs_raw) ? __fmul_rn(__fsub_rn(g.y, g.x), lm1) : 0.0f;
    return t;
}

// Same evaluation with the table row given as a pre-biased 32-bit shared-memory address:
//   row_bias = shared_addr(row) - 8 * 0x4B000000   (mod 2^32)
// The round-down add of 2^23 leaves the bit pattern 0x4B000000 | x0, so `bits * 8 + row_bias` is the byte address
// of the (g0, g1) pair: one shift-add instead of mask + multiply-add + shift-add per element.
__device__ __forceinline__ uint32_t curve_row_bias(const float2 *row) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(row)) - 8u * 0x4B000000u;
}

__device__ __forceinline__ void icrf_linear_biased(float x, uint32_t row_bias, float lm1, float &f, float &fp) {
    const float xs_raw = __fmul_rn(x, lm1);
    const float xs = fminf(fmaxf(xs_raw, 0.0f), lm1);
    const float t = __fadd_rd(xs, 8388608.0f);
    const float w = __fsub_rn(xs, __fsub_rn(t, 8388608.0f));
    float g0, g1;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(g0), "=f"(g1) : "r"(static_cast<uint32_t>(__float_as_int(t)) * 8u + row_bias));
    f = __fadd_rn(__fmul_rn(g0, __fsub_rn(1.0f, w)), __fmul_rn(g1, w));
    fp = (xs == xs_raw) ? __fmul_rn(__fsub_rn(g1, g0), lm1) : 0.0f;
}

// The merge kernels' form over a (g0, g1 - g0) table: f = g0 + w (g1 - g0) in one FMA (within 1 ulp of the reference's
// g0 (1 - w) + g1 w; radiance is gated at 1e-5, only the model's own forward is held to the reference's bits) and
// f' = (g1 - g0)(L - 1), the same bits as above.  4 instructions fewer per element than icrf_linear_biased.
__device__ __forceinline__ void icrf_linear_slope(float x, uint32_t row_bias, float lm1, float &f, float &fp) {
    const float xs_raw = __fmul_rn(x, lm1);
    const float xs = fminf(fmaxf(xs_raw, 0.0f), lm1);
    const float t = __fadd_rd(xs, 8388608.0f);
    const float w = __fsub_rn(xs, __fsub_rn(t, 8388608.0f));
    float g0, dg;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(g0), "=f"(dg) : "r"(static_cast<uint32_t>(__float_as_int(t)) * 8u + row_bias));
    f = fmaf(w, dg, g0);
    fp = (xs == xs_raw) ? __fmul_rn(dg, lm1) : 0.0f;
}

__device__ __forceinline__ float sqrt_approx(float t) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

// LOOKUP mode index: round-half-even, then clamp (models/base.py:145)
__device__ __forceinline__ int icrf_lookup_index(float x, float lm1) {
    const float r = rintf(__fmul_rn(x, lm1));
    return static_cast<int>(fminf(fmaxf(r, 0.0f), lm1));
}

// exp(-scale * (x - 0.5)^2): d^2 rounded as the reference does (training/losses.py:205), then ONE multiply by
// -scale*log2(e) and MUFU.EX2.  Relative error <= ~5e-7 at the largest exponent used (|-30*0.25*log2e| = 10.8),
// far inside the 1e-5 gate; libdevice expf costs 8 more instructions per element.
__device__ __forceinline__ float gaussian_weight(float x, float neg_scale_log2e, float &d) {
    d = __fsub_rn(x, 0.5f);
    return exp2f_approx(__fmul_rn(neg_scale_log2e, __fmul_rn(d, d)));
}

// CATMULL mode (models/base.py:184-226): four taps x0-1 .. x0+2 (clamped), Catmull-Rom weights of t = xs - x0 in the
// reference's left-to-right fp32 op order (so the value is bit-exact), rows per the same k-mod-C rule as LINEAR.
// The derivative is the closed form sum_i w_i'(t) g_i (L-1) (the reference's autograd result differs from it by its own
// fp32 rounding, ~3e-5 of the maximum).
struct CatmullTaps {
    int idx[4];
    float w[4];
    float dw[4];
};

// taps from the clamped scaled position xs = clamp(x (L-1), 0, L-1); `inside` = the clamp changed nothing (its backward
// passes the gradient exactly then)
__device__ __forceinline__ CatmullTaps catmull_taps_xs(float xs, bool inside_clamp, int L) {
    CatmullTaps c;
    const float lm1 = static_cast<float>(L - 1);
    int x0;
    const float fl = floor_small(xs, x0);
#pragma unroll
    for (int k = 0; k < 4; ++k) c.idx[k] = min(max(x0 + k - 1, 0), L - 1);
    const float t = fminf(fmaxf(__fsub_rn(xs, fl), 0.0f), 1.0f);
    const float t2 = __fmul_rn(t, t), t3 = __fmul_rn(t2, t);
    c.w[0] = __fsub_rn(__fadd_rn(__fmul_rn(-0.5f, t3), t2), __fmul_rn(0.5f, t));
    c.w[1] = __fadd_rn(__fsub_rn(__fmul_rn(1.5f, t3), __fmul_rn(2.5f, t2)), 1.0f);
    c.w[2] = __fadd_rn(__fadd_rn(__fmul_rn(-1.5f, t3), __fmul_rn(2.0f, t2)), __fmul_rn(0.5f, t));
    c.w[3] = __fsub_rn(__fmul_rn(0.5f, t3), __fmul_rn(0.5f, t2));
    const float inside = inside_clamp ? lm1 : 0.0f;
    c.dw[0] = inside * (-1.5f * t2 + 2.0f * t - 0.5f);
    c.dw[1] = inside * (4.5f * t2 - 5.0f * t);
    c.dw[2] = inside * (-4.5f * t2 + 4.0f * t + 0.5f);
    c.dw[3] = inside * (1.5f * t2 - t);
    return c;
}

__device__ __forceinline__ CatmullTaps catmull_taps(float x, int L) {
    const float lm1 = static_cast<float>(L - 1);
    const float xs_raw = __fmul_rn(x, lm1);
    const float xs = fminf(fmaxf(xs_raw, 0.0f), lm1);
    return catmull_taps_xs(xs, xs == xs_raw, L);
}

// f and autograd's df/dx of a LOOKUP (models/base.py:138-158: nearest sample, no image edge) or CATMULL model from a
// table row whose .x entries are theta[u][k] (either staging layout above)
template <int MODE>
__device__ __forceinline__ void icrf_mode_eval(float x, const float2 *row, int L, float lm1, float &f, float &fp) {
    if constexpr (MODE == CLAIR_INTERP_LOOKUP) {
        f = row[icrf_lookup_index(x, lm1)].x;
        fp = 0.0f;
    } else {
        const CatmullTaps t = catmull_taps(x, L);
        const float g0 = row[t.idx[0]].x, g1 = row[t.idx[1]].x, g2 = row[t.idx[2]].x, g3 = row[t.idx[3]].x;
        f = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w[0], g0), __fmul_rn(t.w[1], g1)), __fmul_rn(t.w[2], g2)),
                      __fmul_rn(t.w[3], g3));
        fp = t.dw[0] * g0 + t.dw[1] * g1 + t.dw[2] * g2 + t.dw[3] * g3;
    }
}

// runtime-mode form of the same, for kernels that are not specialised per mode
__device__ __forceinline__ void icrf_mode_eval_rt(int mode, float x, const float2 *row, int L, float lm1, float &f, float &fp) {
    if (mode == CLAIR_INTERP_LOOKUP) icrf_mode_eval<CLAIR_INTERP_LOOKUP>(x, row, L, lm1, f, fp);
    else icrf_mode_eval<CLAIR_INTERP_CATMULL>(x, row, L, lm1, f, fp);
}

__device__ __forceinline__ int wrap_inc(int u, int C) { return (u + 1 == C) ? 0 : u + 1; }

// ---- packed fp32x2 arithmetic: sm_100a issues FFMA2 / FMUL2 / FADD2, two IEEE fp32 results per issue slot ----
// The pair kernels are bound by instruction issue, not by HBM, and give every lane two adjacent pixels, so their
// add / mul / fma chains run on 64-bit register pairs.  min / max / compare / MUFU have no packed form.
// ptxas contracts a packed mul feeding a packed add into FFMA2 even with .rn: nothing that selects an integer
// (LUT index, mask) may be written as such a pair.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 splat2(float v) { return pack2(v, v); }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2_rd(f32x2 a, f32x2 b) {      // round towards -inf: the floor trick of floor_small
    f32x2 d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 lds2(const float *p) { return *reinterpret_cast<const f32x2 *>(p); }
__device__ __forceinline__ void sts2(float *p, f32x2 v) { *reinterpret_cast<f32x2 *>(p) = v; }

// ---- bulk-copy staging (sm_90+ TMA without a tensor map: 1-D, 16-byte granular) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


}  // namespace clair
