// The fused HDR-merge kernels (templates) and their launch descriptor, shared by the translation units that instantiate
// them: clair_stack.cu (fp32 stacks, dark-field variant, all-modes kernel) and clair_merge_codes*.cu (integer ingest, one
// file per code layout so that they compile in parallel).  See DESIGN.md §3.1.
#pragma once

#include "clair_common.cuh"
#include "clair_dark.cuh"
#include "clair_host.h"

#include <cstdio>

namespace clair {

constexpr int kBlock = 256;
constexpr int kFrameChunk = 4;          // frames whose loads are in flight together (8 x LDG.128 / thread)
constexpr float kHdrNegScaleLog2e = -43.28085122666890f;  // -30 * log2(e): training/losses.py:193 default scale 30 (hdr_merge.py:95)

// =====================================================================================================
// HDR merge + uncertainty
// =====================================================================================================
struct HdrParams {
    const void *val;          // fp32 values, or uint8 / uint16 codes (integer ingest)
    const float *std;
    const float *theta;       // nullptr = identity
    double *mean_state;
    float *wsum_state;
    float *var_state;
    void *radiance;
    float *sigma;
    int64_t plane;            // pixels of this call (a whole plane, or a band of rows of it)
    int64_t stride;           // elements between channel planes in every buffer (= plane unless the call is a row band)
    int n_frames;
    int n_channels;
    int lut;
    int gaussian;
    int is_first;
    int is_final;
    int radiance_f64;
    int std_mode;             // kStdNone / kStdTensor / kStdMultiplier / kStdConstant
    float std_value;          // multiplier or constant
    float code_max;           // integer ingest: x = fl32(code) / fl32(code_max)   (CastTo + Normalize)
    float code_rcp;           // fl32(1 / code_max) when normalise_code16's short division is exact for all 65536 codes, else 0
    int src;                  // kSrcF32 / kSrcU8 / kSrcU16 (read by the all-modes kernel only; the others take it as a template)
    int hwc;                  // integer codes are (n_frames, H, W, 3) BGR-interleaved instead of planar
    int prefetch;             // camera-layout register kernels: 1 / 2 = prefetch the next trip's codes into L1 / L2, 0 = off
    const float *dark;        // fused dark-field mix (hdr_merge_dark_kernel): dark frames and their std, shaped like val
    int dark_rows;            // hdr_merge_dark_strip_kernel: rows per band of a warp's strip walk
    const float *dark_std;
    DarkGeometry dg;
    CurveRows rows;
    FrameScale scale;
};

constexpr int kStdNone = 0, kStdTensor = 1, kStdMultiplier = 2, kStdConstant = 3;
constexpr int kSrcF32 = 0, kSrcU8 = 1, kSrcU16 = 2, kSrcU8Hwc = 3, kSrcU16Hwc = 4;   // Hwc: interleaved BGR camera layout
// ...Tma: the same camera layout, a block's contiguous codes of every frame brought into shared memory by bulk copies
// (cp.async.bulk + mbarrier, one trip ahead) instead of per-thread strided loads
constexpr int kSrcU8HwcTma = 5, kSrcU16HwcTma = 6;
__host__ __device__ constexpr bool src_is_u8(int src) { return src == kSrcU8 || src == kSrcU8Hwc || src == kSrcU8HwcTma; }
__host__ __device__ constexpr bool src_is_tma(int src) { return src == kSrcU8HwcTma || src == kSrcU16HwcTma; }
__host__ __device__ constexpr bool src_is_hwc(int src) { return src == kSrcU8Hwc || src == kSrcU16Hwc || src_is_tma(src); }

// One frame's VEC pixel values.  SRC = kSrcF32: the fp32 stack the reference hands over.  kSrcU8 / kSrcU16: the raw
// integer codes, normalised in-register exactly like the reference's CPU transforms do it (CastTo(float32) then
// Normalize(max_val, min_val=0): an IEEE fp32 division, clair_torch/common/general_functions.py:378) — 8-bit codes
// go through a 256-entry table of those quotients, 16-bit codes through normalise_code16 (__fdiv_rn, or its exact
// reciprocal form in the register kernels).
// Interleaved camera layout (`hwc`): the codes are (n_frames, H, W, 3) in OpenCV's BGR order (p.stride pixels per frame) and channel c of the planar
// tensor is byte / halfword 2 - c of each pixel — the CvToTorch transform (common/general_functions.py:315-336) folded
// into the address.  Four pixels of one channel are 4 of the 12 codes a thread loads (two pixels: 2 of 6); the register kernels
// merge the three channels in turn in one thread, so the codes leave DRAM once (FOLD, hdr_merge_fixed_kernel).
__device__ __forceinline__ uint32_t pick_byte(uint32_t w0, uint32_t w1, uint32_t w2, int j) {      // byte j of a 12-byte window
    const uint32_t w = j < 4 ? w0 : (j < 8 ? w1 : w2);
    return (w >> (8 * (j & 3))) & 0xffu;
}

__device__ __forceinline__ uint32_t pick_half(const uint2 &a, const uint2 &b, const uint2 &c, int j) {   // halfword j of 24 bytes
    const uint32_t w = j < 2 ? a.x : j < 4 ? a.y : j < 6 ? b.x : j < 8 ? b.y : j < 10 ? c.x : c.y;
    return (j & 1) ? (w >> 16) : (w & 0xffffu);
}

// fl32(code) / fl32(code_max) for a 16-bit code, bit-identical to the IEEE division of the reference's Normalize.
// The code becomes a float without the conversion unit (2^23 + code is exact, subtracting 2^23 leaves fl32(code)).
// SHORT: the quotient is q0 = a * rcp corrected once by its exact residual, q = fma(fma(-q0, d, a), rcp, q0) — the
// correctly rounded a / d when rcp is the correctly rounded 1 / d (Markstein).  The host only selects a SHORT kernel (the
// register kernels) after checking that against a / d for every one of the 65536 codes (exact_code_reciprocal,
// clair_stack.cu; p.code_rcp != 0).  Against __fdiv_rn this drops MUFU.RCP, I2F and the division's range-check branch:
// two of the three XU-pipe operations per element.
template <bool SHORT>
__device__ __forceinline__ float normalise_code16(uint32_t code, const HdrParams &p) {
    const float a = __fsub_rn(__int_as_float(0x4B000000u | code), 8388608.0f);
    if constexpr (SHORT) {
        const float q0 = __fmul_rn(a, p.code_rcp);
        return fmaf(fmaf(-q0, p.code_max, a), p.code_rcp, q0);
    } else {
        return __fdiv_rn(a, p.code_max);
    }
}

template <int SRC, int VEC, bool SHORT = false>
__device__ __forceinline__ Pack<VEC> load_pixels(const HdrParams &p, int n, int c, uint32_t pix, int64_t o, const float *s_x,
                                                 const unsigned char *stage = nullptr, uint32_t stage_frame_bytes = 0) {
    if constexpr (SRC == kSrcF32) {
        return load_stream<VEC>(static_cast<const float *>(p.val) + o);
    } else if constexpr (src_is_u8(SRC)) {
        static_assert(VEC == 4 || VEC == 2, "integer ingest is 4 pixels per thread (2 in the 9..16-frame register kernel)");
        Pack<VEC> r;
        if constexpr (VEC == 2) {
            // two pixels: one 16-bit load, or the 6 bytes of two BGR pixels as three (the pair starts 2-byte aligned)
            if constexpr (SRC == kSrcU8HwcTma) {
                // staged frame: the channel's two bytes straight from the thread's 6-byte window (byte loads, no selects)
                const uint8_t *src = stage + n * stage_frame_bytes + threadIdx.x * 6 + (2 - c);
                r.v[0] = s_x[src[0]];
                r.v[1] = s_x[src[3]];
                return r;
            }
            if constexpr (SRC == kSrcU8Hwc) {
                uint32_t h0, h1, h2;
                {
                    const uint16_t *src = reinterpret_cast<const uint16_t *>(static_cast<const uint8_t *>(p.val) +
                                                                             (static_cast<int64_t>(n) * p.stride + pix) * 3);
                    h0 = __ldca(src); h1 = __ldca(src + 1); h2 = __ldca(src + 2);
                }
                const uint32_t w0 = h0 | (h1 << 16);          // bytes 0..3, h2 = bytes 4..5
                const int j0 = 2 - c;
                const uint32_t c0 = (w0 >> (8 * j0)) & 0xffu;                                   // byte j0
                const uint32_t c1 = j0 == 0 ? (w0 >> 24) : ((h2 >> (8 * (j0 - 1))) & 0xffu);    // byte 3 + j0
                r.v[0] = s_x[c0];
                r.v[1] = s_x[c1];
            } else {
                const uint32_t w = __ldcs(reinterpret_cast<const uint16_t *>(static_cast<const uint8_t *>(p.val) + o));
                r.v[0] = s_x[w & 0xffu];
                r.v[1] = s_x[w >> 8];
            }
            return r;
        }
        if constexpr (SRC == kSrcU8Hwc) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(static_cast<const uint8_t *>(p.val) +
                                                                     (static_cast<int64_t>(n) * p.stride + pix) * 3);
            const uint32_t w0 = __ldca(src), w1 = __ldca(src + 1), w2 = __ldca(src + 2);
            const int j0 = 2 - c;                              // uniform over the block
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t code;
                if (j0 == 0) code = pick_byte(w0, w1, w2, 3 * k);
                else if (j0 == 1) code = pick_byte(w0, w1, w2, 3 * k + 1);
                else code = pick_byte(w0, w1, w2, 3 * k + 2);
                r.v[k] = s_x[code];
            }
            return r;
        }
        const uint32_t w = __ldcs(reinterpret_cast<const uint32_t *>(static_cast<const uint8_t *>(p.val) + o));
#pragma unroll
        for (int k = 0; k < 4; ++k) r.v[k] = s_x[(w >> (8 * k)) & 0xffu];
        return r;
    } else {
        static_assert(VEC == 4 || VEC == 2, "integer ingest is 4 pixels per thread (2 in the 9..16-frame register kernel)");
        Pack<VEC> r;
        if constexpr (VEC == 2) {
            // two pixels: one 32-bit load, or the 12 bytes of two BGR pixels as three
            if constexpr (SRC == kSrcU16HwcTma) {
                // staged frame: the channel's two halfwords straight from the thread's 12-byte window (a lane stride of 3
                // banks: conflict-free), no word loads + selects
                const uint16_t *src = reinterpret_cast<const uint16_t *>(stage + n * stage_frame_bytes) + threadIdx.x * 6 + (2 - c);
                r.v[0] = normalise_code16<SHORT>(src[0], p);
                r.v[1] = normalise_code16<SHORT>(src[3], p);
                return r;
            }
            if constexpr (SRC == kSrcU16Hwc) {
                uint32_t w0, w1, w2;                                                            // halfwords 0..5
                {
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(static_cast<const uint16_t *>(p.val) +
                                                                             (static_cast<int64_t>(n) * p.stride + pix) * 3);
                    w0 = __ldca(src); w1 = __ldca(src + 1); w2 = __ldca(src + 2);
                }
                const int j0 = 2 - c;
                const uint32_t c0 = j0 == 0 ? (w0 & 0xffffu) : (j0 == 1 ? (w0 >> 16) : (w1 & 0xffffu));   // halfword j0
                const uint32_t c1 = j0 == 0 ? (w1 >> 16) : (j0 == 1 ? (w2 & 0xffffu) : (w2 >> 16));       // halfword 3 + j0
                r.v[0] = normalise_code16<SHORT>(c0, p);
                r.v[1] = normalise_code16<SHORT>(c1, p);
            } else {
                const uint32_t w = __ldcs(reinterpret_cast<const uint32_t *>(static_cast<const uint16_t *>(p.val) + o));
                r.v[0] = normalise_code16<SHORT>(w & 0xffffu, p);
                r.v[1] = normalise_code16<SHORT>(w >> 16, p);
            }
            return r;
        }
        if constexpr (SRC == kSrcU16Hwc) {
            const uint2 *src = reinterpret_cast<const uint2 *>(static_cast<const uint16_t *>(p.val) +
                                                               (static_cast<int64_t>(n) * p.stride + pix) * 3);
            const uint2 a = __ldca(src), b = __ldca(src + 1), cc = __ldca(src + 2);
            const int j0 = 2 - c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t code;
                if (j0 == 0) code = pick_half(a, b, cc, 3 * k);
                else if (j0 == 1) code = pick_half(a, b, cc, 3 * k + 1);
                else code = pick_half(a, b, cc, 3 * k + 2);
                r.v[k] = normalise_code16<SHORT>(code, p);
            }
            return r;
        }
        const uint2 w = __ldcs(reinterpret_cast<const uint2 *>(static_cast<const uint16_t *>(p.val) + o));
        r.v[0] = normalise_code16<SHORT>(w.x & 0xffffu, p);
        r.v[1] = normalise_code16<SHORT>(w.x >> 16, p);
        r.v[2] = normalise_code16<SHORT>(w.y & 0xffffu, p);
        r.v[3] = normalise_code16<SHORT>(w.y >> 16, p);
        return r;
    }
}

// The std of one frame's VEC pixels: a tensor, or synthesised the way MultiFileMapDataset does for missing std
// images (clair_torch/datasets/base.py:128-133): value * multiplier, or a constant.
// STD: 1 = tensor, 2 = synthesised (compile-time, so the tensor loads of the fp32 path stay unconditional and are
// hoisted with the value loads)
template <int VEC, int STD>
__device__ __forceinline__ Pack<VEC> load_std(const HdrParams &p, int64_t o, const Pack<VEC> &x) {
    if constexpr (STD == 1) {
        return load_stream<VEC>(p.std + o);
    } else {
        Pack<VEC> r;
#pragma unroll
        for (int k = 0; k < VEC; ++k) r.v[k] = (p.std_mode == kStdMultiplier) ? __fmul_rn(x.v[k], p.std_value) : p.std_value;
        return r;
    }
}

// Per frame element (all fp32):
//   w = exp(-30 (x-.5)^2) | 1,   q = w'/w = -60 (x-.5) | 0,   v = f(x)/t
//   R = s w (f'(x)/t + q v),     Q = s w q
// so that  s * d mean_new/dx = alpha R + gamma Q  with the per-pixel constants
//   alpha = (W_B/W) / (W_B + 1e-6),   gamma = (W_A/W^2)(mean_B - mean_A) - alpha mean_B
// (closed form of the autograd pass of inference/hdr_merge.py:107-115).
struct HdrTerms {
    float w, wv, R, Q;
    float P, v;      // P = s w f'/t (R without the weight-derivative part) and v: only the one-frame case reads them
};

__device__ __forceinline__ HdrTerms hdr_terms(float x, float s, float it, bool has_model, bool gaussian,
                                              uint32_t row_bias, float lm1, bool has_std) {
    HdrTerms o;
    float f = x, fp = 1.0f;
    if (has_model) icrf_linear_slope(x, row_bias, lm1, f, fp);
    float w = 1.0f, q = 0.0f;
    if (gaussian) {
        float d;
        w = gaussian_weight(x, kHdrNegScaleLog2e, d);
        q = __fmul_rn(-60.0f, d);
    }
    // every rounding below is explicit (no implicit contraction): hdr_terms2 is the same sequence on fp32x2 pairs and the
    // two must agree to the bit (callers accumulate sum w v as fmaf(w, v, sum))
    const float v = __fmul_rn(f, it);
    o.w = w;
    o.wv = __fmul_rn(w, v);
    if (has_std) {
        const float ws = __fmul_rn(w, s);
        const float t1 = __fmul_rn(fp, it);
        o.R = __fmul_rn(ws, fmaf(q, v, t1));
        o.Q = __fmul_rn(ws, q);
        o.P = __fmul_rn(ws, t1);
    } else {
        o.R = 0.0f; o.Q = 0.0f; o.P = 0.0f;
    }
    o.v = v;
    return o;
}

// hdr_terms for two adjacent pixels on packed fp32x2 arithmetic (FMUL2 / FADD2 / FFMA2: two results per issue slot;
// the merge kernel is HBM-bound in bursts but runs into the board's power cap in long runs, where fewer issued
// instructions buy clock).  Bit-identical to two hdr_terms calls.  P and v of the one-frame case are not produced.
struct HdrTerms2 {
    f32x2 w, v, R, Q;
};

__device__ __forceinline__ HdrTerms2 hdr_terms2(float x0, float x1, float s0, float s1, float it, bool has_model, bool gaussian,
                                                uint32_t bias0, uint32_t bias1, float lm1, bool has_std) {
    HdrTerms2 o;
    const f32x2 x2 = pack2(x0, x1);
    f32x2 f2 = x2, fp2 = splat2(1.0f);
    if (has_model) {
        float r0, r1;
        unpack2(mul2(x2, splat2(lm1)), r0, r1);                          // image * (L - 1)
        const float xs0 = fminf(fmaxf(r0, 0.0f), lm1), xs1 = fminf(fmaxf(r1, 0.0f), lm1);
        const f32x2 xs2 = pack2(xs0, xs1);
        const f32x2 two23 = splat2(8388608.0f);
        const f32x2 t2 = add2_rd(xs2, two23);
        float t0, t1, w0, w1;
        unpack2(t2, t0, t1);
        unpack2(sub2(xs2, sub2(t2, two23)), w0, w1);
        float g00, dg0, g01, dg1;
        asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(g00), "=f"(dg0) : "r"(static_cast<uint32_t>(__float_as_int(t0)) * 8u + bias0));
        asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(g01), "=f"(dg1) : "r"(static_cast<uint32_t>(__float_as_int(t1)) * 8u + bias1));
        f2 = pack2(fmaf(w0, dg0, g00), fmaf(w1, dg1, g01));
        fp2 = pack2((xs0 == r0) ? __fmul_rn(dg0, lm1) : 0.0f, (xs1 == r1) ? __fmul_rn(dg1, lm1) : 0.0f);
    }
    f32x2 w2 = splat2(1.0f), q2 = 0ull;
    if (gaussian) {
        const f32x2 d2 = add2(x2, splat2(-0.5f));
        float e0, e1;
        unpack2(mul2(splat2(kHdrNegScaleLog2e), mul2(d2, d2)), e0, e1);
        w2 = pack2(exp2f_approx(e0), exp2f_approx(e1));
        q2 = mul2(splat2(-60.0f), d2);
    }
    o.v = mul2(f2, splat2(it));
    o.w = w2;
    if (has_std) {
        const f32x2 ws = mul2(w2, pack2(s0, s1));
        const f32x2 t1 = mul2(fp2, splat2(it));
        o.R = mul2(ws, fma2(q2, o.v, t1));
        o.Q = mul2(ws, q2);
    } else {
        o.R = 0ull; o.Q = 0ull;
    }
    return o;
}

// A one-frame batch: mean_B = w v / (w + 1e-6), so v - mean_B = v * 1e-6 / (w + 1e-6) EXACTLY, while the generic
// R - mean_B Q has to recover that 1e-3..1e-6-sized difference from a cancellation (visible where f' = 0: a saturated
// pixel).  Returns s w [f'/t + q (v - mean_B)], the alpha-part of the gradient, cancellation-free.
__device__ __forceinline__ float one_frame_gradient(const HdrTerms &t) {
    const float wbe = t.w + 1e-6f;
    float inv = rcp_approx(wbe);
    inv = fmaf(fmaf(-wbe, inv, 1.0f), inv, inv);
    return fmaf(t.Q * t.v, 1e-6f * inv, t.P);
}

// Merge of the batch sums with the running state (common/statistics.py:88-109) and the output stage, shared by
// both kernels.  `var_update(k, alpha, gamma)` returns sum_n (alpha R_n + gamma Q_n)^2 for pixel k.
// SINGLE = the whole stack is this one batch (is_first && is_final): no state traffic, no float64.
// ONE = the batch is a single frame and the caller's R already holds the exact alpha-part of the gradient (see
// one_frame_gradient): gamma then carries the running-state term only.
template <int VEC, bool HAS_STD, bool SINGLE, bool ONE = false, typename VarFn>
__device__ __forceinline__ void hdr_finish(const HdrParams &p, int64_t off, const float (&wsum)[VEC], const float (&wv)[VEC],
                                           VarFn var_update) {
    if constexpr (SINGLE) {
        Pack<VEC> rad, sg;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float wbe = wsum[k] + 1e-6f;                    // statistics.py:76 (fp32 add)
            float inv = rcp_approx(wbe);
            inv = fmaf(fmaf(-wbe, inv, 1.0f), inv, inv);          // one Newton step: <= 1 ulp
            const float mean_b = wv[k] * inv;
            // W = 0.0 + W_B so W_B/W is exactly 1; an all-zero-weight pixel gives 0/0 = NaN as in the reference
            const float frac = (wsum[k] != 0.0f) ? 1.0f : __int_as_float(0x7fc00000);
            rad.v[k] = frac * mean_b;
            if constexpr (HAS_STD) {
                // sum_n (alpha R_n - alpha mean_B Q_n)^2 = alpha^2 sum_n (R_n - mean_B Q_n)^2: SINGLE callers' var_update
                // takes rho = -mean_B as its last argument and returns the unit sum (one FMA less per frame)
                const float alpha = frac * inv;
                sg.v[k] = alpha * sqrt_approx(var_update(k, 1.0f, ONE ? 0.0f : -mean_b));
            }
        }
        if (p.radiance_f64) {
            double r64[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) r64[k] = static_cast<double>(rad.v[k]);
            store_stream_f64<VEC>(static_cast<double *>(p.radiance) + off, r64);
        } else {
            store_stream<VEC>(static_cast<float *>(p.radiance) + off, rad);
        }
        if constexpr (HAS_STD) store_stream<VEC>(p.sigma + off, sg);
        return;
    }
    double mean_new[VEC];
    Pack<VEC> wtot, var_new;
    Pack<VEC> w_a, var_a;
    double mean_a[VEC];
    const bool first = p.is_first != 0;
    if (!first) {
        w_a = load_stream<VEC>(p.wsum_state + off);
        if constexpr (HAS_STD) var_a = load_stream<VEC>(p.var_state + off);
#pragma unroll
        for (int k = 0; k < VEC; ++k) mean_a[k] = __ldcs(p.mean_state + off + k);
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) { w_a.v[k] = 0.0f; var_a.v[k] = 0.0f; mean_a[k] = 0.0; }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float wbe = wsum[k] + 1e-6f;
        float inv = rcp_approx(wbe);
        inv = fmaf(fmaf(-wbe, inv, 1.0f), inv, inv);
        const float mean_b = wv[k] * inv;
        float alpha, gamma;
        if (first) {
            const float frac = (wsum[k] != 0.0f) ? 1.0f : __int_as_float(0x7fc00000);
            mean_new[k] = static_cast<double>(frac * mean_b);
            wtot.v[k] = wsum[k];
            alpha = frac * inv;
            gamma = ONE ? 0.0f : -alpha * mean_b;
        } else {
            const float wt = w_a.v[k] + wsum[k];                  // statistics.py:104
            const float frac = wsum[k] / wt;                      // statistics.py:106
            const double dm = static_cast<double>(mean_b) - mean_a[k];
            mean_new[k] = mean_a[k] + static_cast<double>(frac) * dm;
            wtot.v[k] = wt;
            alpha = frac * inv;
            gamma = static_cast<float>(static_cast<double>(w_a.v[k]) / (static_cast<double>(wt) * wt) * dm) - (ONE ? 0.0f : alpha * mean_b);
        }
        if constexpr (HAS_STD) var_new.v[k] = var_a.v[k] + var_update(k, alpha, gamma);
    }
    if (p.is_final) {
        if (p.radiance_f64) {
            store_stream_f64<VEC>(static_cast<double *>(p.radiance) + off, mean_new);
        } else {
            Pack<VEC> r;
#pragma unroll
            for (int k = 0; k < VEC; ++k) r.v[k] = static_cast<float>(mean_new[k]);
            store_stream<VEC>(static_cast<float *>(p.radiance) + off, r);
        }
        if constexpr (HAS_STD) {
            Pack<VEC> sg;
#pragma unroll
            for (int k = 0; k < VEC; ++k) sg.v[k] = sqrt_approx(var_new.v[k]);
            store_stream<VEC>(p.sigma + off, sg);
        }
    } else {
        store_stream_f64<VEC>(p.mean_state + off, mean_new);
        store_stream<VEC>(p.wsum_state + off, wtot);
        if constexpr (HAS_STD) store_stream<VEC>(p.var_state + off, var_new);
    }
}

// Table-row bookkeeping of a persistent thread: element (c, pix) reads row (pix + base(c)) mod C, so the VEC
// row addresses rotate by (stride mod C) per trip; one modulo per thread instead of one per trip.
template <int VEC>
struct RowCursor {
    uint32_t u0, du, C;
    __device__ __forceinline__ RowCursor(uint32_t first_pix, uint32_t pix_stride, uint32_t row_base, uint32_t channels)
        : u0((first_pix + row_base) % channels), du(pix_stride % channels), C(channels) {}
    __device__ __forceinline__ void advance() { u0 += du; u0 = (u0 >= C) ? u0 - C : u0; }
};

// ---- main kernel: N known at compile time (1..kMaxFixedFrames) -------------------------------------------
// Persistent blocks; one thread owns VEC adjacent pixels of one channel per loop trip.  All 2N vector loads of
// the trip are issued before the first use; (R_n, Q_n) stay in registers until mean_B is known, so the variance
// is a plain sum of squares of the actual (small) per-frame terms: no cancellation, everything in fp32.
constexpr int kMaxFixedFrames = 8;         // integer ingest and the dark-field variant
constexpr int kMaxFixedFramesF32 = 16;     // fp32 stacks, 2 pixels per thread: (R_n, Q_n) of 16 frames still fit the register file

// Camera layout: the channel loop costs ~25 registers (per-frame addresses stay live across it).  Where the single-batch kernel
// held three blocks per SM without it, it is held there (measured, c4 as uint16 camera codes: 1.45 ms at two blocks per SM,
// 1.18 ms at three; 11 frames 3.06 -> 1.44 ms); the multi-batch variants would spill under that cap and keep the default.
// 0 = no residency request (a request of 1 lifts ptxas' default register heuristics: the planar kernels grew from 64 to 102+
// registers and lost a quarter of their speed).
#ifndef CLAIR_FOLD_MIN_BLOCKS
#define CLAIR_FOLD_MIN_BLOCKS 3
#endif
__host__ __device__ constexpr int fixed_min_blocks(int src, int vec, int nf, bool single) {
    if (!src_is_hwc(src)) return 0;
    if (single && ((vec == 2 && nf <= 13) || (vec == 4 && nf <= 5))) return CLAIR_FOLD_MIN_BLOCKS;
    return ((vec == 2 && nf == 16) || (vec == 4 && nf == 8)) ? 2 : 0;       // these two would otherwise take 134 / 168 registers
}

template <int VEC, int NF, int STD, bool SINGLE, int SRC>
__global__ void __launch_bounds__(kBlock, fixed_min_blocks(SRC, VEC, NF, SINGLE)) hdr_merge_fixed_kernel(const HdrParams p) {
    constexpr bool HAS_STD = STD != 0;
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    float *s_x = reinterpret_cast<float *>(s_tab + (has_model ? C * L : 0));     // kSrcU8: code -> fl32(code)/code_max
    if (has_model) stage_curve_slopes(s_tab, p.theta, C, L);
    if constexpr (src_is_u8(SRC)) {
        for (int k = threadIdx.x; k < 256; k += blockDim.x) s_x[k] = __fdiv_rn(static_cast<float>(k), p.code_max);
    }
    // Camera layout, staged (TMA): per trip the block's 256 x VEC pixels are 3 * VEC * 256 contiguous codes of every frame.
    // One thread requests them a trip ahead with one bulk copy per frame into a two-stage shared-memory ring (an mbarrier
    // per stage counts the bytes in); the threads then read their 6 / 12 bytes per frame from shared memory, all three
    // channels from the same words, instead of waiting on 12-byte-strided global loads.
    constexpr bool TMA = src_is_tma(SRC);
    static_assert(!TMA || VEC == 2, "the staged camera kernels are the 2-pixel register kernels");
    constexpr uint32_t kStageFrameBytes = kBlock * VEC * 3 * (src_is_u8(SRC) ? 1 : 2);
    constexpr uint32_t kStageBytes = kStageFrameBytes * NF;
    unsigned char *s_stage = nullptr;
    uint64_t *s_bar = nullptr;
    if constexpr (TMA) {
        const uintptr_t after = reinterpret_cast<uintptr_t>(s_x + (src_is_u8(SRC) ? 256 : 0));
        s_stage = reinterpret_cast<unsigned char *>((after + 127) & ~static_cast<uintptr_t>(127));
        s_bar = reinterpret_cast<uint64_t *>(s_stage + 2 * kStageBytes);
        if (threadIdx.x == 0) {
            mbar_init(s_bar, 1);
            mbar_init(s_bar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    // Camera layout (FOLD): the three channels of a pixel share their bytes, so a thread merges them in turn — the first
    // channel's loads bring the codes into L1, the other two hit there (ld.global.ca).  One block per channel (gridDim.y)
    // read the stack from DRAM three times (ncu: 3.89 GB for the 1.30 GB of a c4 stack of uint16 codes).
    constexpr bool FOLD = src_is_hwc(SRC);
    const int c_first = FOLD ? 0 : static_cast<int>(blockIdx.y), c_last = FOLD ? C : c_first + 1;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.stride;
    const float lm1 = static_cast<float>(L - 1);
    const bool gaussian = p.gaussian != 0;
    const uint32_t n_items = static_cast<uint32_t>(p.plane / VEC);
    const uint32_t item_stride = gridDim.x * kBlock;
    const uint32_t first_item = blockIdx.x * kBlock + threadIdx.x;
    RowCursor<VEC> cur(first_item * VEC, item_stride * VEC, FOLD ? 0u : static_cast<uint32_t>(p.rows.base(c_first)), static_cast<uint32_t>(C));
    // FOLD: the cursor follows the pixel alone; channel c adds (row base of c) mod C  (3 channels, checked by the host)
    const uint32_t row_shift0 = FOLD ? static_cast<uint32_t>(p.rows.base(0)) % cur.C : 0u;
    const uint32_t row_shift1 = FOLD ? static_cast<uint32_t>(p.rows.base(1)) % cur.C : 0u;
    const uint32_t row_shift2 = FOLD ? static_cast<uint32_t>(p.rows.base(2)) % cur.C : 0u;
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;

    // TMA: bulk copies of one trip of this block (pixels [first, first + 256 VEC) of every frame) into stage `st`
    auto request_trip = [&](uint32_t block_first_item, uint32_t st) {
        constexpr uint32_t kPixBytes = 3 * (src_is_u8(SRC) ? 1 : 2);
        const uint32_t first_pix = block_first_item * VEC;
        const uint32_t n_pix = min(static_cast<uint32_t>(kBlock * VEC), static_cast<uint32_t>(p.plane) - first_pix);
        const uint32_t bytes = n_pix * kPixBytes;                    // a multiple of 16: the host checked H*W % 16 == 0
        mbar_expect_tx(s_bar + st, bytes * NF);
        const char *src = static_cast<const char *>(p.val) + static_cast<int64_t>(first_pix) * kPixBytes;
        const int64_t frame_bytes = p.stride * kPixBytes;
#pragma unroll
        for (int n = 0; n < NF; ++n) bulk_copy_g2s(s_stage + st * kStageBytes + n * kStageFrameBytes, src + n * frame_bytes, bytes, s_bar + st);
    };
    uint32_t trip = 0;
    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            const uint32_t b0 = blockIdx.x * kBlock;
            if (b0 < n_items) request_trip(b0, 0);
            if (b0 + item_stride < n_items) request_trip(b0 + item_stride, 1);
        }
    }
    // (TMA: the loop bound is the block's, so that every thread reaches the barrier that frees a stage)
    for (uint32_t item = first_item; TMA ? (item - threadIdx.x < n_items) : (item < n_items); item += item_stride, cur.advance(), ++trip) {
        const uint32_t pix = item * VEC;
        const unsigned char *stage = nullptr;
        if constexpr (TMA) {
            mbar_wait(s_bar + (trip & 1u), (trip >> 1) & 1u);
            stage = s_stage + (trip & 1u) * kStageBytes;
        }
        if (!TMA || item < n_items) {
        if constexpr (FOLD && !TMA) {
            // the codes of the next trip, requested while this trip's three channels are merged: their first use then finds
            // them on the chip instead of waiting on DRAM at two or three blocks per SM
            if (p.prefetch != 0 && item + item_stride < n_items) {
                constexpr int kCodeBytes = src_is_u8(SRC) ? 1 : 2;
                const char *next = static_cast<const char *>(p.val) + static_cast<int64_t>(pix + item_stride * VEC) * (3 * kCodeBytes);
                const int64_t frame_bytes = p.stride * (3 * kCodeBytes);
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    if (p.prefetch == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(next + n * frame_bytes));
                    else asm volatile("prefetch.global.L2 [%0];" ::"l"(next + n * frame_bytes));
                }
            }
        }
#pragma unroll 1
        for (int c = c_first; c < c_last; ++c) {
        const int64_t off = static_cast<int64_t>(c) * p.stride + pix;
        uint32_t bias[VEC];
        {
            uint32_t u = cur.u0;
            if constexpr (FOLD) {
                u += (c == 0) ? row_shift0 : ((c == 1) ? row_shift1 : row_shift2);
                u = (u >= cur.C) ? u - cur.C : u;
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                bias[k] = tab_bias + u * row_bytes;
                u = (u + 1 == cur.C) ? 0u : u + 1;
            }
        }
        float wsum[VEC], wv[VEC];
        if constexpr (VEC % 2 == 0 && NF > 1) {
            // pixel pairs on packed fp32x2 arithmetic; (R_n, Q_n) stay packed until the per-pixel epilogue
            constexpr int HV = VEC / 2;
            f32x2 wsum2[HV], wv2[HV], R2[NF][HV], Q2[NF][HV];
#pragma unroll
            for (int h = 0; h < HV; ++h) { wsum2[h] = 0ull; wv2[h] = 0ull; }
            {
                Pack<VEC> xv[NF], sv[NF];
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    const int64_t o = off + static_cast<int64_t>(n) * frame_stride;
                    xv[n] = load_pixels<SRC, VEC, true>(p, n, c, pix, o, s_x, stage, kStageFrameBytes);
                    if constexpr (HAS_STD) sv[n] = load_std<VEC, STD>(p, o, xv[n]);
                }
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    const float it = p.scale.inv_t[n];
#pragma unroll
                    for (int h = 0; h < HV; ++h) {
                        const HdrTerms2 t = hdr_terms2(xv[n].v[2 * h], xv[n].v[2 * h + 1], HAS_STD ? sv[n].v[2 * h] : 0.0f,
                                                       HAS_STD ? sv[n].v[2 * h + 1] : 0.0f, it, has_model, gaussian, bias[2 * h],
                                                       bias[2 * h + 1], lm1, HAS_STD);
                        wsum2[h] = add2(wsum2[h], t.w);
                        wv2[h] = fma2(t.w, t.v, wv2[h]);
                        R2[n][h] = t.R;
                        Q2[n][h] = t.Q;
                    }
                }
            }
            if constexpr (SINGLE) {
                // the single-batch epilogue of hdr_finish on pixel pairs (the same operations, so the same bits)
                Pack<VEC> rad, sg;
#pragma unroll
                for (int h = 0; h < HV; ++h) {
                    const f32x2 wbe = add2(wsum2[h], splat2(1e-6f));                 // statistics.py:76 (fp32 add)
                    float b0, b1, ws0, ws1;
                    unpack2(wbe, b0, b1);
                    unpack2(wsum2[h], ws0, ws1);
                    f32x2 inv = pack2(rcp_approx(b0), rcp_approx(b1));
                    inv = fma2(fma2(sub2(0ull, wbe), inv, splat2(1.0f)), inv, inv);  // one Newton step: <= 1 ulp
                    const f32x2 mean_b = mul2(wv2[h], inv);
                    const float nan = __int_as_float(0x7fc00000);                    // an all-zero-weight pixel: 0/0 as in the reference
                    const f32x2 frac = pack2(ws0 != 0.0f ? 1.0f : nan, ws1 != 0.0f ? 1.0f : nan);
                    unpack2(mul2(frac, mean_b), rad.v[2 * h], rad.v[2 * h + 1]);
                    if constexpr (HAS_STD) {
                        const f32x2 rho = sub2(0ull, mean_b);
                        f32x2 acc = 0ull;
#pragma unroll
                        for (int n = 0; n < NF; ++n) {
                            const f32x2 g = fma2(rho, Q2[n][h], R2[n][h]);
                            acc = fma2(g, g, acc);
                        }
                        float a0, a1;
                        unpack2(acc, a0, a1);
                        unpack2(mul2(mul2(frac, inv), pack2(sqrt_approx(a0), sqrt_approx(a1))), sg.v[2 * h], sg.v[2 * h + 1]);
                    }
                }
                if (p.radiance_f64) {
                    double r64[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) r64[k] = static_cast<double>(rad.v[k]);
                    store_stream_f64<VEC>(static_cast<double *>(p.radiance) + off, r64);
                } else {
                    store_stream<VEC>(static_cast<float *>(p.radiance) + off, rad);
                }
                if constexpr (HAS_STD) store_stream<VEC>(p.sigma + off, sg);
            } else {
#pragma unroll
            for (int h = 0; h < HV; ++h) {
                unpack2(wsum2[h], wsum[2 * h], wsum[2 * h + 1]);
                unpack2(wv2[h], wv[2 * h], wv[2 * h + 1]);
            }
            hdr_finish<VEC, HAS_STD, SINGLE, false>(p, off, wsum, wv, [&](int k, float alpha, float gamma) {
                float acc = 0.0f;
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    float r0, r1, q0, q1;
                    unpack2(R2[n][k >> 1], r0, r1);
                    unpack2(Q2[n][k >> 1], q0, q1);
                    const float r = (k & 1) ? r1 : r0, q = (k & 1) ? q1 : q0;
                    const float g = SINGLE ? fmaf(gamma, q, r) : fmaf(alpha, r, __fmul_rn(gamma, q));
                    acc = fmaf(g, g, acc);
                }
                return acc;
            });
            }
        } else {
        float R[NF][VEC], Q[NF][VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { wsum[k] = 0.0f; wv[k] = 0.0f; }
        {
            Pack<VEC> xv[NF], sv[NF];
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                const int64_t o = off + static_cast<int64_t>(n) * frame_stride;
                xv[n] = load_pixels<SRC, VEC, true>(p, n, c, pix, o, s_x);
                if constexpr (HAS_STD) sv[n] = load_std<VEC, STD>(p, o, xv[n]);
            }
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                const float it = p.scale.inv_t[n];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const HdrTerms t = hdr_terms(xv[n].v[k], HAS_STD ? sv[n].v[k] : 0.0f, it, has_model, gaussian, bias[k], lm1,
                                                 HAS_STD);
                    wsum[k] += t.w;
                    wv[k] = fmaf(t.w, t.v, wv[k]);
                    R[n][k] = (NF == 1) ? one_frame_gradient(t) : t.R;
                    Q[n][k] = t.Q;
                }
            }
        }
        hdr_finish<VEC, HAS_STD, SINGLE, NF == 1>(p, off, wsum, wv, [&](int k, float alpha, float gamma) {
            float acc = 0.0f;
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                const float g = SINGLE ? fmaf(gamma, Q[n][k], R[n][k]) : fmaf(alpha, R[n][k], __fmul_rn(gamma, Q[n][k]));
                acc = fmaf(g, g, acc);
            }
            return acc;
        });
        }
        }   // channel
        }   // item in range
        if constexpr (TMA) {
            __syncthreads();                                         // every thread has read this trip's stage
            const uint32_t ahead = item - threadIdx.x + 2u * item_stride;
            if (threadIdx.x == 0 && ahead < n_items) request_trip(ahead, trip & 1u);
        }
    }
}

// Output stage of the dark-field kernels for the pixel pair a thread owns (packed sums, (R_n, Q_n) of all frames in registers)
template <int NF, bool SINGLE>
__device__ __forceinline__ void dark_finish_pair(const HdrParams &p, int64_t off, f32x2 wsum2, f32x2 wv2, const f32x2 (&R2)[NF],
                                                 const f32x2 (&Q2)[NF]) {
    constexpr int VEC = 2;
    if constexpr (SINGLE) {
        // the single-batch epilogue of hdr_merge_fixed_kernel on the pixel pair (same operations, same bits)
        Pack<VEC> rad, sg;
        const f32x2 wbe = add2(wsum2, splat2(1e-6f));                     // statistics.py:76 (fp32 add)
        float b0, b1, ws0, ws1;
        unpack2(wbe, b0, b1);
        unpack2(wsum2, ws0, ws1);
        f32x2 inv = pack2(rcp_approx(b0), rcp_approx(b1));
        inv = fma2(fma2(sub2(0ull, wbe), inv, splat2(1.0f)), inv, inv);      // one Newton step: <= 1 ulp
        const f32x2 mean_b = mul2(wv2, inv);
        const float nan = __int_as_float(0x7fc00000);                        // an all-zero-weight pixel: 0/0 as in the reference
        const f32x2 frac = pack2(ws0 != 0.0f ? 1.0f : nan, ws1 != 0.0f ? 1.0f : nan);
        unpack2(mul2(frac, mean_b), rad.v[0], rad.v[1]);
        const f32x2 rho = sub2(0ull, mean_b);
        f32x2 acc = 0ull;
#pragma unroll
        for (int n = 0; n < NF; ++n) {
            const f32x2 g = fma2(rho, Q2[n], R2[n]);
            acc = fma2(g, g, acc);
        }
        float a0, a1;
        unpack2(acc, a0, a1);
        unpack2(mul2(mul2(frac, inv), pack2(sqrt_approx(a0), sqrt_approx(a1))), sg.v[0], sg.v[1]);
        if (p.radiance_f64) {
            double r64[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) r64[k] = static_cast<double>(rad.v[k]);
            store_stream_f64<VEC>(static_cast<double *>(p.radiance) + off, r64);
        } else {
            store_stream<VEC>(static_cast<float *>(p.radiance) + off, rad);
        }
        store_stream<VEC>(p.sigma + off, sg);
    } else {
        float wsum[VEC], wv[VEC];
        unpack2(wsum2, wsum[0], wsum[1]);
        unpack2(wv2, wv[0], wv[1]);
        hdr_finish<VEC, true, SINGLE, false>(p, off, wsum, wv, [&](int k, float alpha, float gamma) {
            float acc = 0.0f;
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                float r0, r1, q0, q1;
                unpack2(R2[n], r0, r1);
                unpack2(Q2[n], q0, q1);
                const float r = (k & 1) ? r1 : r0, q = (k & 1) ? q1 : q0;
                const float g = fmaf(alpha, r, __fmul_rn(gamma, q));
                acc = fmaf(g, g, acc);
            }
            return acc;
        });
    }
}

// ---- the fixed-N kernel with the dark-field mix fused into its load ----------------------
// x' = m B(x) + (1 - m) x and s_eff (clair_dark.cuh) are formed in registers from the raw frame, its std and the dark
// frame + std, so the mixed stack never exists in memory: 4 input stacks are read once instead of the pre-pass
// writing and the merge re-reading two more.  2 pixels of one row per thread; the loop bound is warp-uniform because
// the blur takes its left / right neighbours from the adjacent lanes.
constexpr int kDarkChunk = 3;      // frames whose 18 loads are in flight together (chunks of 1..3 measured within 3 %)

template <int NF, bool SINGLE>
__global__ void __launch_bounds__(kBlock, 2) hdr_merge_dark_kernel(const HdrParams p) {
    constexpr int VEC = 2;
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    if (has_model) stage_curve_slopes(s_tab, p.theta, C, L);
    __syncthreads();
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.stride;
    const float lm1 = static_cast<float>(L - 1);
    const bool gaussian = p.gaussian != 0;
    const uint32_t n_items = static_cast<uint32_t>(p.plane / VEC);
    const uint32_t groups_per_row = static_cast<uint32_t>(p.dg.W / VEC);
    const uint32_t item_stride = gridDim.x * kBlock;
    const uint32_t first_item = blockIdx.x * kBlock + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    RowCursor<VEC> cur(first_item * VEC, item_stride * VEC, static_cast<uint32_t>(p.rows.base(c)), static_cast<uint32_t>(C));
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const float *val = static_cast<const float *>(p.val);
    for (uint32_t item = first_item; item - lane < n_items; item += item_stride, cur.advance()) {
        const bool active = item < n_items;
        const uint32_t row = active ? item / groups_per_row : 0u;
        const uint32_t grp = item - row * groups_per_row;
        const int col = static_cast<int>(grp) * VEC;
        const bool chained_left = lane > 0 && grp > 0;
        const bool chained_right = lane < 31 && grp + 1 < groups_per_row;
        const int64_t off = static_cast<int64_t>(c) * p.stride + (active ? item * VEC : 0u);
        uint32_t bias[VEC];
        {
            uint32_t u = cur.u0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                bias[k] = tab_bias + u * row_bytes;
                u = (u + 1 == cur.C) ? 0u : u + 1;
            }
        }
        // The two pixels of the thread go through the merge arithmetic as one packed fp32x2 pair (hdr_terms2, as in
        // hdr_merge_fixed_kernel): the kernel is instruction-bound (blur + mix + merge), not HBM-bound.  A single frame keeps
        // the scalar form (its gradient is the exact one-frame expression).
        constexpr bool PACKED = NF > 1;
        float wsum[VEC], wv[VEC], R[PACKED ? 1 : NF][VEC], Q[PACKED ? 1 : NF][VEC];
        f32x2 wsum2 = 0ull, wv2 = 0ull, R2[PACKED ? NF : 1], Q2[PACKED ? NF : 1];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { wsum[k] = 0.0f; wv[k] = 0.0f; }
        // frames in chunks of kDarkChunk: all loads of a chunk are issued before its first shuffle
#pragma unroll
        for (int n0 = 0; n0 < NF; n0 += kDarkChunk) {
            RowGroup<VEC> rg[kDarkChunk];
            Pack<VEC> sv[kDarkChunk], dk[kDarkChunk], ds[kDarkChunk];
#pragma unroll
            for (int j = 0; j < kDarkChunk; ++j) {
                if (n0 + j < NF) {
                    const int64_t o = off + static_cast<int64_t>(n0 + j) * frame_stride;
                    row_group_load<VEC>(val + static_cast<int64_t>(n0 + j) * frame_stride + static_cast<int64_t>(c) * p.stride,
                                        static_cast<int>(row), col, p.dg, active, chained_left, chained_right, rg[j]);
                    sv[j] = load_stream<VEC>(p.std + o);
                    dk[j] = load_stream<VEC>(p.dark + o);
                    ds[j] = load_stream<VEC>(p.dark_std + o);
                }
            }
#pragma unroll
            for (int j = 0; j < kDarkChunk; ++j) {
                if (n0 + j < NF) {
                    float x[VEC], blur[VEC];
                    row_group_blur<VEC>(rg[j], col, p.dg, chained_left, chained_right, x, blur);
                    const float it = p.scale.inv_t[n0 + j];
                    float xm[VEC], sm[VEC];
                    {
                        f32x2 xm2, sm2;
                        dark_mix_value2<true>(pack2(x[0], x[1]), pack2(blur[0], blur[1]), pack2(sv[j].v[0], sv[j].v[1]),
                                              pack2(dk[j].v[0], dk[j].v[1]), pack2(ds[j].v[0], ds[j].v[1]), p.dg, xm2, sm2);
                        unpack2(xm2, xm[0], xm[1]);
                        unpack2(sm2, sm[0], sm[1]);
                    }
                    if constexpr (PACKED) {
                        const HdrTerms2 t = hdr_terms2(xm[0], xm[1], sm[0], sm[1], it, has_model, gaussian, bias[0], bias[1], lm1, true);
                        wsum2 = add2(wsum2, t.w);
                        wv2 = fma2(t.w, t.v, wv2);
                        R2[n0 + j] = t.R;
                        Q2[n0 + j] = t.Q;
                    } else {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) {
                            const HdrTerms t = hdr_terms(xm[k], sm[k], it, has_model, gaussian, bias[k], lm1, true);
                            wsum[k] += t.w;
                            wv[k] = fmaf(t.w, t.v, wv[k]);
                            R[n0 + j][k] = one_frame_gradient(t);
                            Q[n0 + j][k] = t.Q;
                        }
                    }
                }
            }
        }
        if (!active) continue;
        if constexpr (PACKED) {
            dark_finish_pair<NF, SINGLE>(p, off, wsum2, wv2, R2, Q2);
        } else {
            hdr_finish<VEC, true, SINGLE, true>(p, off, wsum, wv, [&](int k, float alpha, float gamma) {
                float acc = 0.0f;
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    const float g = SINGLE ? fmaf(gamma, Q[n][k], R[n][k]) : fmaf(alpha, R[n][k], gamma * Q[n][k]);
                    acc = fmaf(g, g, acc);
                }
                return acc;
            });
        }
    }
}

// ---- the same merge as a walk down column strips ----------------------------------------------------------------------
// The form above reads three value rows per output row (6 N vector loads + 6 N predicated end-lane loads per pixel pair)
// and alternates between waiting for a chunk of loads and computing on it: loads alone 100 us at c1 size, arithmetic alone
// ~95 us, 142 us together; prefetching, an asynchronous ring and more resident blocks all measured the same.  Here a warp
// owns a strip of 60 output pixels and walks down a band of rows with the previous two value rows of every frame in
// registers, so each value row is loaded once (4 N loads per pixel pair and row); lanes 0 and 31 of the warp hold the
// strip's halo pixel pairs and produce no output, so the neighbour column sums always come from the adjacent lanes and
// there are no end-lane loads or selects; and the loads of row r + 1 are issued before the arithmetic of row r.
// kStripPix = 60: 1920 = 32 strips exactly.  Measured (1080p RGB): 3 frames 85 -> 67 us, 4 frames 127 -> 88 us, 5 frames (c1) 143 ->
// 116 us = 0.72 of the roofline of its four input stacks; 5 x 4K 593 -> 433 us; 6 / 7 / 8 frames (no prefetch set) 159 -> 135,
// 194 -> 156, 211 -> 176 us.
constexpr int kStripPix = 60;
constexpr int kMaxStripFrames = 5;     // most frames whose NEXT row is prefetched into registers (128 registers at 5 frames)

// PIPE: next row's inputs in a second register set (2..5 frames); without it (6..8 frames) a row's own inputs are loaded at its
// start, all frames at once, and only the two-row value window persists
template <int NF, bool SINGLE, bool PIPE = (NF <= kMaxStripFrames)>
__global__ void __launch_bounds__(kBlock, 2) hdr_merge_dark_strip_kernel(const HdrParams p) {
    constexpr int VEC = 2;
    static_assert(NF > 1, "the strip walk is the packed multi-frame kernel");
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    if (has_model) stage_curve_slopes(s_tab, p.theta, C, L);
    __syncthreads();
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.stride;
    const float lm1 = static_cast<float>(L - 1);
    const bool gaussian = p.gaussian != 0;
    const int W = p.dg.W, H = p.dg.H, R = p.dark_rows;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t strips = static_cast<uint32_t>((W + kStripPix - 1) / kStripPix), bands = static_cast<uint32_t>((H + R - 1) / R);
    const uint32_t n_tasks = strips * bands;
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const uint32_t uC = static_cast<uint32_t>(C);
    const float *val_c = static_cast<const float *>(p.val) + static_cast<int64_t>(c) * p.stride;
    const float *std_c = p.std + static_cast<int64_t>(c) * p.stride;
    const float *dark_c = p.dark + static_cast<int64_t>(c) * p.stride;
    const float *dstd_c = p.dark_std + static_cast<int64_t>(c) * p.stride;
    constexpr float t0 = blur_tap(0), t1 = blur_tap(1);
    auto ld = [](const float *q, bool on) { return on ? __ldcs(reinterpret_cast<const unsigned long long *>(q)) : 0ull; };

    for (uint32_t task = blockIdx.x * (kBlock / 32) + warp; task < n_tasks; task += gridDim.x * (kBlock / 32)) {
        const uint32_t band = task / strips, strip = task - band * strips;
        const int col = static_cast<int>(strip) * kStripPix - 2 + 2 * static_cast<int>(lane);     // the lane's pixel pair (W is even)
        const bool in_row = col >= 0 && col < W;
        const bool owner = in_row && lane >= 1u && lane <= 30u;
        const bool wrap_l = col == 0, wrap_r = col + 2 == W;          // x[-1] = x[1], x[W] = x[W-2]
        const int r_first = static_cast<int>(band) * R, r_end = min(r_first + R, H);
        const uint32_t at0 = static_cast<uint32_t>(in_row ? col : 0);
        // the two value rows above the first output row (reflect padding) of every frame, and the first row's other inputs
        f32x2 above[NF], here[NF], below[NF], sv[NF], dk[NF], ds[NF];
        {
            const uint32_t ra = static_cast<uint32_t>(r_first == 0 ? 1 : r_first - 1) * W + at0, rh = static_cast<uint32_t>(r_first) * W + at0;
            const uint32_t rb = static_cast<uint32_t>(r_first == H - 1 ? H - 2 : r_first + 1) * W + at0;
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                const int64_t fo = static_cast<int64_t>(n) * frame_stride;
                above[n] = ld(val_c + fo + ra, in_row);
                here[n] = ld(val_c + fo + rh, in_row);
                if constexpr (PIPE) {
                    below[n] = ld(val_c + fo + rb, in_row);
                    sv[n] = ld(std_c + fo + rh, in_row);
                    dk[n] = ld(dark_c + fo + rh, in_row);
                    ds[n] = ld(dstd_c + fo + rh, in_row);
                }
            }
        }
        uint32_t u_row = (static_cast<uint32_t>(r_first) * W + at0 + static_cast<uint32_t>(p.rows.base(c))) % uC;   // table row of the pair's first pixel
        const uint32_t du_row = static_cast<uint32_t>(W) % uC;
        for (int r = r_first; r < r_end; ++r) {
            // next row's inputs go out before this row's arithmetic (nothing is loaded past the band's last row)
            const bool more = in_row && r + 1 < r_end;
            f32x2 nbelow[PIPE ? NF : 1], nsv[PIPE ? NF : 1], ndk[PIPE ? NF : 1], nds[PIPE ? NF : 1];
            if constexpr (PIPE) {
                const uint32_t rn = static_cast<uint32_t>(r + 1) * W + at0;
                const uint32_t rb = static_cast<uint32_t>(r + 2 >= H ? (r + 2 == H ? H - 2 : H - 1) : r + 2) * W + at0;
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    const int64_t fo = static_cast<int64_t>(n) * frame_stride;
                    nbelow[n] = ld(val_c + fo + rb, more);
                    nsv[n] = ld(std_c + fo + rn, more);
                    ndk[n] = ld(dark_c + fo + rn, more);
                    nds[n] = ld(dstd_c + fo + rn, more);
                }
            } else {
                const uint32_t rh = static_cast<uint32_t>(r) * W + at0;
                const uint32_t rb = static_cast<uint32_t>(r + 1 == H ? H - 2 : r + 1) * W + at0;
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    const int64_t fo = static_cast<int64_t>(n) * frame_stride;
                    below[n] = ld(val_c + fo + rb, in_row);
                    sv[n] = ld(std_c + fo + rh, in_row);
                    dk[n] = ld(dark_c + fo + rh, in_row);
                    ds[n] = ld(dstd_c + fo + rh, in_row);
                }
            }
            uint32_t bias[VEC];
            bias[0] = tab_bias + u_row * row_bytes;
            bias[1] = tab_bias + ((u_row + 1u == uC) ? 0u : u_row + 1u) * row_bytes;
            f32x2 wsum2 = 0ull, wv2 = 0ull, R2[NF], Q2[NF];
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                // separable blur (clair_dark.cuh): column sums of the lane's two pixels, the neighbours' from the adjacent lanes
                const f32x2 v2 = fma2(splat2(t0), add2(above[n], below[n]), mul2(splat2(t1), here[n]));
                float v_lo, v_hi;
                unpack2(v2, v_lo, v_hi);
                float left = __shfl_up_sync(0xffffffffu, v_hi, 1), right = __shfl_down_sync(0xffffffffu, v_lo, 1);
                left = wrap_l ? v_hi : left;
                right = wrap_r ? v_lo : right;
                const f32x2 blur2 = fma2(splat2(t0), add2(pack2(left, v_lo), pack2(v_hi, right)), mul2(splat2(t1), v2));
                f32x2 xm2, sm2;
                dark_mix_value2<true>(here[n], blur2, sv[n], dk[n], ds[n], p.dg, xm2, sm2);
                float xm[VEC], sm[VEC];
                unpack2(xm2, xm[0], xm[1]);
                unpack2(sm2, sm[0], sm[1]);
                const HdrTerms2 t = hdr_terms2(xm[0], xm[1], sm[0], sm[1], p.scale.inv_t[n], has_model, gaussian, bias[0], bias[1], lm1, true);
                wsum2 = add2(wsum2, t.w);
                wv2 = fma2(t.w, t.v, wv2);
                R2[n] = t.R;
                Q2[n] = t.Q;
            }
            if (owner)
                dark_finish_pair<NF, SINGLE>(p, static_cast<int64_t>(c) * p.stride + static_cast<int64_t>(r) * W + col, wsum2, wv2, R2, Q2);
#pragma unroll
            for (int n = 0; n < NF; ++n) {
                above[n] = here[n]; here[n] = below[n];
                if constexpr (PIPE) { below[n] = nbelow[n]; sv[n] = nsv[n]; dk[n] = ndk[n]; ds[n] = nds[n]; }
            }
            u_row += du_row;
            u_row = (u_row >= uC) ? u_row - uC : u_row;
        }
    }
}

// ---- 9 <= N <= ~40: the same two-step evaluation with (R_n, Q_n) parked in shared memory ---------------------
// N is a runtime value; frames are loaded in chunks of four.  Each thread owns a private, conflict-free column
// rq[(n*2 + which)*VEC + k][tid], so registers stay low (occupancy is bounded by shared memory instead: 4 KB x N per
// 256-thread block at 2 px/thread) and the variance is again a plain fp32 sum of squares.
template <int VEC, int STD, bool SINGLE, int SRC>
__global__ void __launch_bounds__(kBlock) hdr_merge_smem_kernel(const HdrParams p) {
    constexpr bool HAS_STD = STD != 0;
    static_assert(HAS_STD, "without std images there is no second step");
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    float *s_x = reinterpret_cast<float *>(s_tab + (has_model ? C * L : 0));
    float *s_rq = s_x + (src_is_u8(SRC) ? 256 : 0) + threadIdx.x;
    if (has_model) stage_curve_slopes(s_tab, p.theta, C, L);
    if constexpr (src_is_u8(SRC)) {
        for (int k = threadIdx.x; k < 256; k += blockDim.x) s_x[k] = __fdiv_rn(static_cast<float>(k), p.code_max);
    }
    __syncthreads();
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.stride;
    const float lm1 = static_cast<float>(L - 1);
    const bool gaussian = p.gaussian != 0;
    const int N = p.n_frames;
    const uint32_t n_items = static_cast<uint32_t>(p.plane / VEC);
    const uint32_t item_stride = gridDim.x * kBlock;
    const uint32_t first_item = blockIdx.x * kBlock + threadIdx.x;
    RowCursor<VEC> cur(first_item * VEC, item_stride * VEC, static_cast<uint32_t>(p.rows.base(c)), static_cast<uint32_t>(C));
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;

    for (uint32_t item = first_item; item < n_items; item += item_stride, cur.advance()) {
        const uint32_t pix = item * VEC;
        const int64_t off = static_cast<int64_t>(c) * p.stride + pix;
        uint32_t bias[VEC];
        {
            uint32_t u = cur.u0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) { bias[k] = tab_bias + u * row_bytes; u = (u + 1 == cur.C) ? 0u : u + 1; }
        }
        float wsum[VEC], wv[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { wsum[k] = 0.0f; wv[k] = 0.0f; }
        for (int n0 = 0; n0 < N; n0 += kFrameChunk) {
            Pack<VEC> xv[kFrameChunk], sv[kFrameChunk];
#pragma unroll
            for (int j = 0; j < kFrameChunk; ++j) {
                if (n0 + j < N) {
                    const int64_t o = off + static_cast<int64_t>(n0 + j) * frame_stride;
                    xv[j] = load_pixels<SRC, VEC>(p, n0 + j, c, pix, o, s_x);
                    sv[j] = load_std<VEC, STD>(p, o, xv[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < kFrameChunk; ++j) {
                if (n0 + j < N) {
                    const float it = p.scale.inv_t[n0 + j];
                    float *col = s_rq + (n0 + j) * (2 * VEC * kBlock);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const HdrTerms t = hdr_terms(xv[j].v[k], sv[j].v[k], it, has_model, gaussian, bias[k], lm1, true);
                        wsum[k] += t.w;
                        wv[k] = fmaf(t.w, t.v, wv[k]);
                        col[k * kBlock] = t.R;
                        col[(VEC + k) * kBlock] = t.Q;
                    }
                }
            }
        }
        hdr_finish<VEC, true, SINGLE>(p, off, wsum, wv, [&](int k, float alpha, float gamma) {
            float acc = 0.0f;                                 // one accumulator, frame order: the same bits as the register kernel
            const float *col = s_rq + k * kBlock;
            int n = 0;
            for (; n + 4 <= N; n += 4) {                      // four frames per trip: eight independent LDS in flight
                float r[4], q[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    r[j] = col[(n + j) * (2 * VEC * kBlock)];
                    q[j] = col[(n + j) * (2 * VEC * kBlock) + VEC * kBlock];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float g = SINGLE ? fmaf(gamma, q[j], r[j]) : fmaf(alpha, r[j], gamma * q[j]);
                    acc = fmaf(g, g, acc);
                }
            }
            for (; n < N; ++n) {
                const float rn = col[n * (2 * VEC * kBlock)], qn = col[n * (2 * VEC * kBlock) + VEC * kBlock];
                const float g = SINGLE ? fmaf(gamma, qn, rn) : fmaf(alpha, rn, gamma * qn);
                acc = fmaf(g, g, acc);
            }
            return acc;
        });
    }
}

// ---- fallback for N > kMaxFixedFrames: single pass, N dynamic --------------------------------------------
// sum_n (alpha R_n + gamma Q_n)^2 = alpha^2 SRR + 2 alpha gamma SRQ + gamma^2 SQQ needs only three running
// sums, but the expansion cancels (|gamma Q| can be ~14x the result), so the three sums are float64.
template <int VEC, int STD, int SRC>
__global__ void __launch_bounds__(kBlock) hdr_merge_kernel(const HdrParams p) {
    constexpr bool HAS_STD = STD != 0;
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    float *s_x = reinterpret_cast<float *>(s_tab + (has_model ? C * L : 0));
    if (has_model) stage_curve_slopes(s_tab, p.theta, C, L);
    if constexpr (src_is_u8(SRC)) {
        for (int k = threadIdx.x; k < 256; k += blockDim.x) s_x[k] = __fdiv_rn(static_cast<float>(k), p.code_max);
    }
    __syncthreads();
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.stride;
    const float lm1 = static_cast<float>(L - 1);
    const bool gaussian = p.gaussian != 0;
    const int N = p.n_frames;
    const uint32_t n_items = static_cast<uint32_t>(p.plane / VEC);
    const uint32_t item_stride = gridDim.x * kBlock;
    const uint32_t first_item = blockIdx.x * kBlock + threadIdx.x;
    RowCursor<VEC> cur(first_item * VEC, item_stride * VEC, static_cast<uint32_t>(p.rows.base(c)), static_cast<uint32_t>(C));
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;

    for (uint32_t item = first_item; item < n_items; item += item_stride, cur.advance()) {
        const uint32_t pix = item * VEC;
        const int64_t off = static_cast<int64_t>(c) * p.stride + pix;
        uint32_t bias[VEC];
        {
            uint32_t u = cur.u0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                bias[k] = tab_bias + u * row_bytes;
                u = (u + 1 == cur.C) ? 0u : u + 1;
            }
        }
        float wsum[VEC], wv[VEC];
        double srr[VEC], srq[VEC], sqq[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { wsum[k] = 0.0f; wv[k] = 0.0f; srr[k] = 0.0; srq[k] = 0.0; sqq[k] = 0.0; }
        for (int n0 = 0; n0 < N; n0 += kFrameChunk) {
            Pack<VEC> xv[kFrameChunk], sv[kFrameChunk];
#pragma unroll
            for (int j = 0; j < kFrameChunk; ++j) {
                if (n0 + j < N) {
                    const int64_t o = off + static_cast<int64_t>(n0 + j) * frame_stride;
                    xv[j] = load_pixels<SRC, VEC>(p, n0 + j, c, pix, o, s_x);
                    if constexpr (HAS_STD) sv[j] = load_std<VEC, STD>(p, o, xv[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < kFrameChunk; ++j) {
                if (n0 + j < N) {
                    const float it = p.scale.inv_t[n0 + j];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const HdrTerms t = hdr_terms(xv[j].v[k], HAS_STD ? sv[j].v[k] : 0.0f, it, has_model, gaussian,
                                                     bias[k], lm1, HAS_STD);
                        wsum[k] += t.w;
                        wv[k] = fmaf(t.w, t.v, wv[k]);
                        if constexpr (HAS_STD) {
                            const double R = static_cast<double>(t.R), Q = static_cast<double>(t.Q);
                            srr[k] = fma(R, R, srr[k]);
                            srq[k] = fma(R, Q, srq[k]);
                            sqq[k] = fma(Q, Q, sqq[k]);
                        }
                    }
                }
            }
        }
        hdr_finish<VEC, HAS_STD, false>(p, off, wsum, wv, [&](int k, float alpha, float gamma) {
            const double a = static_cast<double>(alpha), g = static_cast<double>(gamma);
            return static_cast<float>(fmax(a * a * srr[k] + 2.0 * a * g * srq[k] + g * g * sqq[k], 0.0));
        });
    }
}

// ---- launching ---------------------------------------------------------------------------------------------------
template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    return 0;
}

// Everything hdr_merge_impl has decided before it picks a kernel instantiation.
struct MergeLaunch {
    HdrParams p;
    size_t smem;
    int64_t want_blocks;     // blocks that would cover the plane once
    bool fixed, parked, single;
    int std_mode, n_frames, n_channels;
    cudaStream_t stream;
};

// persistent grid: a whole number of resident waves (blocks/SM from the occupancy calculator), split over the channels
// `fold`: the register kernels on the camera layout loop over the channels inside the thread (gridDim.y = 1)
template <typename K>
int launch_merge_kernel(K kernel, const MergeLaunch &m, bool fold = false) {
    if (int rc = ensure_smem(kernel, m.smem)) return rc;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, m.smem);
    // resident waves per persistent grid (measured): 2 for the register kernel up to 8 frames, 3 beyond and for the
    // shared-memory-parked kernel; 6 for planes of tens of megapixels (c4, values interleaved in one process: fp32 stack 878 ->
    // 869 us, uint16 codes 0.818 -> 0.802 ms, camera layout 0.883 -> 0.869 ms; 1080p stacks lose with more than 2)
    const int grid_channels = fold ? 1 : m.n_channels;
    per_sm = std::max(per_sm, 1);
    const bool big = m.want_blocks >= 64 * resident_blocks_per_channel(per_sm, grid_channels);
    per_sm *= g_tuning.hdr_waves > 0 ? g_tuning.hdr_waves : (big ? 6 : ((m.parked || m.n_frames > kMaxFixedFrames) ? 3 : 2));
    const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(m.want_blocks, resident_blocks_per_channel(per_sm, grid_channels)));
    kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(grid_channels)), kBlock, m.smem, m.stream>>>(m.p);
    return 0;
}

// The (N, std source, single-batch) dispatch for one (VEC, SRC) pair: register kernel up to 8 frames, shared-memory-parked
// kernel while it fits, float64-sum kernel otherwise.
template <int V, int SRC>
int launch_merge_by_std(const MergeLaunch &m) {
#define CLAIR_FIXED_NF(NF, ST) \
    (m.single ? launch_merge_kernel(hdr_merge_fixed_kernel<V, NF, ST, true, SRC>, m, src_is_hwc(SRC)) : launch_merge_kernel(hdr_merge_fixed_kernel<V, NF, ST, false, SRC>, m, src_is_hwc(SRC)))
#define CLAIR_FIXED(ST)                                 \
    switch (m.n_frames) {                               \
        case 1: return CLAIR_FIXED_NF(1, ST);           \
        case 2: return CLAIR_FIXED_NF(2, ST);           \
        case 3: return CLAIR_FIXED_NF(3, ST);           \
        case 4: return CLAIR_FIXED_NF(4, ST);           \
        case 5: return CLAIR_FIXED_NF(5, ST);           \
        case 6: return CLAIR_FIXED_NF(6, ST);           \
        case 7: return CLAIR_FIXED_NF(7, ST);           \
        case 8: return CLAIR_FIXED_NF(8, ST);           \
        default:                                        \
            if constexpr (SRC == kSrcF32 && V <= 2) {   \
                switch (m.n_frames) {                   \
                    case 9: return CLAIR_FIXED_NF(9, ST);   \
                    case 10: return CLAIR_FIXED_NF(10, ST); \
                    case 11: return CLAIR_FIXED_NF(11, ST); \
                    case 12: return CLAIR_FIXED_NF(12, ST); \
                    case 13: return CLAIR_FIXED_NF(13, ST); \
                    case 14: return CLAIR_FIXED_NF(14, ST); \
                    case 15: return CLAIR_FIXED_NF(15, ST); \
                    default: return CLAIR_FIXED_NF(16, ST); \
                }                                       \
            } else {                                    \
                return CLAIR_FIXED_NF(8, ST);           \
            }                                           \
    }
#define CLAIR_PARKED(ST) \
    (m.single ? launch_merge_kernel(hdr_merge_smem_kernel<V, ST, true, SRC>, m) : launch_merge_kernel(hdr_merge_smem_kernel<V, ST, false, SRC>, m))
    if (m.std_mode == kStdNone) return launch_merge_kernel(hdr_merge_kernel<V, 0, SRC>, m);
    if (m.std_mode == kStdTensor) {
        if (m.fixed) { CLAIR_FIXED(1) }
        if (m.parked) {
            if constexpr (SRC == kSrcF32 && V == 4) return fail(CLAIR_E_ARG, "hdr merge: the parked fp32 kernel is 1 or 2 pixels wide");
            else return CLAIR_PARKED(1);
        }
        return launch_merge_kernel(hdr_merge_kernel<V, 1, SRC>, m);
    }
    if constexpr (SRC == kSrcF32) {
        return fail(CLAIR_E_MODE, "hdr merge: synthesised std needs integer codes");
    } else {
        if (m.fixed) { CLAIR_FIXED(2) }
        if (m.parked) return CLAIR_PARKED(2);
        return launch_merge_kernel(hdr_merge_kernel<V, 2, SRC>, m);
    }
#undef CLAIR_PARKED
#undef CLAIR_FIXED
#undef CLAIR_FIXED_NF
}

// Integer ingest of 9..16 frames: the register kernel at 2 codes per thread (the packed pixel-pair path), so that the
// 9 x 24 MP 16-bit stacks of BASELINE config c4 do not fall to the shared-memory-parked kernel when they are handed over as
// codes.  Only the instantiations this range needs (std as a tensor or synthesised; N = 9..16).
template <int SRC>
int launch_merge_codes_wide(const MergeLaunch &m) {
#define CLAIR_WIDE_NF(NF, ST) \
    (m.single ? launch_merge_kernel(hdr_merge_fixed_kernel<2, NF, ST, true, SRC>, m, src_is_hwc(SRC)) : launch_merge_kernel(hdr_merge_fixed_kernel<2, NF, ST, false, SRC>, m, src_is_hwc(SRC)))
#define CLAIR_WIDE(ST)                                \
    switch (m.n_frames) {                             \
        case 9: return CLAIR_WIDE_NF(9, ST);          \
        case 10: return CLAIR_WIDE_NF(10, ST);        \
        case 11: return CLAIR_WIDE_NF(11, ST);        \
        case 12: return CLAIR_WIDE_NF(12, ST);        \
        case 13: return CLAIR_WIDE_NF(13, ST);        \
        case 14: return CLAIR_WIDE_NF(14, ST);        \
        case 15: return CLAIR_WIDE_NF(15, ST);        \
        case 16: return CLAIR_WIDE_NF(16, ST);        \
        default: return fail(CLAIR_E_ARG, "hdr merge: the 2-code register kernel takes 9..16 frames"); \
    }
    if (m.std_mode == kStdTensor) { CLAIR_WIDE(1) }
    if (m.std_mode == kStdMultiplier || m.std_mode == kStdConstant) { CLAIR_WIDE(2) }
    return fail(CLAIR_E_ARG, "hdr merge: the register kernels need a std source");
#undef CLAIR_WIDE
#undef CLAIR_WIDE_NF
}

// integer ingest, 4 codes per thread: planar (clair_merge_codes.cu) and interleaved BGR (clair_merge_codes_hwc.cu);
// 2 codes per thread for 9..16 frames (clair_merge_codes_wide.cu, clair_merge_codes_hwc_wide.cu)
int launch_merge_codes_planar(const MergeLaunch &m, bool u8);
int launch_merge_codes_hwc(const MergeLaunch &m, bool u8);
int launch_merge_codes_planar_wide(const MergeLaunch &m, bool u8);
int launch_merge_codes_hwc_wide(const MergeLaunch &m, bool u8);
int launch_merge_codes_hwc_wide_tma(const MergeLaunch &m, bool u8);

}  // namespace clair
