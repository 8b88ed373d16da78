// Pairwise exposure-ratio kernels: linearity statistics (forward) and the ICRF-table gradient (backward),
// plus the two tiny kernels that close a training step on the device (upstream factors, curve penalties).
//
// Both big kernels are FP32-issue-bound once there are more than a handful of exposure pairs (P grows as N^2/2),
// not HBM-bound: every frame element is read from HBM once per launch, its per-frame terms (ICRF value, sigma,
// Gaussian weight, validity) are evaluated once, and all pairs it takes part in reuse them from shared memory.
// See DESIGN.md §3.3-§3.5.
#include "clair_common.cuh"
#include "clair_host.h"

#include <cstdio>

namespace clair {

constexpr float kPairNegScaleLog2e = -14.426950408889634f;   // -10 * log2(e); training/losses.py:212 default scale 10
constexpr int kMaxPairsPerLaunch = 256;   // the pair table travels as a kernel argument
constexpr int kStatsTile = 128;           // pixels per shared-memory tile in the statistics kernels (power of two)
constexpr int kMeansTile = 256;           // ... of the packed kernel's training variant (sum w, sum w l only: registers to spare)
constexpr int kMaxSlots = 4;              // pairs a warp carries in registers in the statistics kernel
constexpr int kGradCopies = 64;           // replicated gradient tables the REDs of the scatter kernels are spread over
constexpr int kMaxGradCopies = 1024;      // the workspace holds this many
constexpr int kPairGradCopies = 1024;     // pair-gradient kernel: the L2 reductions are the bound, and hot table entries
                                          // serialise — c5 gradient 3.84 / 3.44 / 2.35 / 2.40 ms with 64 / 256 / 1024 / 4096 copies
constexpr int kFlushTiles = 8;            // tiles between flushes of the fp32 partial sums into float64 (32 terms per lane)

struct PairTable {
    float r_hi[kMaxPairsPerLaunch];       // exposure ratio t_i/t_j split into two floats (r = hi + lo to ~48 bits)
    float r_lo[kMaxPairsPerLaunch];
    uint8_t i[kMaxPairsPerLaunch];
    uint8_t j[kMaxPairsPerLaunch];
};

struct PairParams {
    const float *val;
    const float *std;            // nullptr when no per-pair error term is needed
    const float *theta;          // nullptr = identity
    double *sums;                // statistics: (P, C, 5), already offset to this launch's first pair
    const double *upstream;      // gradient: (P, C)
    const double *mean;          // gradient: (P, C)
    float *hist;                 // gradient: kGradCopies x 2 x C x (L + 2) fp32
    int64_t plane;
    int n_frames;
    int n_channels;
    int lut;
    int n_pairs;                 // pairs in this launch
    int unc_weighting;
    int n_copies;                // gradient: replicated tables in use
    int mode;                    // CLAIR_INTERP_* of the model the table belongs to
    int stats_buffers;           // statistics: tile buffers in shared memory (2 = staging overlaps the pair phase)
    float valid_lo, valid_hi;
    uint32_t mod_magic;          // ceil(2^16 / C): x mod C for x < 2^13 without a divide
    CurveRows rows;
    PairTable pairs;
};

__device__ __forceinline__ uint32_t mod_small(uint32_t x, uint32_t C, uint32_t magic) {
    return x - C * ((x * magic) >> 16);
}

// Per-frame quantities shared by all pairs a frame element takes part in.
//   f     linearised value                                     (models/base.py:182)
//   sig   |f'(x) * std|                                         (training/icrf_training.py:124)
//   gw    exp(-10 (x-.5)^2) of the RAW value, or -1 when the raw value is outside [valid_lo, valid_hi]
//         (training/losses.py:229-234 and common/general_functions.py:305 folded into one number)
//   xs    clamped scaled value x*(L-1): floor = lower LUT index, fraction = interpolation weight
struct FrameTerms {
    float f, sig, gw, xs;
};

template <bool HAS_STD>
__device__ __forceinline__ FrameTerms frame_terms(float x, float s, bool has_model, uint32_t row_bias, float lm1, float lo,
                                                  float hi) {
    FrameTerms t;
    float fp = 1.0f;
    t.f = x;
    t.xs = 0.0f;
    if (has_model) {
        icrf_linear_biased(x, row_bias, lm1, t.f, fp);
        t.xs = fminf(fmaxf(__fmul_rn(x, lm1), 0.0f), lm1);
    }
    t.sig = HAS_STD ? fabsf(__fmul_rn(fp, s)) : 0.0f;
    float d;
    const float g = gaussian_weight(x, kPairNegScaleLog2e, d);
    t.gw = (x >= lo && x <= hi) ? g : -1.0f;
    return t;
}

// The same terms for a LOOKUP / CATMULL model (models/base.py:138-158, :184-226); `row` is the table row the element
// reads (.x entries): the true channel row for LOOKUP, the k-mod-C row for CATMULL.
template <bool HAS_STD>
__device__ __forceinline__ FrameTerms frame_terms_mode(int mode, float x, float s, const float2 *row, int L, float lm1, float lo,
                                                       float hi) {
    FrameTerms t;
    float fp;
    icrf_mode_eval_rt(mode, x, row, L, lm1, t.f, fp);
    t.xs = fminf(fmaxf(__fmul_rn(x, lm1), 0.0f), lm1);
    t.sig = HAS_STD ? fabsf(__fmul_rn(fp, s)) : 0.0f;
    float d;
    const float g = gaussian_weight(x, kPairNegScaleLog2e, d);
    t.gw = (x >= lo && x <= hi) ? g : -1.0f;
    return t;
}

// a - b*r with r = r_hi + r_lo: two FMAs give the difference to fp32 relative accuracy OF THE DIFFERENCE
// (the reference forms it in float64, training/losses.py:40-42); the sign is what the gradient hinges on.
__device__ __forceinline__ float ratio_residual(float a, float b, float r_hi, float r_lo) {
    return fmaf(-b, r_lo, fmaf(-b, r_hi, a));
}

__device__ __forceinline__ float rcp_fast(float x) {     // MUFU.RCP + one Newton step: <= 1 ulp
    const float r = rcp_approx(x);
    return fmaf(fmaf(-x, r, 1.0f), r, r);
}

__device__ __forceinline__ float rsqrt_approx(float t) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

// =====================================================================================================
// Statistics kernel.  Block = W warps; a tile of kStatsTile pixels x N frames of one channel is staged in
// shared memory once, then warp w owns pairs w, w+W, ... (at most SLOTS of them) and keeps their running sums in
// registers across all tiles of the persistent loop.
//   FULL  = measure_linearity: sum w, sum w l, sum w l^2, sum err, count.  Per tile the three weighted sums are
//           accumulated in fp32 AROUND A PIVOT k (the first valid loss the thread saw): sum w(l-k), sum w(l-k)^2.
//           Their conversion to pivot 0 happens once per thread in float64, so the variance never sees the
//           k^2 sum(w) cancellation in fp32 and no per-element float64 / conversion instruction is needed.
//   !FULL = training: only sum w and sum w l.
//   ERR   = the per-pair uncertainty term is needed (std images present and either FULL or uncertainty weights).
// Per tile the fp32 partials are flushed into float64 running sums; one warp reduction and <= 5 fp64 atomics per
// (block, pair) at the very end.
// =====================================================================================================
template <int SLOTS, bool ERR, bool RELATIVE, bool FULL, int VA>
__global__ void __launch_bounds__(512) pair_stats_kernel(const PairParams p) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames;
    const bool has_model = p.theta != nullptr;
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw);
    // tile layout [frame][array][pixel]: one base address per frame, the arrays at constant offsets
    //   array 0: f   1: gw   2: sig (ERR)   3: sig / max(f, 1e-6) (ERR && RELATIVE)
    constexpr int kArr = 2 + (ERR ? (RELATIVE ? 2 : 1) : 0);
    constexpr int kFrameFloats = kArr * kStatsTile;
    float *s_tile = reinterpret_cast<float *>(s_tab + (has_model ? C * L : 0));
    if (has_model) stage_curve_pairs(s_tab, p.theta, C, L);

    const int c = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    const float lm1 = static_cast<float>(L - 1);
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const float *val_c = p.val + static_cast<int64_t>(c) * p.plane;
    const int64_t std_minus_val = ERR ? (p.std - p.val) : 0;      // std element = val element + this many floats
    const bool unc = p.unc_weighting != 0;
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const uint32_t uC = static_cast<uint32_t>(C);
    const uint32_t plane = static_cast<uint32_t>(p.plane);

    double d0[SLOTS], d1[SLOTS], d2[SLOTS], d3[SLOTS];
    float t0[SLOTS], t1[SLOTS], t2[SLOTS], t3[SLOTS];   // fp32 partials of the last <= kFlushTiles tiles
    float pivot[SLOTS];
    unsigned int cnt[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        d0[s] = 0.0; d1[s] = 0.0; d2[s] = 0.0; d3[s] = 0.0;
        t0[s] = 0.0f; t1[s] = 0.0f; t2[s] = 0.0f; t3[s] = 0.0f;
        pivot[s] = 0.0f; cnt[s] = 0u;
    }
    auto flush = [&]() {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            d0[s] += static_cast<double>(t0[s]); t0[s] = 0.0f;
            d1[s] += static_cast<double>(t1[s]); t1[s] = 0.0f;
            if constexpr (FULL) {
                d2[s] += static_cast<double>(t2[s]); t2[s] = 0.0f;
                if constexpr (ERR) { d3[s] += static_cast<double>(t3[s]); t3[s] = 0.0f; }
            }
        }
    };
    int since_flush = 0;

    const uint32_t n_tiles = (plane + kStatsTile - 1) / kStatsTile;
    // table row of the tile's first pixel, advanced incrementally (no divide in the loop)
    uint32_t u_tile = (blockIdx.x * kStatsTile + static_cast<uint32_t>(p.rows.base(c))) % uC;
    const uint32_t du_tile = (gridDim.x * kStatsTile) % uC;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();   // previous tile fully consumed (also orders the table staging on the first pass)
        const uint32_t pix0 = tile * kStatsTile;
        // phase A: one item = VA adjacent pixels of one frame (one 128-bit load per input when VA = 4)
        constexpr int kItemsPerFrame = kStatsTile / VA;
        constexpr int kItemShift = (VA == 4) ? 5 : 7;
        for (int item = threadIdx.x; item < N * kItemsPerFrame; item += blockDim.x) {
            const int n = item >> kItemShift;
            const uint32_t q = (static_cast<uint32_t>(item) & (kItemsPerFrame - 1)) * VA;
            const uint32_t pix = pix0 + q;
            const bool live = pix < plane;                     // plane % VA == 0: the whole item is in or out
            Pack<VA> xv, sv;
#pragma unroll
            for (int k = 0; k < VA; ++k) { xv.v[k] = 0.0f; sv.v[k] = 0.0f; }
            if (live) {
                const float *src = val_c + static_cast<int64_t>(n) * frame_stride + pix;
                xv = load_stream<VA>(src);
                if constexpr (ERR) sv = load_stream<VA>(src + std_minus_val);
            }
            uint32_t u = mod_small(u_tile + q, uC, p.mod_magic);
            Pack<VA> of, og, os, ob;
            const bool linear = p.mode == CLAIR_INTERP_LINEAR || !has_model;
#pragma unroll
            for (int k = 0; k < VA; ++k) {
                const FrameTerms t = linear
                    ? frame_terms<ERR>(xv.v[k], sv.v[k], has_model, tab_bias + u * row_bytes, lm1, p.valid_lo, p.valid_hi)
                    : frame_terms_mode<ERR>(p.mode, xv.v[k], sv.v[k], s_tab + (p.mode == CLAIR_INTERP_LOOKUP ? c : u) * L, L, lm1,
                                            p.valid_lo, p.valid_hi);
                u = (u + 1 == uC) ? 0u : u + 1;
                of.v[k] = t.f;
                og.v[k] = live ? t.gw : -1.0f;
                if constexpr (ERR) {
                    os.v[k] = t.sig;
                    if constexpr (RELATIVE) ob.v[k] = t.sig * rcp_fast(fmaxf(t.f, 1e-6f));   // losses.py:55,58
                }
            }
            float *dst = s_tile + n * kFrameFloats + q;
            store_shared<VA>(dst, of);
            store_shared<VA>(dst + kStatsTile, og);
            if constexpr (ERR) {
                store_shared<VA>(dst + 2 * kStatsTile, os);
                if constexpr (RELATIVE) store_shared<VA>(dst + 3 * kStatsTile, ob);
            }
        }
        u_tile += du_tile;
        u_tile = (u_tile >= uC) ? u_tile - uC : u_tile;
        __syncthreads();
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int pr = warp + s * n_warps;
            if (pr < p.n_pairs) {
                const float *fi = s_tile + p.pairs.i[pr] * kFrameFloats + lane;
                const float *fj = s_tile + p.pairs.j[pr] * kFrameFloats + lane;
                const float r_hi = p.pairs.r_hi[pr], r_lo = p.pairs.r_lo[pr];
                float a0 = t0[s], a1 = t1[s], a2 = t2[s], a3 = t3[s];
                float k = pivot[s];
                unsigned int n_valid = cnt[s];
#pragma unroll
                for (int it = 0; it < kStatsTile / 32; ++it) {
                    const int q = 32 * it;
                    const float a = fi[q], b = fj[q];
                    const float gi = fi[kStatsTile + q], gj = fj[kStatsTile + q];
                    const bool valid = (gi >= 0.0f) && (gj >= 0.0f);
                    const float d = ratio_residual(a, b, r_hi, r_lo);
                    float inv = 1.0f, ell;
                    if constexpr (RELATIVE) {
                        inv = rcp_fast(fmaf(b, r_hi, 1e-6f));                 // 1 / (expected + 1e-6), losses.py:45
                        ell = fabsf(d * inv);                                 // es can be negative once the curve dips below 0
                    } else {
                        ell = fabsf(d);
                    }
                    float wt = gi + gj;
                    float err = 0.0f;
                    if constexpr (ERR) {
                        const float sa = fi[2 * kStatsTile + q];
                        if constexpr (RELATIVE) {
                            const float e1 = sa * inv;
                            const float e2 = a * fj[3 * kStatsTile + q] * inv;
                            err = sqrt_approx(fmaf(e1, e1, fmaf(e2, e2, 1e-6f)));   // losses.py:57-60
                        } else {
                            const float rs = r_hi * fj[2 * kStatsTile + q];
                            err = sqrt_approx(fmaf(sa, sa, rs * rs));               // losses.py:62
                        }
                        if (unc) wt += rcp_fast(err + 1e-6f);                       // losses.py:97
                    }
                    const float w = valid ? wt : 0.0f;
                    if constexpr (FULL) {
                        k = (valid && n_valid == 0u) ? ell : k;
                        n_valid += valid ? 1u : 0u;
                        const float dl = valid ? ell - k : 0.0f;
                        const float wdl = w * dl;
                        a0 += w;
                        a1 += wdl;
                        a2 = fmaf(wdl, dl, a2);
                        if constexpr (ERR) a3 += valid ? err : 0.0f;
                    } else {
                        a0 += w;
                        a1 = fmaf(w, ell, a1);
                    }
                }
                t0[s] = a0; t1[s] = a1;
                if constexpr (FULL) {
                    t2[s] = a2;
                    if constexpr (ERR) t3[s] = a3;
                    pivot[s] = k;
                    cnt[s] = n_valid;
                }
            }
        }
        if (++since_flush == kFlushTiles) { flush(); since_flush = 0; }
    }
    flush();

#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int pr = warp + s * n_warps;
        if (pr < p.n_pairs) {     // warp-uniform
            // back to pivot 0 in float64:  sum w l = B + k W,  sum w l^2 = A + 2 k B + k^2 W
            const double kd = static_cast<double>(pivot[s]);
            double v0 = d0[s];
            double v1 = FULL ? d1[s] + kd * d0[s] : d1[s];
            double v2 = FULL ? d2[s] + 2.0 * kd * d1[s] + kd * kd * d0[s] : 0.0;
            double v3 = d3[s];
            unsigned int v4 = cnt[s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                if constexpr (FULL) {
                    v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                    if constexpr (ERR) v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                    v4 += __shfl_xor_sync(0xffffffffu, v4, o);
                }
            }
            if (lane == 0) {
                double *out = p.sums + (static_cast<int64_t>(pr) * C + c) * 5;
                atomicAdd(out + 0, v0);
                atomicAdd(out + 1, v1);
                if constexpr (FULL) {
                    atomicAdd(out + 2, v2);
                    if constexpr (ERR) atomicAdd(out + 3, v3);
                    atomicAdd(out + 4, static_cast<double>(v4));
                }
            }
        }
    }
}

// =====================================================================================================
// Statistics kernel, packed form (the default: H*W % 4 == 0, 16-byte aligned stacks).  Same block / tile / slot
// organisation as above, but every lane works on two adjacent pixels held as fp32x2 register pairs — the kernel
// is bound by instruction issue and FFMA2 / FMUL2 / FADD2 retire two fp32 results per slot:
//   * staging: one item = 4 pixels of one frame = two packed halves; table over (g0, g1 - g0);
//   * pair phase: 64-bit shared loads, validity folded into the Gaussian weight as kMaskedWeight (gi + gj < 0
//     <=> masked), 0/1 mask as a float pair so the masked sums are plain packed multiply-adds;
//   * FULL: the fp32 partial sums of both halves share one pivot per (lane, slot): the first valid loss seen.
// Measured on the c3 stack (16 x 4K, 29 pairs): 8930 -> 7540 warp instructions per tile (1.74 G -> 1.46 G per launch);
// c2 means pass 156 M -> 105 M.
// =====================================================================================================
#ifndef STATS_MAXT
#define STATS_MAXT(S) 512
#endif
#ifndef STATS_MINB
#define STATS_MINB(S) ((S) == 2 ? 2 : 1)
#endif
constexpr float kMaskedWeight = -1.0e30f;  // gw of a masked element: no sum with finite weights gets back above 0

__device__ __forceinline__ f32x2 rcp_fast2(f32x2 x) {     // 2 x MUFU.RCP + one packed Newton step
    float x0, x1;
    unpack2(x, x0, x1);
    const f32x2 r = pack2(rcp_approx(x0), rcp_approx(x1));
    return fma2(fma2(sub2(0ull, x), r, splat2(1.0f)), r, r);
}

template <bool HAS_STD, bool RELATIVE>
struct FrameTerms2 {
    f32x2 f, gw, sig, rel;
};

// table rows of a LOOKUP / CATMULL model for the two pixels (only read by the MODES instantiations)
struct ModeRows {
    const float2 *row0, *row1;
    int mode, L;
};

// per-frame terms of two adjacent pixels (see FrameTerms); `live` = the pixels exist.  MODES = the table belongs to a
// LOOKUP / CATMULL model (evaluated per pixel through `mr`); otherwise LINEAR over the (g0, g1 - g0) table.
template <bool HAS_STD, bool RELATIVE, bool MODES = false>
__device__ __forceinline__ FrameTerms2<HAS_STD, RELATIVE> frame_terms2(float x0, float x1, float s0, float s1, bool has_model,
                                                                      uint32_t bias0, uint32_t bias1, float lm1, float lo, float hi,
                                                                      bool live, const ModeRows mr = ModeRows{}) {
    FrameTerms2<HAS_STD, RELATIVE> t;
    const f32x2 x2 = pack2(x0, x1);
    float f0 = x0, f1 = x1, fp0 = 1.0f, fp1 = 1.0f;
    if constexpr (MODES) {
        icrf_mode_eval_rt(mr.mode, x0, mr.row0, mr.L, lm1, f0, fp0);
        icrf_mode_eval_rt(mr.mode, x1, mr.row1, mr.L, lm1, f1, fp1);
    } else if (has_model) {
        float r0, r1;
        unpack2(mul2(x2, splat2(lm1)), r0, r1);                          // image * (L - 1), rounded once
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
        f0 = fmaf(w0, dg0, g00);
        f1 = fmaf(w1, dg1, g01);
        if constexpr (HAS_STD) {
            fp0 = (xs0 == r0) ? __fmul_rn(dg0, lm1) : 0.0f;
            fp1 = (xs1 == r1) ? __fmul_rn(dg1, lm1) : 0.0f;
        }
    }
    const f32x2 d2 = add2(x2, splat2(-0.5f));
    float e0, e1;
    unpack2(mul2(mul2(d2, d2), splat2(kPairNegScaleLog2e)), e0, e1);
    const bool ok0 = live && x0 >= lo && x0 <= hi, ok1 = live && x1 >= lo && x1 <= hi;
    t.gw = pack2(ok0 ? exp2f_approx(e0) : kMaskedWeight, ok1 ? exp2f_approx(e1) : kMaskedWeight);
    // a masked pixel carries f = 1: every pair it takes part in then has a finite loss and error, and the pair phase
    // removes it with a 0 / 1 factor instead of selects (0 * Inf would be NaN)
    f0 = ok0 ? f0 : 1.0f;
    f1 = ok1 ? f1 : 1.0f;
    t.f = pack2(f0, f1);
    t.sig = 0ull;
    t.rel = 0ull;
    if constexpr (HAS_STD) {
        const float sg0 = ok0 ? fabsf(__fmul_rn(fp0, s0)) : 0.0f, sg1 = ok1 ? fabsf(__fmul_rn(fp1, s1)) : 0.0f;
        t.sig = pack2(sg0, sg1);
        // sigma / max(f, 1e-6), losses.py:55,58 (MUFU.RCP alone: <= 1 ulp, the statistics are gated at 1e-5)
        if constexpr (RELATIVE) t.rel = mul2(t.sig, pack2(rcp_approx(fmaxf(f0, 1e-6f)), rcp_approx(fmaxf(f1, 1e-6f))));
    }
    return t;
}

__device__ __forceinline__ float mul_sat(float a, float b) {        // clamp(a * b, 0, 1); NaN -> 0
    float r;
    asm("mul.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// TRIPS = staging items per thread (N * 32 items <= TRIPS * blockDim.x): an item's frame, tile offset, table rows and
// shared-memory slot never change, and the loads of the NEXT tile are issued before the pair phase of the current one,
// so their HBM latency is hidden behind it.
// MODES: the model is LOOKUP / CATMULL (only the staging phase differs: the pair phase works on the staged per-frame terms).
// TILE: pixels per tile — 128 for the FULL statistics (five packed accumulators per slot at 64 registers), 256 for the
// training variant (two accumulators: the larger tile halves the barriers and the per-tile bookkeeping; c2 means pass
// 0.167 -> 0.148 ms, c3-sized 1.02 -> 0.91 ms; the FULL kernel at 256 spills and loses: 1.69 -> 1.82 ms)
template <int SLOTS, bool ERR, bool RELATIVE, bool FULL, int TRIPS, bool MODES = false, int TILE = kStatsTile>
// (two-slot kernels are held to 64 registers = two 16-warp blocks per SM: c3 2.07 -> 1.93 ms with 48 B of spills)
__global__ void __launch_bounds__(STATS_MAXT(SLOTS), STATS_MINB(SLOTS)) pair_stats2_kernel(const PairParams p) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames;
    const bool has_model = p.theta != nullptr;
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw);
    // tile layout [frame][array][pixel]:  0: f   1: gw   2: sig (ERR)   3: sig / max(f, 1e-6) (ERR && RELATIVE)
    constexpr int kArr = 2 + (ERR ? (RELATIVE ? 2 : 1) : 0);
    constexpr int kFrameFloats = kArr * TILE;
    float *s_tile = reinterpret_cast<float *>(s_tab + (has_model ? C * L : 0));
    if (has_model) stage_curve_slopes(s_tab, p.theta, C, L);

    const int c = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    const float lm1 = static_cast<float>(L - 1);
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const float *val_c = p.val + static_cast<int64_t>(c) * p.plane;
    const int64_t std_minus_val = ERR ? (p.std - p.val) : 0;
    const bool unc = p.unc_weighting != 0;
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const uint32_t uC = static_cast<uint32_t>(C);
    const uint32_t plane = static_cast<uint32_t>(p.plane);

    // Running sums: fp32 partials (two halves per lane) of the last <= kFlushTiles tiles around a per-lane pivot, flushed
    // into float64 running sums in registers.  (Tried and measured slower on c3, 1.76 -> 1.89-1.97 ms: no float64 state in
    // registers, each flush a float64 warp reduction + atomics; pair constants hoisted out of the tile loop; one block of
    // 106 uncapped registers per SM; 8 warps x 4 slots at 80 / 126 registers.  The 64-register cap makes ptxas recompute
    // staging addresses and thread indices per tile, but 32 resident warps per SM are worth more than those instructions.)
    double d0[SLOTS], d1[SLOTS], d2[SLOTS], d3[SLOTS], d4[SLOTS];
    f32x2 t0[SLOTS], t1[SLOTS], t2[SLOTS], t3[SLOTS], t4[SLOTS];
    float pivot[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        d0[s] = 0.0; d1[s] = 0.0; d2[s] = 0.0; d3[s] = 0.0; d4[s] = 0.0;
        t0[s] = 0ull; t1[s] = 0ull; t2[s] = 0ull; t3[s] = 0ull; t4[s] = 0ull;
        pivot[s] = -1.0f;       // < 0: this lane has not seen a valid loss yet (losses are >= 0)
    }
    auto halves = [](f32x2 v) { float a, b; unpack2(v, a, b); return static_cast<double>(a) + static_cast<double>(b); };
    auto flush = [&]() {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            d0[s] += halves(t0[s]); t0[s] = 0ull;
            d1[s] += halves(t1[s]); t1[s] = 0ull;
            if constexpr (FULL) {
                d2[s] += halves(t2[s]); t2[s] = 0ull;
                if constexpr (ERR) { d3[s] += halves(t3[s]); t3[s] = 0ull; }
                d4[s] += halves(t4[s]); t4[s] = 0ull;
            }
        }
    };
    int since_flush = 0;

    const uint32_t n_tiles = (plane + TILE - 1) / TILE;
    // Per-thread staging items: everything that does not change from tile to tile is set up once — the source pointer only
    // advances by the grid's tile stride, the pixel position by the same (it is the bounds check), and because the host
    // sizes the grid so that its stride in pixels is a multiple of C, the table rows of an item's four pixels never change
    // (their biased shared-memory addresses are kept, not recomputed).  ncu's source page had 230 of 490 instructions per
    // tile in this phase, a third of them address arithmetic.
    const float *src[TRIPS];             // this thread's item in the tile that is loaded next
    uint32_t dst[TRIPS];                 // float offset of the item inside a tile buffer
    uint32_t pixn[TRIPS];                // first pixel of the item in the tile that is loaded next (>= plane: nothing left)
    uint32_t pixs[TRIPS];                // ... in the tile that is staged next
    uint32_t bias[TRIPS][4];             // biased table-row addresses of the item's four pixels
    const uint32_t tile_stride = gridDim.x * TILE;
    const bool rows_fixed = (tile_stride % uC) == 0u;
    uint32_t urow[TRIPS];
#pragma unroll
    for (int t = 0; t < TRIPS; ++t) {
        const int item = threadIdx.x + t * blockDim.x;
        constexpr int kItemsPerFrame = TILE / 4;
        const int n = item / kItemsPerFrame;
        const bool on = n < N;
        const uint32_t qoff = (static_cast<uint32_t>(item) & (kItemsPerFrame - 1)) * 4u;
        pixn[t] = on ? blockIdx.x * TILE + qoff : plane;
        pixs[t] = pixn[t];
        src[t] = val_c + static_cast<int64_t>(on ? n : 0) * frame_stride + blockIdx.x * TILE + qoff;
        dst[t] = static_cast<uint32_t>((on ? n : 0) * kFrameFloats) + qoff;
        urow[t] = (blockIdx.x * TILE + qoff + static_cast<uint32_t>(p.rows.base(c))) % uC;   // row of the item's first pixel
        uint32_t u = urow[t];
#pragma unroll
        for (int k = 0; k < 4; ++k) { bias[t][k] = tab_bias + u * row_bytes; u = (u + 1 == uC) ? 0u : u + 1; }
    }
    float4 xv[TRIPS], sv[TRIPS];
    auto prefetch = [&]() {
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            xv[t] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            sv[t] = xv[t];
            if (pixn[t] < plane) {                                             // plane % 4 == 0: the whole item is in or out
                xv[t] = __ldcs(reinterpret_cast<const float4 *>(src[t]));
                if constexpr (ERR) sv[t] = __ldcs(reinterpret_cast<const float4 *>(src[t] + std_minus_val));
            }
            // tile_stride < 2^32 - plane is checked on the host, so the position saturates instead of wrapping
            pixn[t] = (pixn[t] < plane) ? pixn[t] + tile_stride : plane;
            src[t] += tile_stride;
        }
    };
    // phase A: one item = 4 adjacent pixels of one frame, loaded one tile ahead; per-frame terms into tile buffer `buf`
    auto stage = [&](float *buf) {
#pragma unroll
        for (int t = 0; t < TRIPS; ++t) {
            if (threadIdx.x + t * blockDim.x < static_cast<unsigned>(N) * (TILE / 4)) {
                const bool live = pixs[t] < plane;
                pixs[t] = live ? pixs[t] + tile_stride : plane;
                if (!rows_fixed) {              // (grids smaller than C blocks per channel)
                    uint32_t u = urow[t];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { bias[t][k] = tab_bias + u * row_bytes; u = (u + 1 == uC) ? 0u : u + 1; }
                    urow[t] += tile_stride % uC;
                    urow[t] = (urow[t] >= uC) ? urow[t] - uC : urow[t];
                }
                ModeRows mra{}, mrb{};
                if constexpr (MODES) {
                    // LOOKUP reads the true channel row (base.py:148-158), CATMULL the k-mod-C row like LINEAR (:217-219)
                    const bool lookup = p.mode == CLAIR_INTERP_LOOKUP;
                    const float2 *rows[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) rows[k] = s_tab + (lookup ? static_cast<uint32_t>(c) : (bias[t][k] - tab_bias) / row_bytes) * L;
                    mra = ModeRows{rows[0], rows[1], p.mode, L};
                    mrb = ModeRows{rows[2], rows[3], p.mode, L};
                }
                const auto a = frame_terms2<ERR, RELATIVE, MODES>(xv[t].x, xv[t].y, sv[t].x, sv[t].y, has_model, bias[t][0], bias[t][1], lm1,
                                                                  p.valid_lo, p.valid_hi, live, mra);
                const auto b = frame_terms2<ERR, RELATIVE, MODES>(xv[t].z, xv[t].w, sv[t].z, sv[t].w, has_model, bias[t][2], bias[t][3], lm1,
                                                                  p.valid_lo, p.valid_hi, live, mrb);
                float *out = buf + dst[t];
                *reinterpret_cast<ulonglong2 *>(out) = make_ulonglong2(a.f, b.f);
                *reinterpret_cast<ulonglong2 *>(out + TILE) = make_ulonglong2(a.gw, b.gw);
                if constexpr (ERR) {
                    *reinterpret_cast<ulonglong2 *>(out + 2 * TILE) = make_ulonglong2(a.sig, b.sig);
                    if constexpr (RELATIVE) *reinterpret_cast<ulonglong2 *>(out + 3 * TILE) = make_ulonglong2(a.rel, b.rel);
                }
            }
        }
    };
    // phase B: warp w walks its pairs over the staged tile; all masking is arithmetic (vm = 0 / 1 per pixel)
    auto pairs = [&](const float *buf) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int pr = warp + s * n_warps;
            if (pr < p.n_pairs) {
                const float *fi = buf + p.pairs.i[pr] * kFrameFloats + lane * 2;
                const float *fj = buf + p.pairs.j[pr] * kFrameFloats + lane * 2;
                const float r_hi = p.pairs.r_hi[pr];
                const f32x2 nrh2 = splat2(-r_hi), nrl2 = splat2(-p.pairs.r_lo[pr]);
                f32x2 a0 = t0[s], a1 = t1[s], a2 = t2[s], a3 = t3[s], a4 = t4[s];
                float k = pivot[s];
#pragma unroll
                for (int it = 0; it < TILE / 64; ++it) {
                    const int q = 64 * it;
                    const f32x2 av = lds2(fi + q), bv = lds2(fj + q);
                    const f32x2 wg = add2(lds2(fi + TILE + q), lds2(fj + TILE + q));   // < 0 <=> masked
                    const f32x2 d = fma2(bv, nrl2, fma2(bv, nrh2, av));                          // a - b r
                    float g0, g1;
                    unpack2(wg, g0, g1);
                    const f32x2 vm = pack2(mul_sat(g0, 1.0e30f), mul_sat(g1, 1.0e30f));           // 1 valid, 0 masked
                    f32x2 wt = wg;                                                                // Gaussian part of the weight; x vm below
                    f32x2 inv = 0ull, ell2;
                    if constexpr (RELATIVE) {
                        float es0, es1;
                        unpack2(fma2(bv, splat2(r_hi), splat2(1e-6f)), es0, es1);                 // expected + 1e-6, losses.py:45
                        inv = pack2(rcp_approx(es0), rcp_approx(es1));
                        ell2 = mul2(d, inv);
                    } else {
                        ell2 = d;
                    }
                    ell2 &= 0x7fffffff7fffffffull;                                                // |.|: es can be negative once the curve dips below 0
                    f32x2 err = 0ull;
                    if constexpr (ERR) {
                        const f32x2 sa = lds2(fi + 2 * TILE + q);
                        if constexpr (RELATIVE) {
                            const f32x2 e1 = mul2(sa, inv);
                            const f32x2 e2 = mul2(mul2(av, lds2(fj + 3 * TILE + q)), inv);
                            const f32x2 T = fma2(e1, e1, fma2(e2, e2, splat2(1e-6f)));          // losses.py:57-60; T >= 1e-6
                            float T0, T1;
                            unpack2(T, T0, T1);
                            const f32x2 rs = pack2(rsqrt_approx(T0), rsqrt_approx(T1));           // 1 / err
                            err = mul2(T, rs);
                            if (unc) {                                                            // losses.py:97: 1 / (err + 1e-6)
                                // = rs / (1 + x), x = 1e-6 rs <= 1e-3: 1 - x + x^2 is exact to 1e-9
                                const f32x2 x = mul2(rs, splat2(1e-6f));
                                const f32x2 pq = sub2(fma2(x, x, splat2(1.0f)), x);
                                wt = fma2(rs, pq, wt);
                            }
                        } else {
                            const f32x2 rsb = mul2(lds2(fj + 2 * TILE + q), splat2(r_hi));
                            float s0, s1;
                            unpack2(fma2(sa, sa, mul2(rsb, rsb)), s0, s1);                        // losses.py:62
                            err = pack2(sqrt_approx(s0), sqrt_approx(s1));
                            if (unc) {
                                float u0, u1;
                                unpack2(add2(err, splat2(1e-6f)), u0, u1);
                                wt = add2(pack2(rcp_approx(u0), rcp_approx(u1)), wt);
                            }
                        }
                    }
                    wt = mul2(wt, vm);                                // masked pixels: finite (sanitised terms) times 0
                    if constexpr (FULL) {
                        {                                             // the pivot: the first valid loss this lane sees; < 0: none yet
                            float l0, l1, m0, m1;                     // (losses are >= 0; while there is none every weight was 0,
                            unpack2(ell2, l0, l1);                    //  so the value subtracted so far does not matter)
                            unpack2(vm, m0, m1);
                            const bool none = k < 0.0f;
                            k = (none && m1 != 0.0f) ? l1 : k;
                            k = (none && m0 != 0.0f) ? l0 : k;
                        }
                        const f32x2 dl = sub2(ell2, splat2(k));       // wt carries the mask
                        const f32x2 wdl = mul2(wt, dl);
                        a0 = add2(a0, wt);
                        a1 = add2(a1, wdl);
                        a2 = fma2(wdl, dl, a2);
                        if constexpr (ERR) a3 = fma2(err, vm, a3);
                        a4 = add2(a4, vm);
                    } else {
                        a0 = add2(a0, wt);
                        a1 = fma2(wt, ell2, a1);
                    }
                }
                t0[s] = a0; t1[s] = a1;
                if constexpr (FULL) {
                    t2[s] = a2;
                    if constexpr (ERR) t3[s] = a3;
                    t4[s] = a4;
                    pivot[s] = k;
                }
            }
        }
    };

    // Tile pipeline.  Two tile buffers (p.stats_buffers == 2): phase A of tile k+1 and phase B of tile k run between the
    // same pair of barriers, so a tile costs ONE block barrier and a warp goes from staging straight into its pairs; with
    // one buffer (very long stacks, where a second buffer would cost a resident block) it is A, barrier, B, barrier.
    const uint32_t nb = static_cast<uint32_t>(p.stats_buffers);
    const uint32_t buf_floats = static_cast<uint32_t>(N) * kFrameFloats;
    // (Tried: three buffers handed over through mbarriers instead of the block barrier — a warp then only waits for one a
    // whole pair phase behind.  Same sums, same time on c3 (1.69 ms): the barrier stalls ncu shows are not what bounds it.)
    prefetch();
    __syncthreads();                    // the table is staged
    uint32_t k_it = 0;
    for (uint32_t tile = blockIdx.x;; tile += gridDim.x, ++k_it) {
        const bool stage_live = tile < n_tiles;
        const bool pairs_live = (nb == 1) ? stage_live : (k_it > 0);
        if (!stage_live && !pairs_live) break;
        if (nb == 1) __syncthreads();   // previous tile fully consumed (also orders the table staging on the first pass)
        if (stage_live) {
            stage(s_tile + (nb == 2 ? (k_it & 1u) * buf_floats : 0u));
            prefetch();
        }
        if (nb == 1) __syncthreads();
        if (pairs_live) {
            pairs(s_tile + (nb == 2 ? ((k_it + 1u) & 1u) * buf_floats : 0u));
            if (++since_flush == kFlushTiles) { flush(); since_flush = 0; }
        }
        if (nb == 2) __syncthreads();
        if (!stage_live) break;
    }
    flush();

#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int pr = warp + s * n_warps;
        if (pr < p.n_pairs) {     // warp-uniform
            // back to pivot 0 in float64:  sum w l = B + k W,  sum w l^2 = A + 2 k B + k^2 W
            const double kd = static_cast<double>(pivot[s]);
            double v0 = d0[s];
            double v1 = FULL ? d1[s] + kd * d0[s] : d1[s];
            double v2 = FULL ? d2[s] + 2.0 * kd * d1[s] + kd * kd * d0[s] : 0.0;
            double v3 = d3[s];
            double v4 = d4[s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                if constexpr (FULL) {
                    v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                    if constexpr (ERR) v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                    v4 += __shfl_xor_sync(0xffffffffu, v4, o);
                }
            }
            if (lane == 0) {
                double *out = p.sums + (static_cast<int64_t>(pr) * C + c) * 5;
                atomicAdd(out + 0, v0);
                atomicAdd(out + 1, v1);
                if constexpr (FULL) {
                    atomicAdd(out + 2, v2);
                    if constexpr (ERR) atomicAdd(out + 3, v3);
                    atomicAdd(out + 4, v4);
                }
            }
        }
    }
}

// =====================================================================================================
// Gradient kernel.  Each warp owns 32 pixels of one channel at a time: the per-frame terms of those pixels
// go to the warp's private shared-memory slice, every pair is visited by the same lane that owns the pixel
// (so the per-frame upstream G[n] accumulates without atomics; runs of pairs with the same first frame keep
// that frame's sum in a register), and each frame element ends with ONE vectorised reduction
// red.global.add.v2.f32 {G(1-w), G w} into one of kGradCopies replicated tables.
// ERR = the weights contain the inverse-uncertainty term (std images present and uncertainty weighting on).
// =====================================================================================================
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Table layout per copy: A rows then B rows, each C x (L+2).  An element with even x0 adds its two taps at
// A[u][x0], A[u][x0+1]; one with odd x0 adds them at B[u][x0+1], B[u][x0+2] (8-byte aligned in both cases).
// grad[u][k] = A[u][k] + B[u][k+1].
__device__ __forceinline__ void scatter_taps(float *copy, int C, int L, int u, float xs, float g) {
    int x0;
    const float fl = floor_small(xs, x0);
    const float w = xs - fl;
    const int lp = L + 2;
    float *base = copy + ((x0 & 1) ? (C * lp + u * lp + x0 + 1) : (u * lp + x0));
    red_add_v2(base, g * (1.0f - w), g * w);
}

// LOOKUP / CATMULL models: one tap of weight 1 at the nearest sample (the gather of models/base.py:158 under autograd) or
// the four Catmull-Rom taps (:219-226), plain fp32 reductions into the A half of a replicated table.  `row` = the true
// channel for LOOKUP, the k-mod-C row for CATMULL; `xs` = clamp(x (L-1), 0, L-1).
__device__ __forceinline__ void scatter_taps_mode(float *copy, int L, int mode, int row, float xs, float g) {
    float *dst = copy + row * (L + 2);
    if (mode == CLAIR_INTERP_LOOKUP) {
        atomicAdd(dst + static_cast<int>(rintf(xs)), g);        // rint(clamp(.)) == clamp(rint(.)): the bounds are integers
    } else {
        const CatmullTaps t = catmull_taps_xs(xs, true, L);
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(dst + t.idx[k], g * t.w[k]);
    }
}

template <bool ERR, bool RELATIVE, int PIX>
__global__ void __launch_bounds__(256) pair_grad_kernel(const PairParams p) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames, P = p.n_pairs;
    const int c = blockIdx.y;
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw);
    float *s_up = reinterpret_cast<float *>(s_tab + C * L);      // U[p, c] and mean[p, c] of this channel
    float *s_mean = s_up + P;
    stage_curve_pairs(s_tab, p.theta, C, L);
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        s_up[k] = static_cast<float>(p.upstream[static_cast<int64_t>(k) * C + c]);
        s_mean[k] = static_cast<float>(p.mean[static_cast<int64_t>(k) * C + c]);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    // warp-private slice, layout [frame][array][lane][PIX]: one base address per frame, arrays at constant offsets
    //   0: f   1: gw   2: xs   3: G (per-frame upstream)   4: sigma (ERR)   5: 1 / max(f, 1e-6) (ERR)
    constexpr int kArrays = ERR ? 6 : 4;
    constexpr int kArr = 32 * PIX;                       // floats per array
    constexpr int kFrameFloats = kArrays * kArr;
    // s_up and s_mean hold 2P floats after the float2 table, so the slices stay 8-byte aligned for paired accesses
    float *slice = s_mean + P + warp * (N * kFrameFloats) + lane * PIX;

    const float lm1 = static_cast<float>(L - 1);
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const float *val_c = p.val + static_cast<int64_t>(c) * p.plane;
    const int64_t std_minus_val = ERR ? (p.std - p.val) : 0;
    const int lp = L + 2;
    float *copy = p.hist + static_cast<int64_t>((blockIdx.x * n_warps + warp) % p.n_copies) * (2 * C * lp);
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const uint32_t uC = static_cast<uint32_t>(C);
    const uint32_t plane = static_cast<uint32_t>(p.plane);

    constexpr uint32_t kGroup = 32 * PIX;                // pixels per warp trip
    const uint32_t n_groups = (plane + kGroup - 1) / kGroup;
    const uint32_t grp_stride = gridDim.x * n_warps;
    uint32_t grp = blockIdx.x * n_warps + warp;
    uint32_t u0 = (grp * kGroup + lane * PIX + static_cast<uint32_t>(p.rows.base(c))) % uC;
    const uint32_t du = (grp_stride * kGroup) % uC;
    for (; grp < n_groups; grp += grp_stride) {
        const uint32_t pix = grp * kGroup + lane * PIX;
        const bool live = pix < plane;                   // plane % PIX == 0
        uint32_t bias[PIX];
        {
            uint32_t u = u0;
#pragma unroll
            for (int k = 0; k < PIX; ++k) { bias[k] = tab_bias + u * row_bytes; u = (u + 1 == uC) ? 0u : u + 1; }
        }
        __syncwarp();
        // loads of four frames in flight together, then their arithmetic
        for (int n0 = 0; n0 < N; n0 += 4) {
            Pack<PIX> xv[4], sv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int k = 0; k < PIX; ++k) { xv[j].v[k] = 0.0f; sv[j].v[k] = 0.0f; }
                if (n0 + j < N && live) {
                    const float *src = val_c + static_cast<int64_t>(n0 + j) * frame_stride + pix;
                    xv[j] = load_stream<PIX>(src);
                    if constexpr (ERR) sv[j] = load_stream<PIX>(src + std_minus_val);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n0 + j < N) {
                    Pack<PIX> of, og, ox, oz, os, oi;
#pragma unroll
                    for (int k = 0; k < PIX; ++k) {
                        const FrameTerms t = (p.mode == CLAIR_INTERP_LINEAR)
                            ? frame_terms<ERR>(xv[j].v[k], sv[j].v[k], true, bias[k], lm1, p.valid_lo, p.valid_hi)
                            : frame_terms_mode<ERR>(p.mode, xv[j].v[k], sv[j].v[k],
                                                    s_tab + (p.mode == CLAIR_INTERP_LOOKUP ? static_cast<uint32_t>(c) : (u0 + k) % uC) * L,
                                                    L, lm1, p.valid_lo, p.valid_hi);
                        of.v[k] = t.f; og.v[k] = live ? t.gw : -1.0f; ox.v[k] = t.xs; oz.v[k] = 0.0f;
                        if constexpr (ERR) { os.v[k] = t.sig; oi.v[k] = rcp_approx(fmaxf(t.f, 1e-6f)); }
                    }
                    float *dst = slice + (n0 + j) * kFrameFloats;
                    store_shared<PIX>(dst, of);
                    store_shared<PIX>(dst + kArr, og);
                    store_shared<PIX>(dst + 2 * kArr, ox);
                    store_shared<PIX>(dst + 3 * kArr, oz);
                    if constexpr (ERR) {
                        store_shared<PIX>(dst + 4 * kArr, os);
                        store_shared<PIX>(dst + 5 * kArr, oi);
                    }
                }
            }
        }
        __syncwarp();
        int cur_i = -1;
        Pack<PIX> acc_i;
#pragma unroll
        for (int k = 0; k < PIX; ++k) acc_i.v[k] = 0.0f;
        auto flush_i = [&]() {
            float *gi_ptr = slice + cur_i * kFrameFloats + 3 * kArr;
            Pack<PIX> g = load_shared<PIX>(gi_ptr);
#pragma unroll
            for (int k = 0; k < PIX; ++k) { g.v[k] += acc_i.v[k]; acc_i.v[k] = 0.0f; }
            store_shared<PIX>(gi_ptr, g);
        };
        for (int pr = 0; pr < P; ++pr) {
            const int pi = p.pairs.i[pr];
            if (pi != cur_i) {                                   // uniform across the warp
                if (cur_i >= 0) flush_i();
                cur_i = pi;
            }
            const float *fi = slice + pi * kFrameFloats;
            float *fj = slice + p.pairs.j[pr] * kFrameFloats;
            const float r_hi = p.pairs.r_hi[pr], r_lo = p.pairs.r_lo[pr];
            const float up_pr = s_up[pr];
            const Pack<PIX> gi = load_shared<PIX>(fi + kArr), gj = load_shared<PIX>(fj + kArr);
            const Pack<PIX> av = load_shared<PIX>(fi), bv = load_shared<PIX>(fj);
            Pack<PIX> gjacc = load_shared<PIX>(fj + 3 * kArr);
            Pack<PIX> sav, sbv, ibv;
            if constexpr (ERR) {
                sav = load_shared<PIX>(fi + 4 * kArr);
                sbv = load_shared<PIX>(fj + 4 * kArr);
                if constexpr (RELATIVE) ibv = load_shared<PIX>(fj + 5 * kArr);
            }
#pragma unroll
            for (int k = 0; k < PIX; ++k) {
                const bool valid = (gi.v[k] >= 0.0f) && (gj.v[k] >= 0.0f);     // masked pair elements carry no gradient
                const float up = valid ? up_pr : 0.0f;
                const float a = av.v[k], b = bv.v[k];
                const float d = ratio_residual(a, b, r_hi, r_lo);
                float wt = gi.v[k] + gj.v[k];
                float ga, gb;
                if constexpr (RELATIVE) {
                    // MUFU.RCP alone (~1 ulp): the gate on the gradient is 1e-5 of its maximum
                    const float inv = rcp_approx(fmaf(b, r_hi, 1e-6f));
                    const float q = d * inv;
                    float extra_a = 0.0f, extra_b = 0.0f;
                    if constexpr (ERR) {
                        // the inverse-uncertainty weight depends on the curve through a, es and max(b, 1e-6)
                        const float e1 = sav.v[k] * inv;
                        const float c2 = sbv.v[k] * ibv.v[k] * inv;
                        const float e2 = a * c2;
                        const float T = fmaf(e1, e1, fmaf(e2, e2, 1e-6f));
                        const float rerr = rsqrt_approx(T);                    // 1 / err
                        const float rw = rcp_approx(fmaf(T, rerr, 1e-6f));     // 1 / (err + 1e-6)
                        wt += rw;
                        // dm/dWt * dWt/derr * derr/dT = (l - m) U * (-rw^2) * 1/(2 err)
                        const float kk = (fabsf(q) - s_mean[pr]) * up * (-0.5f * rw * rw) * rerr;
                        const float dT_da = 2.0f * e2 * c2;
                        const float dT_des = -2.0f * inv * (e1 * e1 + e2 * e2);
                        const float dT_dbs = (b >= 1e-6f) ? (-2.0f * e2 * e2 * ibv.v[k]) : 0.0f;
                        extra_a = kk * dT_da;
                        extra_b = kk * fmaf(dT_des, r_hi, dT_dbs);
                    }
                    // sign(q) * (wt U / es) with sign(0) = 0, as torch.abs' backward has it
                    const float mag = (q != 0.0f) ? wt * up * inv : 0.0f;
                    const float base = __int_as_float((__float_as_int(mag) ^ (__float_as_int(q) & 0x80000000)));
                    ga = base + extra_a;                                   // dl/da = sgn / es
                    gb = fmaf(-base * r_hi, (a + 1e-6f) * inv, extra_b);   // dl/db = -sgn r (a + 1e-6) / es^2
                } else {
                    if constexpr (ERR) {
                        const float rs = r_hi * sbv.v[k];
                        wt += rcp_approx(sqrt_approx(fmaf(sav.v[k], sav.v[k], rs * rs)) + 1e-6f);   // constant wrt the curve
                    }
                    const float mag = (d != 0.0f) ? wt * up : 0.0f;
                    const float base = __int_as_float((__float_as_int(mag) ^ (__float_as_int(d) & 0x80000000)));
                    ga = base;
                    gb = -base * r_hi;
                }
                acc_i.v[k] += ga;
                gjacc.v[k] += gb;
            }
            store_shared<PIX>(fj + 3 * kArr, gjacc);
        }
        if (cur_i >= 0) flush_i();
        __syncwarp();
        for (int n = 0; n < N; ++n) {
            const Pack<PIX> g = load_shared<PIX>(slice + n * kFrameFloats + 3 * kArr);
            const Pack<PIX> xs = load_shared<PIX>(slice + n * kFrameFloats + 2 * kArr);
            uint32_t u = u0;
#pragma unroll
            for (int k = 0; k < PIX; ++k) {
                if (g.v[k] != 0.0f) {
                    if (p.mode == CLAIR_INTERP_LINEAR) scatter_taps(copy, C, L, static_cast<int>(u), xs.v[k], g.v[k]);
                    else scatter_taps_mode(copy, L, p.mode, p.mode == CLAIR_INTERP_LOOKUP ? c : static_cast<int>(u), xs.v[k], g.v[k]);
                }
                u = (u + 1 == uC) ? 0u : u + 1;
            }
        }
        u0 += du;
        u0 = (u0 >= uC) ? u0 - uC : u0;
    }
}

// =====================================================================================================
// Gradient kernel, packed form (the default: H*W even, 8-byte aligned stacks).  Same algorithm as above with two
// adjacent pixels per lane held as fp32x2 register pairs, so the add / mul / fma chains issue as FFMA2 / FMUL2 /
// FADD2 — the kernel is bound by instruction issue and shared-memory bandwidth, not by HBM.  Differences:
//   * the body of a trip is lane-private (the slice columns are [lane][2]): no warp synchronisation, the last
//     partial group simply drops its dead lanes;
//   * validity is folded into the Gaussian weight as a large negative number: gi + gj < 0 <=> the pair element is
//     masked, and max(gi + gj, 0) is the weight (one FMNMX instead of two compares and two selects);
//   * a run of pairs with the same first frame keeps that frame's terms and its gradient sum in registers
//     (4 shared-memory accesses per pair instead of 6);
//   * table over (g0, g1 - g0): f = g0 + w (g1 - g0) in one FMA (within 1 ulp of the reference's form);
//   * frame pointers advance by a constant, tap addresses are 32-bit offsets from one base: a trip of N = 2,
//     P = 1 went from 459 to 283 warp instructions, the c2 launch from 253 M to 156 M.
// =====================================================================================================

// One predicated vector reduction per element: both taps in one L2 operation into one of kPairGradCopies replicated
// tables.  red.global.add.v2.f32 sustains 176 G operations/s on random table entries and ~110 G/s on 16 hot ones
// (scratch/smem_atomics.cu); hot entries are what real exposure stacks produce (a dark frame puts a warp's 64 pixels on
// a few dozen entries), hence the 1024 copies.  Block-private tables in shared memory were tried three ways and all
// lost on the c2 / c5 stacks (0.35 / 2.35 ms in this form):
//   * fp32 atomicAdd x2 — sm_100a has no native floating-point shared atomic, it is an ATOMS.CAST.SPIN loop: 0.52 / 3.19 ms;
//   * one 64-bit CAS loop per element: collapses on hot entries, 3.1 / 8.3 ms;
//   * 64-bit fixed point on native integer shared atomics (returning add on the low word, carry into the high word,
//     pipelined one frame deep; > 790 G/s in isolation): 0.44 / 3.18 ms — the 12 KB table costs a resident block and the
//     kernel is then bound by latency at 12-15 warps per SM, not by the reductions.
// One copy per lane instead of per warp is slower too (0.38 ms): lanes of a warp instruction that meet in a sector are merged.
__device__ __forceinline__ void red_add_v2_if(float *addr, float a, float b, float g) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.neu.f32 p, %3, 0f00000000;\n\t@p red.global.add.v2.f32 [%0], {%1, %2};\n\t}"
                 ::"l"(addr), "f"(a), "f"(b), "f"(g));
}

// MODES: the model is LOOKUP / CATMULL — per-frame terms and the final tap scatter differ, the pair algebra does not.
// FUSED (one exposure pair, no uncertainty weights: the "100 MP exposure pair" of BASELINE config 5): statistics and
// gradient in ONE pass over the stack.  With a single pair the upstream factor U[c] = dLoss_c / dmean / D is one scalar
// per channel, so the kernel scatters the un-normalised gradient (U = 1) into a table PER CHANNEL BLOCK c — the table
// rows u interleave channels (Q1), hence C tables of C rows — and accumulates sum M Wt, sum M Wt l beside it;
// clair_pair_fused_combine applies U afterwards (after the all-reduce when the image is sharded).
template <bool ERR, bool RELATIVE, bool MODES = false, bool FUSED = false>
__global__ void __launch_bounds__(256) pair_grad2_kernel(const PairParams p) {
    static_assert(!(FUSED && ERR), "the fused single-pair pass has no uncertainty weights");
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames, P = p.n_pairs;
    const int c = blockIdx.y;
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw);
    float *s_up = reinterpret_cast<float *>(s_tab + C * L);      // U[p, c] and mean[p, c] of this channel
    float *s_mean = s_up + P;
    const uint32_t lp = static_cast<uint32_t>(L + 2);
    stage_curve_slopes(s_tab, p.theta, C, L);
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        s_up[k] = FUSED ? 1.0f : static_cast<float>(p.upstream[static_cast<int64_t>(k) * C + c]);
        s_mean[k] = FUSED ? 0.0f : static_cast<float>(p.mean[static_cast<int64_t>(k) * C + c]);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    // FUSED: running sum M Wt, sum M Wt l of the single pair (fp32 per lane for <= 16 trips, then float64)
    f32x2 fs0 = 0ull, fs1 = 0ull;
    double fd0 = 0.0, fd1 = 0.0;
    int fused_trips = 0;
    // warp-private slice, layout [frame][array][lane][2]
    //   0: f   1: gw   2: xs   3: G (per-frame upstream)   4: sigma (ERR)   5: 1 / max(f, 1e-6) (ERR && RELATIVE)
    constexpr int kArrays = ERR ? 6 : 4;
    constexpr int kArr = 64;
    constexpr int kFrameFloats = kArrays * kArr;
    float *slice = s_mean + P + warp * (N * kFrameFloats) + lane * 2;

    const float lm1 = static_cast<float>(L - 1);
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const float *val_c = p.val + static_cast<int64_t>(c) * p.plane;
    const int64_t std_minus_val = ERR ? (p.std - p.val) : 0;
    const int64_t copy_index = FUSED ? static_cast<int64_t>((blockIdx.x * n_warps + warp) % p.n_copies) * C + c
                                     : static_cast<int64_t>((blockIdx.x * n_warps + warp) % p.n_copies);
    float *copy = p.hist + copy_index * (2 * C * lp);
    const uint32_t odd_step = static_cast<uint32_t>(C) * lp + 1u;     // A[u][x0] -> B[u][x0 + 1]
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const uint32_t uC = static_cast<uint32_t>(C);
    const uint32_t plane = static_cast<uint32_t>(p.plane);
    const float lo = p.valid_lo, hi = p.valid_hi;
    const f32x2 two23 = splat2(8388608.0f);

    const uint32_t n_groups = (plane + 63u) / 64u;
    const uint32_t grp_stride = gridDim.x * n_warps;
    uint32_t grp = blockIdx.x * n_warps + warp;
    uint32_t u0 = (grp * 64u + lane * 2u + static_cast<uint32_t>(p.rows.base(c))) % uC;
    const uint32_t du = (grp_stride * 64u) % uC;
    for (; grp < n_groups; grp += grp_stride) {
        const uint32_t pix = grp * 64u + lane * 2u;
        if (pix >= plane) break;                         // plane % 2 == 0; nothing below crosses lanes
        const uint32_t u1 = (u0 + 1u == uC) ? 0u : u0 + 1u;
        const uint32_t bias0 = tab_bias + u0 * row_bytes, bias1 = tab_bias + u1 * row_bytes;

        // ---- per-frame terms into the slice ----
        auto stage = [&](int n, float2 xin, float2 sin) {
            const f32x2 x2 = pack2(xin.x, xin.y);
            float r0, r1;
            unpack2(mul2(x2, splat2(lm1)), r0, r1);                       // image * (L - 1), rounded once
            const float xs0 = fminf(fmaxf(r0, 0.0f), lm1), xs1 = fminf(fmaxf(r1, 0.0f), lm1);
            const f32x2 xs2 = pack2(xs0, xs1);
            float f0, f1, fp0 = 0.0f, fp1 = 0.0f;
            if constexpr (MODES) {
                const bool lookup = p.mode == CLAIR_INTERP_LOOKUP;
                icrf_mode_eval_rt(p.mode, xin.x, s_tab + (lookup ? static_cast<uint32_t>(c) : u0) * L, L, lm1, f0, fp0);
                icrf_mode_eval_rt(p.mode, xin.y, s_tab + (lookup ? static_cast<uint32_t>(c) : u1) * L, L, lm1, f1, fp1);
            } else {
                const f32x2 t2 = add2_rd(xs2, two23);
                const f32x2 w2 = sub2(xs2, sub2(t2, two23));
                float t0, t1, w0, w1;
                unpack2(t2, t0, t1);
                unpack2(w2, w0, w1);
                float g00, dg0, g01, dg1;
                asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(g00), "=f"(dg0) : "r"(static_cast<uint32_t>(__float_as_int(t0)) * 8u + bias0));
                asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(g01), "=f"(dg1) : "r"(static_cast<uint32_t>(__float_as_int(t1)) * 8u + bias1));
                f0 = fmaf(w0, dg0, g00);
                f1 = fmaf(w1, dg1, g01);
                if constexpr (ERR) {
                    fp0 = (xs0 == r0) ? __fmul_rn(dg0, lm1) : 0.0f;
                    fp1 = (xs1 == r1) ? __fmul_rn(dg1, lm1) : 0.0f;
                }
            }
            const f32x2 d2 = add2(x2, splat2(-0.5f));
            float e0, e1;
            unpack2(mul2(mul2(d2, d2), splat2(kPairNegScaleLog2e)), e0, e1);
            const float gw0 = (xin.x >= lo && xin.x <= hi) ? exp2f_approx(e0) : kMaskedWeight;
            const float gw1 = (xin.y >= lo && xin.y <= hi) ? exp2f_approx(e1) : kMaskedWeight;
            float *dst = slice + n * kFrameFloats;
            sts2(dst, pack2(f0, f1));
            sts2(dst + kArr, pack2(gw0, gw1));
            sts2(dst + 2 * kArr, xs2);
            sts2(dst + 3 * kArr, 0ull);
            if constexpr (ERR) {
                sts2(dst + 4 * kArr, pack2(fabsf(__fmul_rn(fp0, sin.x)), fabsf(__fmul_rn(fp1, sin.y))));
                if constexpr (RELATIVE) sts2(dst + 5 * kArr, pack2(rcp_approx(fmaxf(f0, 1e-6f)), rcp_approx(fmaxf(f1, 1e-6f))));
            }
        };
        {
            const float *src = val_c + pix;
            int n = 0;
            for (; n + 4 <= N; n += 4) {                 // four frames' loads in flight
                float2 xv[4], sv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    xv[j] = __ldcs(reinterpret_cast<const float2 *>(src));
                    sv[j] = ERR ? __ldcs(reinterpret_cast<const float2 *>(src + std_minus_val)) : make_float2(0.0f, 0.0f);
                    src += frame_stride;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) stage(n + j, xv[j], sv[j]);
            }
            if (n + 2 <= N) {
                float2 xv[2], sv[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    xv[j] = __ldcs(reinterpret_cast<const float2 *>(src));
                    sv[j] = ERR ? __ldcs(reinterpret_cast<const float2 *>(src + std_minus_val)) : make_float2(0.0f, 0.0f);
                    src += frame_stride;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) stage(n + j, xv[j], sv[j]);
                n += 2;
            }
            if (n < N) {
                const float2 xv = __ldcs(reinterpret_cast<const float2 *>(src));
                const float2 sv = ERR ? __ldcs(reinterpret_cast<const float2 *>(src + std_minus_val)) : make_float2(0.0f, 0.0f);
                stage(n, xv, sv);
            }
        }

        // ---- all pairs of the lane's two pixels ----
        int cur_i = -1;
        const float *fi = slice;
        f32x2 acc_i = 0ull, a2 = 0ull, gi2 = 0ull, sa2 = 0ull;
        for (int pr = 0; pr < P; ++pr) {
            const int pi = p.pairs.i[pr];
            if (pi != cur_i) {                                   // uniform across the warp
                if (cur_i >= 0) sts2(const_cast<float *>(fi) + 3 * kArr, add2(lds2(fi + 3 * kArr), acc_i));
                cur_i = pi;
                fi = slice + pi * kFrameFloats;
                acc_i = 0ull;
                a2 = lds2(fi);
                gi2 = lds2(fi + kArr);
                if constexpr (ERR) sa2 = lds2(fi + 4 * kArr);
            }
            float *fj = slice + p.pairs.j[pr] * kFrameFloats;
            const float r_hi = p.pairs.r_hi[pr];
            const f32x2 nrh2 = splat2(-r_hi), nrl2 = splat2(-p.pairs.r_lo[pr]);
            const float up_pr = s_up[pr];
            const f32x2 b2 = lds2(fj);
            const f32x2 wg2 = add2(gi2, lds2(fj + kArr));                    // < 0 <=> masked
            const f32x2 gj_acc = lds2(fj + 3 * kArr);
            const f32x2 d2 = fma2(b2, nrl2, fma2(b2, nrh2, a2));             // a - b r, r = r_hi + r_lo
            float wg0, wg1;
            unpack2(wg2, wg0, wg1);
            f32x2 ga2, gb2;
            if constexpr (RELATIVE) {
                float es0, es1;
                unpack2(fma2(b2, splat2(r_hi), splat2(1e-6f)), es0, es1);
                // MUFU.RCP alone (~1 ulp): the gate on the gradient is 1e-5 of its maximum
                const f32x2 inv2 = pack2(rcp_approx(es0), rcp_approx(es1));
                const f32x2 q2 = mul2(d2, inv2);
                float q0, q1;
                unpack2(q2, q0, q1);
                if constexpr (FUSED) {          // the statistics of the pair: weight and weight * |loss|, masked elements out
                    const f32x2 wv = pack2(fmaxf(wg0, 0.0f), fmaxf(wg1, 0.0f));
                    fs0 = add2(fs0, wv);
                    fs1 = fma2(wv, pack2(wg0 >= 0.0f ? fabsf(q0) : 0.0f, wg1 >= 0.0f ? fabsf(q1) : 0.0f), fs1);
                }
                f32x2 wu2;                                                   // weight * upstream, 0 where masked
                f32x2 extra_a = 0ull, extra_b = 0ull;
                if constexpr (ERR) {
                    // the inverse-uncertainty weight depends on the curve through a, es and max(b, 1e-6)
                    const f32x2 up2 = pack2(wg0 >= 0.0f ? up_pr : 0.0f, wg1 >= 0.0f ? up_pr : 0.0f);
                    const f32x2 ib2 = lds2(fj + 5 * kArr);
                    const f32x2 e1 = mul2(sa2, inv2);
                    const f32x2 c2 = mul2(mul2(lds2(fj + 4 * kArr), ib2), inv2);
                    const f32x2 e2 = mul2(a2, c2);
                    const f32x2 ss = fma2(e1, e1, mul2(e2, e2));             // e1^2 + e2^2
                    float T0, T1;
                    unpack2(add2(ss, splat2(1e-6f)), T0, T1);
                    const float rerr0 = rsqrt_approx(T0), rerr1 = rsqrt_approx(T1);          // 1 / err
                    const f32x2 rerr2 = pack2(rerr0, rerr1);
                    float ew0, ew1;
                    unpack2(fma2(pack2(T0, T1), rerr2, splat2(1e-6f)), ew0, ew1);            // err + 1e-6
                    const f32x2 rw2 = pack2(rcp_approx(ew0), rcp_approx(ew1));
                    wu2 = mul2(add2(wg2, rw2), up2);
                    // dm/dWt * dWt/derr * derr/dT = (l - m) U * (-rw^2) * 1/(2 err)
                    const f32x2 lm2 = add2(pack2(fabsf(q0), fabsf(q1)), splat2(-s_mean[pr]));
                    const f32x2 kk = mul2(mul2(mul2(lm2, up2), mul2(mul2(rw2, rw2), splat2(-0.5f))), rerr2);
                    float b0, b1;
                    unpack2(b2, b0, b1);
                    const f32x2 e2sq_ib = mul2(mul2(e2, e2), ib2);
                    const f32x2 dT_dbs = mul2(e2sq_ib, pack2(b0 >= 1e-6f ? -2.0f : 0.0f, b1 >= 1e-6f ? -2.0f : 0.0f));
                    const f32x2 dT_des = mul2(mul2(inv2, splat2(-2.0f)), ss);
                    extra_a = mul2(kk, mul2(mul2(e2, c2), splat2(2.0f)));
                    extra_b = mul2(kk, fma2(dT_des, splat2(r_hi), dT_dbs));
                } else {
                    wu2 = mul2(pack2(fmaxf(wg0, 0.0f), fmaxf(wg1, 0.0f)), splat2(up_pr));
                }
                // sign(q) * (wt U / es) with sign(0) = 0, as torch.abs' backward has it
                float m0, m1;
                unpack2(mul2(wu2, inv2), m0, m1);
                m0 = (q0 != 0.0f) ? m0 : 0.0f;
                m1 = (q1 != 0.0f) ? m1 : 0.0f;
                const f32x2 base2 = pack2(__int_as_float(__float_as_int(m0) ^ (__float_as_int(q0) & 0x80000000)),
                                          __int_as_float(__float_as_int(m1) ^ (__float_as_int(q1) & 0x80000000)));
                // dl/da = sgn / es;  dl/db = -sgn r (a + 1e-6) / es^2
                const f32x2 ae2 = mul2(add2(a2, splat2(1e-6f)), inv2);
                if constexpr (ERR) {
                    ga2 = add2(base2, extra_a);
                    gb2 = fma2(mul2(base2, nrh2), ae2, extra_b);
                } else {
                    ga2 = base2;
                    gb2 = mul2(mul2(base2, nrh2), ae2);
                }
            } else {
                float d0, d1;
                unpack2(d2, d0, d1);
                if constexpr (FUSED) {
                    const f32x2 wv = pack2(fmaxf(wg0, 0.0f), fmaxf(wg1, 0.0f));
                    fs0 = add2(fs0, wv);
                    fs1 = fma2(wv, pack2(wg0 >= 0.0f ? fabsf(d0) : 0.0f, wg1 >= 0.0f ? fabsf(d1) : 0.0f), fs1);
                }
                f32x2 wu2;
                if constexpr (ERR) {
                    const f32x2 rs2 = mul2(lds2(fj + 4 * kArr), splat2(r_hi));
                    float v0, v1;
                    unpack2(fma2(sa2, sa2, mul2(rs2, rs2)), v0, v1);
                    // constant with respect to the curve
                    const f32x2 rw2 = pack2(rcp_approx(sqrt_approx(v0) + 1e-6f), rcp_approx(sqrt_approx(v1) + 1e-6f));
                    wu2 = mul2(add2(wg2, rw2), pack2(wg0 >= 0.0f ? up_pr : 0.0f, wg1 >= 0.0f ? up_pr : 0.0f));
                } else {
                    wu2 = mul2(pack2(fmaxf(wg0, 0.0f), fmaxf(wg1, 0.0f)), splat2(up_pr));
                }
                float m0, m1;
                unpack2(wu2, m0, m1);
                m0 = (d0 != 0.0f) ? m0 : 0.0f;
                m1 = (d1 != 0.0f) ? m1 : 0.0f;
                ga2 = pack2(__int_as_float(__float_as_int(m0) ^ (__float_as_int(d0) & 0x80000000)),
                            __int_as_float(__float_as_int(m1) ^ (__float_as_int(d1) & 0x80000000)));
                gb2 = mul2(ga2, nrh2);
            }
            acc_i = add2(acc_i, ga2);
            sts2(fj + 3 * kArr, add2(gj_acc, gb2));
        }
        if (cur_i >= 0) sts2(const_cast<float *>(fi) + 3 * kArr, add2(lds2(fi + 3 * kArr), acc_i));

        // ---- one vectorised reduction per frame element: {G (1 - w), G w} at the element's two taps ----
        const uint32_t pre0 = u0 * lp - 0x4B000000u, pre1 = u1 * lp - 0x4B000000u;
        for (int n = 0; n < N; ++n) {
            const float *fr = slice + n * kFrameFloats;
            const f32x2 g2 = lds2(fr + 3 * kArr);
            const f32x2 xs2 = lds2(fr + 2 * kArr);
            if constexpr (MODES) {
                float ga, gb, xa, xb;
                unpack2(g2, ga, gb);
                unpack2(xs2, xa, xb);
                const bool lookup = p.mode == CLAIR_INTERP_LOOKUP;
                if (ga != 0.0f) scatter_taps_mode(copy, L, p.mode, lookup ? c : static_cast<int>(u0), xa, ga);
                if (gb != 0.0f) scatter_taps_mode(copy, L, p.mode, lookup ? c : static_cast<int>(u1), xb, gb);
                continue;
            }
            const f32x2 t2 = add2_rd(xs2, two23);
            const f32x2 w2 = sub2(xs2, sub2(t2, two23));
            float t0, t1, g0, g1, lo0, lo1, hi0, hi1;
            unpack2(t2, t0, t1);
            unpack2(g2, g0, g1);
            unpack2(mul2(g2, sub2(splat2(1.0f), w2)), lo0, lo1);
            unpack2(mul2(g2, w2), hi0, hi1);
            const uint32_t bits0 = static_cast<uint32_t>(__float_as_int(t0)), bits1 = static_cast<uint32_t>(__float_as_int(t1));
            // bits = 0x4B000000 | x0: even x0 -> A[u][x0], odd x0 -> B[u][x0 + 1] (8-byte aligned either way)
            const uint32_t off0 = (bits0 & 1u) * odd_step + bits0 + pre0;
            const uint32_t off1 = (bits1 & 1u) * odd_step + bits1 + pre1;
            red_add_v2_if(copy + off0, lo0, hi0, g0);
            red_add_v2_if(copy + off1, lo1, hi1, g1);
        }
        u0 += du;
        u0 = (u0 >= uC) ? u0 - uC : u0;
        if constexpr (FUSED) {
            if (++fused_trips == 16) {
                float a, b;
                unpack2(fs0, a, b); fd0 += static_cast<double>(a) + static_cast<double>(b);
                unpack2(fs1, a, b); fd1 += static_cast<double>(a) + static_cast<double>(b);
                fs0 = 0ull; fs1 = 0ull; fused_trips = 0;
            }
        }
    }
    if constexpr (FUSED) {
        float a, b;
        unpack2(fs0, a, b); fd0 += static_cast<double>(a) + static_cast<double>(b);
        unpack2(fs1, a, b); fd1 += static_cast<double>(a) + static_cast<double>(b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fd0 += __shfl_xor_sync(0xffffffffu, fd0, o);
            fd1 += __shfl_xor_sync(0xffffffffu, fd1, o);
        }
        if (lane == 0) {
            atomicAdd(p.sums + static_cast<int64_t>(c) * 5 + 0, fd0);
            atomicAdd(p.sums + static_cast<int64_t>(c) * 5 + 1, fd1);
        }
    }
}

constexpr int kFusedChunk = 64;
// Fused single-pair pass, second half: T[c][u][k] += sum over copies of A[u][k] + B[u][k+1] of channel block c's tables
// (float64, laid out (C, C, L) behind the (1, C, 5) sums).
__global__ void __launch_bounds__(256) fused_finalize_kernel(const float *__restrict__ hist, double *tables, int C, int L, int n_copies) {
    __shared__ double s_part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;                     // entry (u, k) of one channel block's table
    const int c = blockIdx.y;
    const int lp = L + 2;
    const bool live = i < C * L;
    const int u = live ? i / L : 0, k = live ? i - u * L : 0;
    // blockIdx.z: a chunk of kFusedChunk copies (the reduction over 341 copies by 72 blocks took 16 us, a third of what a
    // step on an eighth of the c5 image costs besides its main kernel)
    const int first = blockIdx.z * kFusedChunk, last = min(first + kFusedChunk, n_copies);
    double acc = 0.0;
    if (live) {
        const int64_t table = static_cast<int64_t>(2) * C * lp, stride = table * C;
        const float *a = hist + c * table + u * lp + k, *b = hist + c * table + C * lp + u * lp + k + 1;
        for (int r = first + warp; r < last; r += 8)
            acc += static_cast<double>(a[r * stride]) + static_cast<double>(b[r * stride]);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && live) {
#pragma unroll
        for (int w = 1; w < 8; ++w) acc += s_part[w][lane];
        double *out = tables + (static_cast<int64_t>(c) * C + u) * L + k;
        if (gridDim.z == 1) *out += acc;
        else atomicAdd(out, acc);
    }
}

// Fused single-pair pass, last step (after the all-reduce of a sharded image): mean[c] = s1 / max(s0, 1e-8), loss[c] =
// sqrt(mean^2), U[c] = mean / loss / max(s0, 1e-8) (training/icrf_training.py:133-136 and the head of the closed-form
// backward) and grad[u][k] += sum_c U[c] T[c][u][k].
__global__ void __launch_bounds__(256) fused_combine_kernel(const double *__restrict__ fused, int C, int L, double *linloss,
                                                            double *mean, double *grad) {
    __shared__ double s_u[CLAIR_MAX_CHANNELS];
    if (threadIdx.x < C) {
        const double s0 = fused[threadIdx.x * 5 + 0], s1 = fused[threadIdx.x * 5 + 1];
        const double m = s1 / fmax(s0, 1e-8);
        const double l = sqrt(m * m);
        s_u[threadIdx.x] = (l > 0.0) ? m / l / fmax(s0, 1e-8) : 0.0;
        if (blockIdx.x == 0) { mean[threadIdx.x] = m; linloss[threadIdx.x] = l; }
    }
    __syncthreads();
    const double *tables = fused + C * 5;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < C * L; e += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < C; ++c) acc += s_u[c] * tables[static_cast<int64_t>(c) * C * L + e];
        if (grad != nullptr) grad[e] += acc;
    }
}

// Scatter of an arbitrary upstream image: the table gradient of clair_icrf_forward (LINEAR).
__global__ void __launch_bounds__(256) icrf_backward_theta_kernel(const float *__restrict__ x, const float *__restrict__ gy,
                                                                  float *hist, int64_t plane, int C, int L, CurveRows rows) {
    const int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    const int slab = blockIdx.y, c = slab % C;
    const int64_t o = static_cast<int64_t>(slab) * plane + pix;
    const float g = __ldcs(gy + o);
    if (g == 0.0f) return;
    const float lm1 = static_cast<float>(L - 1);
    const float xs = fminf(fmaxf(__fmul_rn(__ldcs(x + o), lm1), 0.0f), lm1);
    const int u = static_cast<int>((pix + rows.base(c)) % C);
    float *copy = hist + static_cast<int64_t>((blockIdx.x + blockIdx.y) % kGradCopies) * (2 * C * (L + 2));
    scatter_taps(copy, C, L, u, xs, g);
}

// CATMULL table gradient: four taps per element, plain fp32 reductions into the A half of the replicated tables
__global__ void __launch_bounds__(256) icrf_catmull_backward_kernel(const float *__restrict__ x, const float *__restrict__ gy,
                                                                    float *hist, int64_t plane, int C, int L, CurveRows rows) {
    const int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    const int slab = blockIdx.y, c = slab % C;
    const int64_t o = static_cast<int64_t>(slab) * plane + pix;
    const float g = __ldcs(gy + o);
    if (g == 0.0f) return;
    const float lm1 = static_cast<float>(L - 1);
    const float xs = fminf(fmaxf(__fmul_rn(__ldcs(x + o), lm1), 0.0f), lm1);
    int x0;
    const float fl = floor_small(xs, x0);
    const float t = fminf(fmaxf(xs - fl, 0.0f), 1.0f);
    const float t2 = t * t, t3 = t2 * t;
    const float w[4] = {-0.5f * t3 + t2 - 0.5f * t, 1.5f * t3 - 2.5f * t2 + 1.0f, -1.5f * t3 + 2.0f * t2 + 0.5f * t,
                        0.5f * t3 - 0.5f * t2};
    const int u = static_cast<int>((pix + rows.base(c)) % C);
    float *copy = hist + static_cast<int64_t>((blockIdx.x + blockIdx.y) % kGradCopies) * (2 * C * (L + 2)) + u * (L + 2);
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(copy + min(max(x0 + k - 1, 0), L - 1), g * w[k]);
}

// LOOKUP table gradient (models/base.py:145-158 under autograd: an index gather): one tap, the true channel row
__global__ void __launch_bounds__(256) icrf_lookup_backward_kernel(const float *__restrict__ x, const float *__restrict__ gy,
                                                                   float *hist, int64_t plane, int C, int L) {
    const int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= plane) return;
    const int slab = blockIdx.y, c = slab % C;
    const int64_t o = static_cast<int64_t>(slab) * plane + pix;
    const float g = __ldcs(gy + o);
    if (g == 0.0f) return;
    float *copy = hist + static_cast<int64_t>((blockIdx.x + blockIdx.y) % kGradCopies) * (2 * C * (L + 2)) + c * (L + 2);
    atomicAdd(copy + icrf_lookup_index(__ldcs(x + o), static_cast<float>(L - 1)), g);
}

// grad[u][k] += sum over copies of A[u][k] + B[u][k+1], in float64.  Lanes own consecutive table entries (coalesced
// reads of a copy's row), the 8 warps of a block own interleaved copies of the block's chunk of 128, and one float64
// atomic per (entry, chunk) folds the chunks: 26 -> 5 us at 1024 copies against one warp per entry striding over the copies.
constexpr int kFinalizeChunk = 128;
__global__ void __launch_bounds__(256) grad_finalize_kernel(const float *__restrict__ hist, double *grad, int C, int L, int n_copies) {
    __shared__ double s_part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;                     // table entry
    const int lp = L + 2;
    const bool live = i < C * L;
    const int u = live ? i / L : 0, k = live ? i - u * L : 0;
    const int first = blockIdx.y * kFinalizeChunk, last = min(first + kFinalizeChunk, n_copies);
    double acc = 0.0;
    if (live) {
        const int64_t stride = static_cast<int64_t>(2) * C * lp;
        const float *a = hist + u * lp + k, *b = hist + C * lp + u * lp + k + 1;
        for (int r = first + warp; r < last; r += 8)
            acc += static_cast<double>(a[r * stride]) + static_cast<double>(b[r * stride]);
    }
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && live) {
#pragma unroll
        for (int w = 1; w < 8; ++w) acc += s_part[w][lane];
        if (gridDim.y == 1) grad[i] += acc;
        else atomicAdd(grad + i, acc);
    }
}

// =====================================================================================================
// Closing a training step on the device (training/icrf_training.py:133-146): (P, C)-sized
// algebra and the four curve penalties with their gradients.  One block each; they exist because the same
// arithmetic as ~100 separate 3-microsecond ATen launches dominated the step time.
// =====================================================================================================
// mean[p,c] = s1/max(s0,1e-8); linloss[c] = sqrt(sum_p mean^2); upstream[p,c] = mean/linloss/max(s0,1e-8) (0 if linloss = 0);
// mean_for_grad = mean, or 0 where the denominator was clamped (no d/dWt through a clamped denominator)
__global__ void __launch_bounds__(256) pair_upstream_kernel(const double *__restrict__ sums, int P, int C, double *linloss,
                                                            double *mean, double *upstream, double *mean_for_grad) {
    __shared__ double s_lin[CLAIR_MAX_CHANNELS];
    const int total = P * C;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const double s0 = sums[e * 5 + 0], s1 = sums[e * 5 + 1];
        mean[e] = s1 / fmax(s0, 1e-8);                            // common/general_functions.py:156-160
    }
    __syncthreads();
    if (threadIdx.x < C) {
        double acc = 0.0;
        for (int k = 0; k < P; ++k) {
            const double m = mean[k * C + threadIdx.x];
            acc += m * m;
        }
        const double l = sqrt(acc);                               // training/icrf_training.py:136
        s_lin[threadIdx.x] = l;
        linloss[threadIdx.x] = l;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const double s0 = sums[e * 5 + 0];
        const double l = s_lin[e % C];
        const double m = mean[e];
        upstream[e] = (l > 0.0) ? m / l / fmax(s0, 1e-8) : 0.0;
        mean_for_grad[e] = (s0 < 1e-8) ? 0.0 : m;
    }
}

// pen[c] = alpha*monotonicity + beta*range + gamma*endpoints + delta*smoothness (training/losses.py:111-190, per channel)
// and grad[c][k] += d pen[c] / d theta[c][k].  One block per channel, fp32 arithmetic like the reference's.
__global__ void __launch_bounds__(256) curve_penalty_kernel(const float *__restrict__ theta, int C, int L, float alpha, float beta,
                                                            float gamma, float delta, double *pen, double *grad) {
    __shared__ float s_red[256];
    const int c = blockIdx.x;
    const float *th = theta + static_cast<int64_t>(c) * L;
    float local = 0.0f;
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        const float t = th[k];
        float g = 0.0f;
        // monotonicity: sum [df <= 0] df^2 over df = th[k+1] - th[k]
        if (k + 1 < L) {
            const float df = th[k + 1] - t;
            if (df <= 0.0f) { local += alpha * df * df; g -= alpha * 2.0f * df; }
        }
        if (k >= 1) {
            const float df = t - th[k - 1];
            if (df <= 0.0f) g += alpha * 2.0f * df;
        }
        // range: relu(-t) + relu(t - 1)
        local += beta * (fmaxf(-t, 0.0f) + fmaxf(t - 1.0f, 0.0f));
        g += beta * ((t > 1.0f ? 1.0f : 0.0f) - (t < 0.0f ? 1.0f : 0.0f));
        // endpoints: th[0]^2 + (th[L-1] - 1)^2
        if (k == 0) { local += gamma * t * t; g += gamma * 2.0f * t; }
        if (k == L - 1) { local += gamma * (t - 1.0f) * (t - 1.0f); g += gamma * 2.0f * (t - 1.0f); }
        // smoothness: sum (th[k] - 2 th[k+1] + th[k+2])^2; entry k is the left, middle and right tap of three terms
        if (k + 2 < L) {
            const float sd = t - 2.0f * th[k + 1] + th[k + 2];
            local += delta * sd * sd;
            g += delta * 2.0f * sd;
        }
        if (k >= 1 && k + 1 < L) g -= delta * 4.0f * (th[k - 1] - 2.0f * t + th[k + 1]);
        if (k >= 2) g += delta * 2.0f * (th[k - 2] - 2.0f * th[k - 1] + t);
        grad[static_cast<int64_t>(c) * L + k] += static_cast<double>(g);
    }
    s_red[threadIdx.x] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) pen[c] = static_cast<double>(s_red[0]);
}

}  // namespace clair

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace clair;

namespace {

int fill_pairs(const char *fn, PairTable &t, const int32_t *pi, const int32_t *pj, const double *pr, int first, int count,
               int n_frames) {
    for (int k = 0; k < count; ++k) {
        const int32_t i = pi[first + k], j = pj[first + k];
        if (i < 0 || j < 0 || i >= n_frames || j >= n_frames) {
            char buf[128];
            std::snprintf(buf, sizeof(buf), "%s: pair index out of range", fn);
            return fail(CLAIR_E_ARG, buf);
        }
        const double r = pr[first + k];
        const float hi = static_cast<float>(r);
        t.i[k] = static_cast<uint8_t>(i);
        t.j[k] = static_cast<uint8_t>(j);
        t.r_hi[k] = hi;
        t.r_lo[k] = static_cast<float>(r - static_cast<double>(hi));
    }
    return 0;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    return 0;
}

// Choose warps per block W and register slots S (W*S >= pairs in the launch) from an instruction-count model of one
// tile: every warp stages ceil(N*items/ (32 W)) items of VA pixels (cost_item each) and then walks S pair slots of
// kStatsTile/32 pixels each (cost_pair per pixel).  Block time per tile ~ that sum; SM throughput ~ 1 / (W * sum).
// The packed kernel (va == 4) keeps its staging items in registers, at most two per thread: W >= ceil(N / 2).
void pick_stats_shape(int count, int n_frames, int va, bool err, bool full, int tile, int &warps, int &slots) {
    const bool packed = va == 4;
    const int cost_item = packed ? (err ? 150 : 100) : va * (err ? 50 : 40) + 12;
    const int cost_pair = packed ? (full ? (err ? 130 : 75) : (err ? 100 : 55))
                                 : (tile / 32) * ((full ? 30 : 21) + (err ? 16 : 0));
    const int items = n_frames * (tile / va);
    long best = -1;
    warps = 8; slots = kMaxSlots;
    for (int s = 1; s <= kMaxSlots; ++s) {
        if (s == 3) continue;                         // slot counts instantiated: 1, 2, 4
        for (int w = packed ? 1 : 2; w <= 16; ++w) {    // a single exposure pair (N = 2, one pair) runs as one-warp blocks
            if (w * s < count) continue;
            if (packed && 2 * w < n_frames * (tile / 128)) continue;
            const int trips = (items + 32 * w - 1) / (32 * w);
            const long cost = static_cast<long>(w) * (static_cast<long>(trips) * cost_item + static_cast<long>(s) * cost_pair);
            if (best < 0 || cost < best) { best = cost; warps = w; slots = s; }
        }
    }
}

void common_params(PairParams &p, const float *val, const float *std, const float *theta, int n_frames, int n_channels,
                   int64_t plane, int lut, const int32_t *row_base_host, float lo, float hi, int unc, int mode) {
    p.val = val; p.std = std; p.theta = theta;
    p.mode = theta ? mode : CLAIR_INTERP_LINEAR;
    p.plane = plane; p.n_frames = n_frames; p.n_channels = n_channels; p.lut = lut;
    p.unc_weighting = unc; p.valid_lo = lo; p.valid_hi = hi;
    p.mod_magic = (65536u + n_channels - 1) / n_channels;
    fill_rows(p.rows, row_base_host, n_channels, plane);
}

}  // namespace

// the mode argument shared by the pair entry points: LINEAR, or LOOKUP / CATMULL with the reference's restriction that a
// LOOKUP model has no derivative to carry std images through (measure_linearity.py:57-63 raises there)
static int check_pair_mode(const char *fn, int interp_mode, const float *theta_dev, const float *std_dev) {
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_LOOKUP && interp_mode != CLAIR_INTERP_CATMULL) {
        char buf[160];
        std::snprintf(buf, sizeof(buf), "%s: interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL", fn);
        return fail(CLAIR_E_MODE, buf);
    }
    if (interp_mode == CLAIR_INTERP_LOOKUP && theta_dev && std_dev) {
        char buf[160];
        std::snprintf(buf, sizeof(buf), "%s: a LOOKUP model has no derivative to propagate std images through", fn);
        return fail(CLAIR_E_MODE, buf);
    }
    return 0;
}

static int pair_stats_impl(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                           const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host, int n_pairs,
                           const float *theta_dev, int lut_size, int interp_mode, const int32_t *curve_row_base_host, float valid_lo,
                           float valid_hi, int relative, int unc_weighting, int full, double *sums_dev, void *stream) {
    const char *fn = full ? "clair_pair_stats" : "clair_pair_means";
    NvtxRange nvtx_range_("clair_pair_stats");
    if (int rc = check_pair_mode(fn, interp_mode, theta_dev, std_dev)) return rc;
    const bool modes = theta_dev != nullptr && interp_mode != CLAIR_INTERP_LINEAR;
    if (!val_dev || !sums_dev) return fail(CLAIR_E_ARG, "clair_pair_stats: null buffer");
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_i_host || !pair_j_host || !pair_ratio_host)))
        return fail(CLAIR_E_ARG, "clair_pair_stats: pair table missing");
    if (theta_dev == nullptr && lut_size <= 0) lut_size = 2;
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, true)) return rc;
    if (n_pairs > CLAIR_MAX_PAIRS) return fail(CLAIR_E_LIMIT, "clair_pair_stats: more than CLAIR_MAX_PAIRS pairs");
    if (n_pairs == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the uncertainty term is only evaluated (and the std images only read) when something uses it
    const bool err = std_dev != nullptr && (full || unc_weighting);
    const int arrays = 2 + (err ? (relative ? 2 : 1) : 0);
    const size_t table_bytes = theta_dev ? sizeof(float2) * n_channels * lut_size : 0;
    const int per_launch = 16 * kMaxSlots;   // 64 pairs: 16 warps x 4 register slots
    // 128-bit staging loads need H*W % 4 == 0 and 16-byte aligned stacks (every frame / channel slab then is, too)
    const bool vec_ok = plane % 4 == 0 && plane < (1ll << 31) && reinterpret_cast<uintptr_t>(val_dev) % 16 == 0 &&
                        (!err || reinterpret_cast<uintptr_t>(std_dev) % 16 == 0) &&
                        (theta_dev == nullptr || (n_channels * lut_size) % 2 == 0);   // tile starts 16-byte aligned after the table
    // the packed training variant works on 256-pixel tiles when its staging items still fit two per thread
    // (a single exposure pair is all staging and no pair work: c5 means pass 0.89 ms on 128-pixel tiles, 1.04 ms on 256)
    const int tile = (!full && vec_ok && !modes && n_frames * 2 <= 32 && n_pairs >= 4) ? kMeansTile : kStatsTile;
    const size_t tile_bytes = sizeof(float) * arrays * n_frames * tile;
    size_t smem = table_bytes + tile_bytes;
    const int va = (vec_ok && n_frames * (tile / 128) <= 32) ? 4 : 1;     // the packed kernel holds <= 2 staging items per thread, 16 warps
    // packed kernel: a second tile buffer lets staging overlap the pair phase (one barrier per tile) as long as two
    // blocks still fit an SM
    int buffers = (va == 4 && table_bytes + 2 * tile_bytes <= 110 * 1024) ? 2 : 1;
    if (va == 4 && (g_tuning.stats_buffers == 1 || g_tuning.stats_buffers == 2)) buffers = g_tuning.stats_buffers;
    if (va == 4 && buffers == 2) smem = table_bytes + 2 * tile_bytes;
    const int n_launches = (n_pairs + per_launch - 1) / per_launch;
    int first = 0;
    for (int l = 0; l < n_launches; ++l) {
        const int count = (n_pairs - first + (n_launches - l) - 1) / (n_launches - l);   // balanced chunks
        PairParams p{};
        common_params(p, val_dev, err ? std_dev : nullptr, theta_dev, n_frames, n_channels, plane, lut_size, curve_row_base_host,
                      valid_lo, valid_hi, unc_weighting, interp_mode);
        p.sums = sums_dev + static_cast<int64_t>(first) * n_channels * 5;
        p.n_pairs = count;
        p.stats_buffers = buffers;
        if (int rc = fill_pairs(fn, p.pairs, pair_i_host, pair_j_host, pair_ratio_host, first, count, n_frames)) return rc;
        int warps, slots;
        pick_stats_shape(count, n_frames, va, err, full != 0, va == 4 ? tile : kStatsTile, warps, slots);
        if (g_tuning.stats_warps > 0 && g_tuning.stats_slots > 0 && g_tuning.stats_warps * g_tuning.stats_slots >= count &&
            g_tuning.stats_warps <= 16 && 2 * g_tuning.stats_warps >= n_frames * (tile / 128) &&
            (g_tuning.stats_slots == 1 || g_tuning.stats_slots == 2 || g_tuning.stats_slots == 4)) {
            warps = g_tuning.stats_warps;
            slots = g_tuning.stats_slots;
        }
        if (modes && va == 4) {
            // LOOKUP / CATMULL models: one shape (4 register slots, two staging items per thread) keeps the number of
            // instantiations down; the pair phase is the LINEAR kernels' own
            slots = kMaxSlots;
            warps = std::max((count + kMaxSlots - 1) / kMaxSlots, (n_frames * (tile / 128) + 1) / 2);
        }
        const int trips = (n_frames * (tile / 128) + warps - 1) / warps;      // staging items per thread of the packed kernel (1 or 2)
        const int64_t n_tiles = (plane + (va == 4 ? tile : kStatsTile) - 1) / (va == 4 ? tile : kStatsTile);
        auto launch = [&](auto kernel) -> int {
            if (int rc = set_smem(kernel, smem)) return rc;
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem);
            per_sm = g_tuning.stats_blocks_per_sm > 0 ? g_tuning.stats_blocks_per_sm : std::max(per_sm, 1);
            // Several pairs per tile: 4 x the resident blocks (c3 1.656 -> 1.55 ms, c2 0.271 -> 0.261 ms: blocks that start late
            // or sit on a slower memory partition no longer hold the launch back); a single exposure pair is all staging, where
            // the per-block table staging and final atomics weigh more (c5/8: 0.194 ms at 1 x, 0.222 ms at 4 x).
            const int waves = g_tuning.stats_waves > 0 ? g_tuning.stats_waves : (count >= 4 ? 4 : 1);
            if (n_tiles >= 8 * waves * resident_blocks_per_channel(per_sm, n_channels)) per_sm *= waves;
            int64_t gx = std::min<int64_t>(n_tiles, std::max<int64_t>(1, resident_blocks_per_channel(per_sm, n_channels)));
            // a grid stride that is a multiple of C pixels keeps every staging item on the same table rows in all its tiles
            if (gx > n_channels) gx -= gx % n_channels;
            kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels)), warps * 32, smem, s>>>(p);
            return 0;
        };
        int rc = 0;
#define STATS_TRIPS(S, E, R, F, M, T) \
    (trips == 1 ? launch(pair_stats2_kernel<S, E, R, F, 1, M, T>) : launch(pair_stats2_kernel<S, E, R, F, 2, M, T>))
#define STATS_FLAGS(S, E, R, F)                                                                        \
    (va == 4 ? (modes ? launch(pair_stats2_kernel<kMaxSlots, E, R, F, 2, true, kStatsTile>)               \
                      : ((!F && tile == kMeansTile) ? STATS_TRIPS(S, E, R, F, false, kMeansTile)           \
                                                    : STATS_TRIPS(S, E, R, F, false, kStatsTile)))         \
             : launch(pair_stats_kernel<S, E, R, F, 1>))
#define STATS_CASE(S)                                                                                   \
    case S:                                                                                             \
        if (err) {                                                                                      \
            if (relative) rc = full ? STATS_FLAGS(S, true, true, true) : STATS_FLAGS(S, true, true, false);     \
            else rc = full ? STATS_FLAGS(S, true, false, true) : STATS_FLAGS(S, true, false, false);            \
        } else {                                                                                        \
            if (relative) rc = full ? STATS_FLAGS(S, false, true, true) : STATS_FLAGS(S, false, true, false);   \
            else rc = full ? STATS_FLAGS(S, false, false, true) : STATS_FLAGS(S, false, false, false);          \
        }                                                                                               \
        break;
        switch (slots) {
            STATS_CASE(1)
            STATS_CASE(2)
            default:
            STATS_CASE(4)
        }
#undef STATS_FLAGS
#undef STATS_TRIPS
#undef STATS_CASE
        if (rc) return rc;
        if (int rc2 = launched("pair_stats_kernel")) return rc2;
        first += count;
    }
    return 0;
}

extern "C" int clair_pair_stats(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                                const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                                int n_pairs, const float *theta_dev, int lut_size, int interp_mode,
                                const int32_t *curve_row_base_host,
                                float valid_lo, float valid_hi, int relative, int unc_weighting, double *sums_dev,
                                void *stream) {
    return pair_stats_impl(val_dev, std_dev, n_frames, n_channels, plane, pair_i_host, pair_j_host, pair_ratio_host, n_pairs,
                           theta_dev, lut_size, interp_mode, curve_row_base_host, valid_lo, valid_hi, relative, unc_weighting, 1, sums_dev,
                           stream);
}

extern "C" int clair_pair_means(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                                const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                                int n_pairs, const float *theta_dev, int lut_size, int interp_mode,
                                const int32_t *curve_row_base_host,
                                float valid_lo, float valid_hi, int relative, int unc_weighting, double *sums_dev,
                                void *stream) {
    return pair_stats_impl(val_dev, std_dev, n_frames, n_channels, plane, pair_i_host, pair_j_host, pair_ratio_host, n_pairs,
                           theta_dev, lut_size, interp_mode, curve_row_base_host, valid_lo, valid_hi, relative, unc_weighting, 0, sums_dev,
                           stream);
}

extern "C" size_t clair_grad_workspace_bytes(int n_channels, int lut_size) {
    if (n_channels <= 0 || lut_size <= 0) return 0;
    return sizeof(float) * kMaxGradCopies * 2 * static_cast<size_t>(n_channels) * (lut_size + 2);
}

namespace {
int finalize_grad(const float *hist, double *grad, int C, int L, int n_copies, cudaStream_t s) {
    const int n = C * L;
    const dim3 grid(static_cast<unsigned>((n + 31) / 32), static_cast<unsigned>((n_copies + kFinalizeChunk - 1) / kFinalizeChunk));
    grad_finalize_kernel<<<grid, 256, 0, s>>>(hist, grad, C, L, n_copies);
    return launched("grad_finalize_kernel");
}
}  // namespace

extern "C" int clair_pair_grad(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                               const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                               int n_pairs, const float *theta_dev, int lut_size, int interp_mode,
                               const int32_t *curve_row_base_host,
                               float valid_lo, float valid_hi, int relative, int unc_weighting,
                               const double *upstream_dev, const double *mean_dev, double *grad_theta_dev,
                               void *workspace_dev, size_t workspace_bytes, void *stream) {
    const char *fn = "clair_pair_grad";
    NvtxRange nvtx_range_("clair_pair_grad");
    if (!val_dev || !theta_dev || !upstream_dev || !mean_dev || !grad_theta_dev || !workspace_dev)
        return fail(CLAIR_E_ARG, "clair_pair_grad: null buffer");
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_i_host || !pair_j_host || !pair_ratio_host)))
        return fail(CLAIR_E_ARG, "clair_pair_grad: pair table missing");
    if (int rc = check_pair_mode(fn, interp_mode, theta_dev, std_dev)) return rc;
    const bool modes = interp_mode != CLAIR_INTERP_LINEAR;
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, true)) return rc;
    if (n_pairs > CLAIR_MAX_PAIRS) return fail(CLAIR_E_LIMIT, "clair_pair_grad: more than CLAIR_MAX_PAIRS pairs");
    const size_t need = clair_grad_workspace_bytes(n_channels, lut_size);
    if (workspace_bytes < need || reinterpret_cast<uintptr_t>(workspace_dev) % 16 != 0)
        return fail(CLAIR_E_ARG, "clair_pair_grad: workspace too small or misaligned (see clair_grad_workspace_bytes)");
    if (n_pairs == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_copies = g_tuning.grad_copies > 0 ? std::min(g_tuning.grad_copies, kMaxGradCopies) : kPairGradCopies;
    if (cudaError_t e = cudaMemsetAsync(workspace_dev, 0, need / kMaxGradCopies * n_copies, s); e != cudaSuccess)
        return fail_cuda(e, "cudaMemsetAsync(workspace)");
    const bool err = std_dev != nullptr && unc_weighting;
    int first = 0;
    while (first < n_pairs) {
        const int count = std::min(kMaxPairsPerLaunch, n_pairs - first);
        const bool pair_ok = plane % 2 == 0 && reinterpret_cast<uintptr_t>(val_dev) % 8 == 0 &&
                             (!err || reinterpret_cast<uintptr_t>(std_dev) % 8 == 0);
        const int pixn = (pair_ok && g_tuning.grad_pix != 1) ? 2 : 1;   // pixels per lane
        const size_t fixed_bytes = sizeof(float2) * n_channels * lut_size + sizeof(float) * (2 * count + 2);
        const size_t per_warp = sizeof(float) * (err ? 6 : 4) * n_frames * 32 * pixn;
        // 4 warps per block: more, smaller blocks balance better across the SMs than 8-warp blocks
        const size_t want_warps = g_tuning.grad_warps > 0 ? static_cast<size_t>(g_tuning.grad_warps) : 4;
        const int warps = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want_warps, (200 * 1024 - fixed_bytes) / per_warp)));
        const size_t smem = fixed_bytes + per_warp * warps;
        PairParams p{};
        common_params(p, val_dev, err ? std_dev : nullptr, theta_dev, n_frames, n_channels, plane, lut_size, curve_row_base_host,
                      valid_lo, valid_hi, unc_weighting, interp_mode);
        p.upstream = upstream_dev + static_cast<int64_t>(first) * n_channels;
        p.mean = mean_dev + static_cast<int64_t>(first) * n_channels;
        p.hist = static_cast<float *>(workspace_dev);
        p.n_copies = n_copies;
        p.n_pairs = count;
        if (int rc = fill_pairs(fn, p.pairs, pair_i_host, pair_j_host, pair_ratio_host, first, count, n_frames)) return rc;
        const int64_t n_groups = (plane + 32 * pixn - 1) / (32 * pixn);
        auto launch = [&](auto kernel) -> int {
            if (int rc = set_smem(kernel, smem)) return rc;
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem);
            // The kernel is bound by L2 reductions and runs best at 6 blocks of 4 warps per SM, not at the 8 that fit (two boxes,
        // blocks per SM 4 / 5 / 6 / 7 / 8 / 12: c5 2.95 / 2.62 / 2.41 / 2.65 / 2.76 / 3.09 ms; half / a quarter of c5 1.50 -> 1.21,
        // 0.71 -> 0.62 ms; 8K 0.91 -> 0.79 ms; an eighth of c5 is flat from 6 to 8), and on exactly that many blocks (2 x: 3.1 ms).
        per_sm = g_tuning.grad_blocks_per_sm > 0 ? g_tuning.grad_blocks_per_sm : std::min(std::max(per_sm, 1), 6);
            // one or two pairs (little work per trip): 4 x the resident blocks (c5 over three boxes: 2.39 / 2.59 / 2.73 ms at 1 x,
            // 2.37 / 2.26 / 2.32 ms at 4 x — the L2-reduction-bound kernel is erratic at 1 x .. 3 x; an eighth of c5 0.41 -> 0.32 ms);
            // many pairs are fastest on exactly the resident blocks (c2 0.340 ms, 0.349 ms at 4 x)
            per_sm *= g_tuning.grad_waves > 0 ? g_tuning.grad_waves : (count <= 2 ? 4 : 1);
            const int64_t gx = std::min<int64_t>((n_groups + warps - 1) / warps,
                                                 std::max<int64_t>(1, resident_blocks_per_channel(per_sm, n_channels)));
            kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels)), warps * 32, smem, s>>>(p);
            return 0;
        };
        int rc;
#define GRAD_FLAGS(E, R) \
    (pixn == 2 ? (modes ? launch(pair_grad2_kernel<E, R, true>) : launch(pair_grad2_kernel<E, R>)) : launch(pair_grad_kernel<E, R, 1>))
        if (err) rc = relative ? GRAD_FLAGS(true, true) : GRAD_FLAGS(true, false);
        else rc = relative ? GRAD_FLAGS(false, true) : GRAD_FLAGS(false, false);
#undef GRAD_FLAGS
        if (rc) return rc;
        if (int rc2 = launched("pair_grad_kernel")) return rc2;
        first += count;
    }
    return finalize_grad(static_cast<const float *>(workspace_dev), grad_theta_dev, n_channels, lut_size, n_copies, s);
}

extern "C" size_t clair_pair_fused_doubles(int n_channels, int lut_size) {
    if (n_channels <= 0 || lut_size <= 0) return 0;
    return static_cast<size_t>(n_channels) * 5 + static_cast<size_t>(n_channels) * n_channels * lut_size;
}

extern "C" int clair_pair_fused(const float *val_dev, int n_frames, int n_channels, int64_t plane, const int32_t *pair_i_host,
                                const int32_t *pair_j_host, const double *pair_ratio_host, int n_pairs, const float *theta_dev,
                                int lut_size, int interp_mode, const int32_t *curve_row_base_host, float valid_lo, float valid_hi,
                                int relative, double *fused_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
    const char *fn = "clair_pair_fused";
    NvtxRange nvtx_range_(fn);
    if (!val_dev || !theta_dev || !fused_dev || !workspace_dev || !pair_i_host || !pair_j_host || !pair_ratio_host)
        return fail(CLAIR_E_ARG, "clair_pair_fused: null buffer");
    if (n_pairs != 1) return fail(CLAIR_E_MODE, "clair_pair_fused: the single-pass form is defined for exactly one exposure pair");
    if (int rc = check_pair_mode(fn, interp_mode, theta_dev, nullptr)) return rc;
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, true)) return rc;
    if (plane % 2 != 0 || reinterpret_cast<uintptr_t>(val_dev) % 8 != 0)
        return fail(CLAIR_E_ARG, "clair_pair_fused: H*W must be even and the stack 8-byte aligned (use the two-pass entry points otherwise)");
    const size_t need = clair_grad_workspace_bytes(n_channels, lut_size);
    if (workspace_bytes < need || reinterpret_cast<uintptr_t>(workspace_dev) % 16 != 0)
        return fail(CLAIR_E_ARG, "clair_pair_fused: workspace too small or misaligned (see clair_grad_workspace_bytes)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n_copies = std::max(1, (g_tuning.grad_copies > 0 ? std::min(g_tuning.grad_copies, kMaxGradCopies) : kPairGradCopies) / n_channels);
    const size_t table_bytes = need / kMaxGradCopies;
    if (cudaError_t e = cudaMemsetAsync(workspace_dev, 0, table_bytes * n_copies * n_channels, s); e != cudaSuccess)
        return fail_cuda(e, "cudaMemsetAsync(workspace)");
    const bool modes = interp_mode != CLAIR_INTERP_LINEAR;
    const size_t fixed_bytes = sizeof(float2) * n_channels * lut_size + sizeof(float) * (2 * 1 + 2);
    const size_t per_warp = sizeof(float) * 4 * n_frames * 64;
    const size_t want_warps = g_tuning.grad_warps > 0 ? static_cast<size_t>(g_tuning.grad_warps) : 4;
    const int warps = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want_warps, (200 * 1024 - fixed_bytes) / per_warp)));
    const size_t smem = fixed_bytes + per_warp * warps;
    PairParams p{};
    common_params(p, val_dev, nullptr, theta_dev, n_frames, n_channels, plane, lut_size, curve_row_base_host, valid_lo, valid_hi, 0,
                  interp_mode);
    p.sums = fused_dev;
    p.hist = static_cast<float *>(workspace_dev);
    p.n_copies = n_copies;
    p.n_pairs = 1;
    if (int rc = fill_pairs(fn, p.pairs, pair_i_host, pair_j_host, pair_ratio_host, 0, 1, n_frames)) return rc;
    const int64_t n_groups = (plane + 63) / 64;
    auto launch = [&](auto kernel) -> int {
        if (int rc = set_smem(kernel, smem)) return rc;
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem);
        // The kernel is bound by L2 reductions and runs best at 6 blocks of 4 warps per SM, not at the 8 that fit (two boxes,
        // blocks per SM 4 / 5 / 6 / 7 / 8 / 12: c5 2.95 / 2.62 / 2.41 / 2.65 / 2.76 / 3.09 ms; half / a quarter of c5 1.50 -> 1.21,
        // 0.71 -> 0.62 ms; 8K 0.91 -> 0.79 ms; an eighth of c5 is flat from 6 to 8), and on exactly that many blocks (2 x: 3.1 ms).
        per_sm = g_tuning.grad_blocks_per_sm > 0 ? g_tuning.grad_blocks_per_sm : std::min(std::max(per_sm, 1), 6);
        const int64_t gx = std::min<int64_t>((n_groups + warps - 1) / warps, std::max<int64_t>(1, resident_blocks_per_channel(per_sm, n_channels)));
        kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels)), warps * 32, smem, s>>>(p);
        return 0;
    };
    int rc;
    if (relative) rc = modes ? launch(pair_grad2_kernel<false, true, true, true>) : launch(pair_grad2_kernel<false, true, false, true>);
    else rc = modes ? launch(pair_grad2_kernel<false, false, true, true>) : launch(pair_grad2_kernel<false, false, false, true>);
    if (rc) return rc;
    if (int rc2 = launched("pair_grad2_kernel<fused>")) return rc2;
    const dim3 grid(static_cast<unsigned>((n_channels * lut_size + 31) / 32), static_cast<unsigned>(n_channels),
                    static_cast<unsigned>((n_copies + kFusedChunk - 1) / kFusedChunk));
    fused_finalize_kernel<<<grid, 256, 0, s>>>(static_cast<const float *>(workspace_dev), fused_dev + n_channels * 5, n_channels, lut_size,
                                               n_copies);
    return launched("fused_finalize_kernel");
}

extern "C" int clair_pair_fused_combine(const double *fused_dev, int n_channels, int lut_size, double *linloss_dev, double *mean_dev,
                                        double *grad_theta_dev, void *stream) {
    if (!fused_dev || !linloss_dev || !mean_dev) return fail(CLAIR_E_ARG, "clair_pair_fused_combine: null buffer");
    if (n_channels <= 0 || n_channels > CLAIR_MAX_CHANNELS || lut_size < 2 || lut_size > CLAIR_MAX_LUT)
        return fail(CLAIR_E_LIMIT, "clair_pair_fused_combine: bad table shape");
    const int blocks = std::max(1, (n_channels * lut_size + 255) / 256);
    fused_combine_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(fused_dev, n_channels, lut_size, linloss_dev, mean_dev,
                                                                                grad_theta_dev);
    return launched("fused_combine_kernel");
}

extern "C" int clair_icrf_backward_theta(const float *x_dev, const float *grad_y_dev, double *grad_theta_dev,
                                         int n_frames, int n_channels, int64_t plane, int lut_size, int interp_mode,
                                         const int32_t *curve_row_base_host, void *workspace_dev, size_t workspace_bytes,
                                         void *stream) {
    if (!x_dev || !grad_y_dev || !grad_theta_dev || !workspace_dev) return fail(CLAIR_E_ARG, "clair_icrf_backward_theta: null buffer");
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_CATMULL && interp_mode != CLAIR_INTERP_LOOKUP)
        return fail(CLAIR_E_MODE, "clair_icrf_backward_theta: interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL");
    if (int rc = check_geometry("clair_icrf_backward_theta", n_frames, n_channels, plane, lut_size, false)) return rc;
    const size_t need = clair_grad_workspace_bytes(n_channels, lut_size);
    if (workspace_bytes < need || reinterpret_cast<uintptr_t>(workspace_dev) % 16 != 0)
        return fail(CLAIR_E_ARG, "clair_icrf_backward_theta: workspace too small or misaligned");
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_icrf_backward_theta: n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaError_t e = cudaMemsetAsync(workspace_dev, 0, need / kMaxGradCopies * kGradCopies, s); e != cudaSuccess)
        return fail_cuda(e, "cudaMemsetAsync(workspace)");
    CurveRows rows;
    fill_rows(rows, curve_row_base_host, n_channels, plane);
    dim3 grid(static_cast<unsigned>((plane + 255) / 256), static_cast<unsigned>(slabs));
    if (interp_mode == CLAIR_INTERP_CATMULL)
        icrf_catmull_backward_kernel<<<grid, 256, 0, s>>>(x_dev, grad_y_dev, static_cast<float *>(workspace_dev), plane,
                                                          n_channels, lut_size, rows);
    else if (interp_mode == CLAIR_INTERP_LOOKUP)
        icrf_lookup_backward_kernel<<<grid, 256, 0, s>>>(x_dev, grad_y_dev, static_cast<float *>(workspace_dev), plane,
                                                         n_channels, lut_size);
    else
        icrf_backward_theta_kernel<<<grid, 256, 0, s>>>(x_dev, grad_y_dev, static_cast<float *>(workspace_dev), plane, n_channels,
                                                        lut_size, rows);
    if (int rc = launched("icrf_backward_theta_kernel")) return rc;
    return finalize_grad(static_cast<const float *>(workspace_dev), grad_theta_dev, n_channels, lut_size, kGradCopies, s);
}

extern "C" int clair_pair_upstream(const double *sums_dev, int n_pairs, int n_channels, double *linloss_dev, double *mean_dev,
                                   double *upstream_dev, double *mean_for_grad_dev, void *stream) {
    if (!sums_dev || !linloss_dev || !mean_dev || !upstream_dev || !mean_for_grad_dev)
        return fail(CLAIR_E_ARG, "clair_pair_upstream: null buffer");
    if (n_pairs <= 0 || n_channels <= 0 || n_channels > CLAIR_MAX_CHANNELS || n_pairs > CLAIR_MAX_PAIRS)
        return fail(CLAIR_E_LIMIT, "clair_pair_upstream: bad pair / channel count");
    pair_upstream_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums_dev, n_pairs, n_channels, linloss_dev, mean_dev,
                                                                           upstream_dev, mean_for_grad_dev);
    return launched("pair_upstream_kernel");
}

extern "C" int clair_curve_penalties(const float *theta_dev, int n_channels, int lut_size, float alpha, float beta, float gamma,
                                     float delta, double *penalty_dev, double *grad_theta_dev, void *stream) {
    if (!theta_dev || !penalty_dev || !grad_theta_dev) return fail(CLAIR_E_ARG, "clair_curve_penalties: null buffer");
    if (n_channels <= 0 || n_channels > CLAIR_MAX_CHANNELS || lut_size < 3)
        return fail(CLAIR_E_LIMIT, "clair_curve_penalties: bad table shape");
    curve_penalty_kernel<<<n_channels, 256, 0, static_cast<cudaStream_t>(stream)>>>(theta_dev, n_channels, lut_size, alpha, beta,
                                                                                    gamma, delta, penalty_dev, grad_theta_dev);
    return launched("curve_penalty_kernel");
}
