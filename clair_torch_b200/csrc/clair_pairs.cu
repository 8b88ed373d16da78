// Pairwise exposure-ratio kernels: linearity statistics (forward) and the ICRF-table gradient (backward).
//
// Both are FP32-issue-bound once there are more than a handful of exposure pairs (P grows as N^2/2), not
// HBM-bound: every frame element is read from HBM once per launch and then reused from shared memory for
// all pairs it takes part in.  See DESIGN.md §4-§5.
#include "clair_common.cuh"
#include "clair_host.h"

#include <cstdio>

namespace clair {

constexpr float kPairNegScaleLog2e = -14.426950408889634f;   // -10 * log2(e); training/losses.py:212 default scale 10
constexpr int kMaxPairsPerLaunch = 256;   // the pair table travels as a kernel argument
constexpr int kStatsTile = 128;           // pixels per shared-memory tile in the statistics kernel
constexpr int kMaxSlots = 4;              // pairs a warp carries in registers in the statistics kernel
constexpr int kGradCopies = 64;           // replicated gradient tables the REDs are spread over

struct PairTable {
    float r_hi[kMaxPairsPerLaunch];       // exposure ratio t_i/t_j split into two floats (r = hi + lo to ~48 bits)
    float r_lo[kMaxPairsPerLaunch];
    uint8_t i[kMaxPairsPerLaunch];
    uint8_t j[kMaxPairsPerLaunch];
};

struct PairParams {
    const float *val;
    const float *std;
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
    float valid_lo, valid_hi;
    CurveRows rows;
    PairTable pairs;
};

// Per-frame quantities shared by all pairs a frame element takes part in.
//   f     linearised value                                     (models/base.py:182)
//   sig   |f'(x) * std|                                         (training/icrf_training.py:124)
//   gw    exp(-10 (x-.5)^2) of the RAW value, or -1 when the raw value is outside [valid_lo, valid_hi]
//         (training/losses.py:229-234 and common/general_functions.py:305 folded into one number)
struct FrameTerms {
    float f, sig, gw, xs;
};

__device__ __forceinline__ FrameTerms frame_terms(float x, float s, bool has_model, const float2 *row, float lm1,
                                                  float lo, float hi, bool has_std) {
    FrameTerms t;
    float fp = 1.0f;
    t.f = x;
    t.xs = 0.0f;
    if (has_model) {
        const IcrfTap tap = icrf_linear(x, row, lm1);
        t.f = tap.f;
        fp = tap.fp;
        t.xs = static_cast<float>(tap.x0) + tap.w;     // exact: x0 + w reproduces the clamped scaled value
    }
    t.sig = has_std ? fabsf(__fmul_rn(fp, s)) : 0.0f;
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

// =====================================================================================================
// Statistics kernel.  Block = W warps; a tile of kStatsTile pixels x N frames of one channel is staged in
// shared memory once (ICRF, sigma, Gaussian weight, validity evaluated once per frame element), then warp w
// owns pairs w, w+W, ... (at most SLOTS of them) and keeps their five running sums in registers across all
// tiles of the persistent loop.  One warp reduction and 5 fp64 atomics per (block, pair) at the very end.
// =====================================================================================================
template <int SLOTS, bool HAS_STD, bool RELATIVE>
__global__ void __launch_bounds__(512) pair_stats_kernel(const PairParams p) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames;
    const bool has_model = p.theta != nullptr;
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw);
    float *s_f = reinterpret_cast<float *>(s_tab + (has_model ? C * L : 0));
    float *s_gw = s_f + N * kStatsTile;
    float *s_sig = s_gw + N * kStatsTile;               // HAS_STD only
    float *s_sb = s_sig + N * kStatsTile;               // HAS_STD && RELATIVE only: sig / max(f, 1e-6)
    if (has_model) stage_curve_pairs(s_tab, p.theta, C, L);

    const int c = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    const float lm1 = static_cast<float>(L - 1);
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const int64_t chan_off = static_cast<int64_t>(c) * p.plane;
    const int row_base = p.rows.base(c);
    const bool unc = p.unc_weighting != 0;

    double s0[SLOTS], s1[SLOTS], s2[SLOTS], s3[SLOTS];
    unsigned int s4[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) { s0[s] = 0.0; s1[s] = 0.0; s2[s] = 0.0; s3[s] = 0.0; s4[s] = 0u; }

    const int64_t n_tiles = (p.plane + kStatsTile - 1) / kStatsTile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();   // previous tile fully consumed (also orders the table staging on the first pass)
        const int64_t pix0 = tile * kStatsTile;
        for (int e = threadIdx.x; e < N * kStatsTile; e += blockDim.x) {
            const int n = e / kStatsTile, q = e - n * kStatsTile;
            const int64_t pix = pix0 + q;
            FrameTerms t;
            t.f = 1.0f; t.sig = 0.0f; t.gw = -1.0f;
            if (pix < p.plane) {
                const int64_t o = static_cast<int64_t>(n) * frame_stride + chan_off + pix;
                const float x = __ldcs(p.val + o);
                const float s = HAS_STD ? __ldcs(p.std + o) : 0.0f;
                const int u = static_cast<int>((pix + row_base) % C);
                t = frame_terms(x, s, has_model, s_tab + u * L, lm1, p.valid_lo, p.valid_hi, HAS_STD);
            }
            s_f[e] = t.f;
            s_gw[e] = t.gw;
            if constexpr (HAS_STD) {
                s_sig[e] = t.sig;
                if constexpr (RELATIVE) s_sb[e] = t.sig / fmaxf(t.f, 1e-6f);   // losses.py:55,58
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int pr = warp + s * n_warps;
            if (pr < p.n_pairs) {
                const int fi = p.pairs.i[pr] * kStatsTile, fj = p.pairs.j[pr] * kStatsTile;
                const float r_hi = p.pairs.r_hi[pr], r_lo = p.pairs.r_lo[pr];
                float t3 = 0.0f;
#pragma unroll
                for (int it = 0; it < kStatsTile / 32; ++it) {
                    const int q = lane + 32 * it;
                    const float a = s_f[fi + q], b = s_f[fj + q];
                    const float gi = s_gw[fi + q], gj = s_gw[fj + q];
                    const bool valid = (gi >= 0.0f) && (gj >= 0.0f);
                    const float d = ratio_residual(a, b, r_hi, r_lo);
                    float inv = 1.0f, ell;
                    if constexpr (RELATIVE) {
                        inv = __frcp_rn(fmaf(b, r_hi, 1e-6f));              // 1 / (expected + 1e-6), losses.py:45
                        ell = fabsf(d * inv);                               // es can be negative once the curve dips below 0
                    } else {
                        ell = fabsf(d);
                    }
                    float wt = gi + gj;
                    float err = 0.0f;
                    if constexpr (HAS_STD) {
                        const float sa = s_sig[fi + q];
                        if constexpr (RELATIVE) {
                            const float t1 = sa * inv;
                            const float t2 = a * s_sb[fj + q] * inv;
                            err = sqrtf(fmaf(t1, t1, fmaf(t2, t2, 1e-6f)));   // losses.py:57-60
                        } else {
                            const float rs = r_hi * s_sig[fj + q];
                            err = sqrtf(fmaf(sa, sa, rs * rs));               // losses.py:62
                        }
                        if (unc) wt += __frcp_rn(err + 1e-6f);                // losses.py:97
                    }
                    if (valid) {
                        const double W = static_cast<double>(wt), Ld = static_cast<double>(ell);
                        const double WL = W * Ld;
                        s0[s] += W;
                        s1[s] += WL;
                        s2[s] = fma(WL, Ld, s2[s]);
                        t3 += err;
                        s4[s] += 1u;
                    }
                }
                if constexpr (HAS_STD) s3[s] += static_cast<double>(t3);
            }
        }
    }

#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int pr = warp + s * n_warps;
        if (pr < p.n_pairs) {     // warp-uniform
            double v0 = s0[s], v1 = s1[s], v2 = s2[s], v3 = s3[s];
            unsigned int v4 = s4[s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                v4 += __shfl_xor_sync(0xffffffffu, v4, o);
            }
            if (lane == 0) {
                double *out = p.sums + (static_cast<int64_t>(pr) * C + c) * 5;
                atomicAdd(out + 0, v0);
                atomicAdd(out + 1, v1);
                atomicAdd(out + 2, v2);
                if constexpr (HAS_STD) atomicAdd(out + 3, v3);
                atomicAdd(out + 4, static_cast<double>(v4));
            }
        }
    }
}

// =====================================================================================================
// Gradient kernel.  Each warp owns 32 pixels of one channel at a time: the per-frame terms of those pixels
// go to the warp's private shared-memory slice, every pair is visited by the same lane that owns the pixel
// (so the per-frame upstream G[n] accumulates without atomics), and each frame element ends with ONE
// vectorised reduction  red.global.add.v2.f32 {G(1-w), G w}  into one of kGradCopies replicated tables.
// =====================================================================================================
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Table layout per copy: A rows then B rows, each C x (L+2).  An element with even x0 adds its two taps at
// A[u][x0], A[u][x0+1]; one with odd x0 adds them at B[u][x0+1], B[u][x0+2] (8-byte aligned in both cases).
// grad[u][k] = A[u][k] + B[u][k+1].
__device__ __forceinline__ void scatter_taps(float *copy, int C, int L, int u, float xs, float g) {
    const float fl = floorf(xs);
    const int x0 = static_cast<int>(fl);
    const float w = xs - fl;
    const int lp = L + 2;
    float *base = copy + ((x0 & 1) ? (C * lp + u * lp + x0 + 1) : (u * lp + x0));
    red_add_v2(base, g * (1.0f - w), g * w);
}

template <bool HAS_STD, bool RELATIVE>
__global__ void __launch_bounds__(256) pair_grad_kernel(const PairParams p) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames;
    float2 *s_tab = reinterpret_cast<float2 *>(s_raw);
    stage_curve_pairs(s_tab, p.theta, C, L);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    constexpr int kArrays = 6;
    float *slice = reinterpret_cast<float *>(s_tab + C * L) + warp * (kArrays * N * 32);
    float *s_f = slice, *s_gw = slice + N * 32, *s_xs = slice + 2 * N * 32, *s_g = slice + 3 * N * 32;
    float *s_sig = slice + 4 * N * 32, *s_ib = slice + 5 * N * 32;   // HAS_STD: sigma, 1 / max(f, 1e-6)

    const int c = blockIdx.y;
    const float lm1 = static_cast<float>(L - 1);
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const int64_t chan_off = static_cast<int64_t>(c) * p.plane;
    const int row_base = p.rows.base(c);
    const bool unc = p.unc_weighting != 0;
    const int lp = L + 2;
    float *copy = p.hist + static_cast<int64_t>((blockIdx.x * n_warps + warp) % kGradCopies) * (2 * C * lp);

    const int64_t n_groups = (p.plane + 31) / 32;
    for (int64_t grp = static_cast<int64_t>(blockIdx.x) * n_warps + warp; grp < n_groups;
         grp += static_cast<int64_t>(gridDim.x) * n_warps) {
        const int64_t pix = grp * 32 + lane;
        const bool live = pix < p.plane;
        const int u = static_cast<int>((pix + row_base) % C);
        __syncwarp();
        for (int n = 0; n < N; ++n) {
            FrameTerms t;
            t.f = 1.0f; t.sig = 0.0f; t.gw = -1.0f; t.xs = 0.0f;
            if (live) {
                const int64_t o = static_cast<int64_t>(n) * frame_stride + chan_off + pix;
                const float x = __ldcs(p.val + o);
                const float s = HAS_STD ? __ldcs(p.std + o) : 0.0f;
                t = frame_terms(x, s, true, s_tab + u * L, lm1, p.valid_lo, p.valid_hi, HAS_STD);
            }
            const int e = n * 32 + lane;
            s_f[e] = t.f; s_gw[e] = t.gw; s_xs[e] = t.xs; s_g[e] = 0.0f;
            if constexpr (HAS_STD) {
                s_sig[e] = t.sig;
                s_ib[e] = __frcp_rn(fmaxf(t.f, 1e-6f));
            }
        }
        __syncwarp();
        for (int pr = 0; pr < p.n_pairs; ++pr) {
            const int ei = p.pairs.i[pr] * 32 + lane, ej = p.pairs.j[pr] * 32 + lane;
            const float gi = s_gw[ei], gj = s_gw[ej];
            if (!((gi >= 0.0f) && (gj >= 0.0f))) continue;      // masked pair elements carry no gradient
            const float r_hi = p.pairs.r_hi[pr], r_lo = p.pairs.r_lo[pr];
            const float up = static_cast<float>(__ldg(p.upstream + static_cast<int64_t>(pr) * C + c));
            const float a = s_f[ei], b = s_f[ej];
            const float d = ratio_residual(a, b, r_hi, r_lo);
            float wt = gi + gj;
            float ga, gb;
            if constexpr (RELATIVE) {
                const float inv = __frcp_rn(fmaf(b, r_hi, 1e-6f));
                const float q = d * inv;
                const float sgn = (q > 0.0f) ? 1.0f : ((q < 0.0f) ? -1.0f : 0.0f);
                float extra_a = 0.0f, extra_b = 0.0f;
                if constexpr (HAS_STD) {
                    if (unc) {
                        // the inverse-uncertainty weight depends on the curve through a, es and max(b, 1e-6)
                        const float sa = s_sig[ei], sb = s_sig[ej], ib = s_ib[ej];
                        const float t1 = sa * inv;
                        const float c2 = sb * ib * inv;
                        const float t2 = a * c2;
                        const float T = fmaf(t1, t1, fmaf(t2, t2, 1e-6f));
                        const float err = sqrtf(T);
                        const float rw = __frcp_rn(err + 1e-6f);
                        wt += rw;
                        const float ell = fabsf(q);
                        const float m = static_cast<float>(__ldg(p.mean + static_cast<int64_t>(pr) * C + c));
                        // dm/dWt * dWt/derr * derr/dT = (l - m) U * (-rw^2) * 1/(2 err)
                        const float k = (ell - m) * up * (-0.5f * rw * rw) * __frcp_rn(err);
                        const float dT_da = 2.0f * t2 * c2;
                        const float dT_des = -2.0f * inv * (t1 * t1 + t2 * t2);
                        const float dT_dbs = (b >= 1e-6f) ? (-2.0f * t2 * t2 * ib) : 0.0f;
                        extra_a = k * dT_da;
                        extra_b = k * (dT_des * r_hi + dT_dbs);
                    }
                }
                const float base = wt * up * sgn * inv;
                ga = base + extra_a;                                   // dl/da = sgn / es
                gb = -base * r_hi * (a + 1e-6f) * inv + extra_b;       // dl/db = -sgn r (a + 1e-6) / es^2
            } else {
                const float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
                if constexpr (HAS_STD) {
                    if (unc) {
                        const float sa = s_sig[ei], rs = r_hi * s_sig[ej];
                        wt += __frcp_rn(sqrtf(fmaf(sa, sa, rs * rs)) + 1e-6f);   // constant wrt the curve
                    }
                }
                const float base = wt * up * sgn;
                ga = base;
                gb = -base * r_hi;
            }
            s_g[ei] += ga;
            s_g[ej] += gb;
        }
        __syncwarp();
        for (int n = 0; n < N; ++n) {
            const float g = s_g[n * 32 + lane];
            if (g != 0.0f) scatter_taps(copy, C, L, u, s_xs[n * 32 + lane], g);
        }
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

// grad[u][k] += sum over copies of A[u][k] + B[u][k+1], in float64
__global__ void grad_finalize_kernel(const float *__restrict__ hist, double *grad, int C, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * L) return;
    const int u = i / L, k = i - u * L;
    const int lp = L + 2;
    double acc = 0.0;
    for (int r = 0; r < kGradCopies; ++r) {
        const float *copy = hist + static_cast<int64_t>(r) * (2 * C * lp);
        acc += static_cast<double>(copy[u * lp + k]) + static_cast<double>(copy[C * lp + u * lp + k + 1]);
    }
    grad[i] += acc;
}

}  // namespace clair

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace clair;

namespace {

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

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

// choose warps per block W (4..16) and register slots S (1..kMaxSlots) with W*S >= count and the least idle slots
void pick_stats_shape(int count, int &warps, int &slots) {
    int best_w = 8, best_s = kMaxSlots, best_waste = 1 << 30;
    for (int s = 1; s <= kMaxSlots; ++s) {
        for (int w = 4; w <= 16; ++w) {
            if (w * s < count) continue;
            const int waste = (w * s - count) * 1000 / (w * s) * 8 + s;   // idle fraction first, then fewer registers
            if (waste < best_waste) { best_waste = waste; best_w = w; best_s = s; }
        }
    }
    warps = best_w; slots = best_s;
}

}  // namespace

extern "C" int clair_pair_stats(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                                const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                                int n_pairs, const float *theta_dev, int lut_size, const int32_t *curve_row_base_host,
                                float valid_lo, float valid_hi, int relative, int unc_weighting, double *sums_dev,
                                void *stream) {
    if (!val_dev || !sums_dev) return fail(CLAIR_E_ARG, "clair_pair_stats: null buffer");
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_i_host || !pair_j_host || !pair_ratio_host)))
        return fail(CLAIR_E_ARG, "clair_pair_stats: pair table missing");
    if (theta_dev == nullptr && lut_size <= 0) lut_size = 2;
    if (int rc = check_geometry("clair_pair_stats", n_frames, n_channels, plane, lut_size, true)) return rc;
    if (n_pairs > CLAIR_MAX_PAIRS) return fail(CLAIR_E_LIMIT, "clair_pair_stats: more than CLAIR_MAX_PAIRS pairs");
    if (n_pairs == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool has_std = std_dev != nullptr;
    const int arrays = 2 + (has_std ? (relative ? 2 : 1) : 0);
    const size_t smem = (theta_dev ? sizeof(float2) * n_channels * lut_size : 0) + sizeof(float) * arrays * n_frames * kStatsTile;
    const int per_launch = 16 * kMaxSlots < kMaxPairsPerLaunch ? 16 * kMaxSlots : kMaxPairsPerLaunch;   // 64
    const int n_launches = (n_pairs + per_launch - 1) / per_launch;
    int first = 0;
    for (int l = 0; l < n_launches; ++l) {
        const int count = (n_pairs - first + (n_launches - l) - 1) / (n_launches - l);   // balanced chunks
        PairParams p{};
        p.val = val_dev; p.std = std_dev; p.theta = theta_dev;
        p.sums = sums_dev + static_cast<int64_t>(first) * n_channels * 5;
        p.plane = plane; p.n_frames = n_frames; p.n_channels = n_channels; p.lut = lut_size;
        p.n_pairs = count; p.unc_weighting = unc_weighting; p.valid_lo = valid_lo; p.valid_hi = valid_hi;
        fill_rows(p.rows, curve_row_base_host, n_channels, plane);
        if (int rc = fill_pairs("clair_pair_stats", p.pairs, pair_i_host, pair_j_host, pair_ratio_host, first, count, n_frames)) return rc;
        int warps, slots;
        pick_stats_shape(count, warps, slots);
        const int64_t n_tiles = (plane + kStatsTile - 1) / kStatsTile;
        const int blocks_per_sm = std::max(1, std::min<int>(2048 / (warps * 32), static_cast<int>((200 * 1024) / std::max<size_t>(smem, 1))));
        const int64_t gx = std::min<int64_t>(n_tiles, std::max<int64_t>(1, static_cast<int64_t>(sm_count()) * blocks_per_sm / n_channels));
        dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels));
#define LAUNCH_STATS(S, HS, RL)                                                             \
    do {                                                                                    \
        if (int rc = set_smem(pair_stats_kernel<S, HS, RL>, smem)) return rc;               \
        pair_stats_kernel<S, HS, RL><<<grid, warps * 32, smem, s>>>(p);                     \
    } while (0)
#define DISPATCH_STATS(S)                                                                   \
    do {                                                                                    \
        if (has_std) { if (relative) LAUNCH_STATS(S, true, true); else LAUNCH_STATS(S, true, false); } \
        else { if (relative) LAUNCH_STATS(S, false, true); else LAUNCH_STATS(S, false, false); }       \
    } while (0)
        switch (slots) {
            case 1: DISPATCH_STATS(1); break;
            case 2: DISPATCH_STATS(2); break;
            case 3: DISPATCH_STATS(3); break;
            default: DISPATCH_STATS(4); break;
        }
#undef DISPATCH_STATS
#undef LAUNCH_STATS
        if (int rc = launched("pair_stats_kernel")) return rc;
        first += count;
    }
    return 0;
}

extern "C" size_t clair_grad_workspace_bytes(int n_channels, int lut_size) {
    if (n_channels <= 0 || lut_size <= 0) return 0;
    return sizeof(float) * kGradCopies * 2 * static_cast<size_t>(n_channels) * (lut_size + 2);
}

namespace {
int finalize_grad(const float *hist, double *grad, int C, int L, cudaStream_t s) {
    const int n = C * L;
    grad_finalize_kernel<<<(n + 255) / 256, 256, 0, s>>>(hist, grad, C, L);
    return launched("grad_finalize_kernel");
}
}  // namespace

extern "C" int clair_pair_grad(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                               const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                               int n_pairs, const float *theta_dev, int lut_size, const int32_t *curve_row_base_host,
                               float valid_lo, float valid_hi, int relative, int unc_weighting,
                               const double *upstream_dev, const double *mean_dev, double *grad_theta_dev,
                               void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!val_dev || !theta_dev || !upstream_dev || !mean_dev || !grad_theta_dev || !workspace_dev)
        return fail(CLAIR_E_ARG, "clair_pair_grad: null buffer");
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_i_host || !pair_j_host || !pair_ratio_host)))
        return fail(CLAIR_E_ARG, "clair_pair_grad: pair table missing");
    if (int rc = check_geometry("clair_pair_grad", n_frames, n_channels, plane, lut_size, true)) return rc;
    if (n_pairs > CLAIR_MAX_PAIRS) return fail(CLAIR_E_LIMIT, "clair_pair_grad: more than CLAIR_MAX_PAIRS pairs");
    const size_t need = clair_grad_workspace_bytes(n_channels, lut_size);
    if (workspace_bytes < need || reinterpret_cast<uintptr_t>(workspace_dev) % 16 != 0)
        return fail(CLAIR_E_ARG, "clair_pair_grad: workspace too small or misaligned (see clair_grad_workspace_bytes)");
    if (n_pairs == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaError_t e = cudaMemsetAsync(workspace_dev, 0, need, s); e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(workspace)");
    const bool has_std = std_dev != nullptr;
    const size_t tab_bytes = sizeof(float2) * n_channels * lut_size;
    const size_t per_warp = sizeof(float) * 6 * n_frames * 32;
    const int warps = static_cast<int>(std::max<size_t>(1, std::min<size_t>(8, (200 * 1024 - tab_bytes) / per_warp)));
    const size_t smem = tab_bytes + per_warp * warps;
    const int64_t n_groups = (plane + 31) / 32;
    const int blocks_per_sm = std::max(1, std::min<int>(8, static_cast<int>((200 * 1024) / smem)));
    int first = 0;
    while (first < n_pairs) {
        const int count = std::min(kMaxPairsPerLaunch, n_pairs - first);
        PairParams p{};
        p.val = val_dev; p.std = std_dev; p.theta = theta_dev;
        p.upstream = upstream_dev + static_cast<int64_t>(first) * n_channels;
        p.mean = mean_dev + static_cast<int64_t>(first) * n_channels;
        p.hist = static_cast<float *>(workspace_dev);
        p.plane = plane; p.n_frames = n_frames; p.n_channels = n_channels; p.lut = lut_size;
        p.n_pairs = count; p.unc_weighting = unc_weighting; p.valid_lo = valid_lo; p.valid_hi = valid_hi;
        fill_rows(p.rows, curve_row_base_host, n_channels, plane);
        if (int rc = fill_pairs("clair_pair_grad", p.pairs, pair_i_host, pair_j_host, pair_ratio_host, first, count, n_frames)) return rc;
        const int64_t gx = std::min<int64_t>((n_groups + warps - 1) / warps,
                                             std::max<int64_t>(1, static_cast<int64_t>(sm_count()) * blocks_per_sm / n_channels));
        dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels));
#define LAUNCH_GRAD(HS, RL)                                                                 \
    do {                                                                                    \
        if (int rc = set_smem(pair_grad_kernel<HS, RL>, smem)) return rc;                   \
        pair_grad_kernel<HS, RL><<<grid, warps * 32, smem, s>>>(p);                         \
    } while (0)
        if (has_std) { if (relative) LAUNCH_GRAD(true, true); else LAUNCH_GRAD(true, false); }
        else { if (relative) LAUNCH_GRAD(false, true); else LAUNCH_GRAD(false, false); }
#undef LAUNCH_GRAD
        if (int rc = launched("pair_grad_kernel")) return rc;
        first += count;
    }
    return finalize_grad(static_cast<const float *>(workspace_dev), grad_theta_dev, n_channels, lut_size, s);
}

extern "C" int clair_icrf_backward_theta(const float *x_dev, const float *grad_y_dev, double *grad_theta_dev,
                                         int n_frames, int n_channels, int64_t plane, int lut_size,
                                         const int32_t *curve_row_base_host, void *workspace_dev, size_t workspace_bytes,
                                         void *stream) {
    if (!x_dev || !grad_y_dev || !grad_theta_dev || !workspace_dev) return fail(CLAIR_E_ARG, "clair_icrf_backward_theta: null buffer");
    if (int rc = check_geometry("clair_icrf_backward_theta", n_frames, n_channels, plane, lut_size, false)) return rc;
    const size_t need = clair_grad_workspace_bytes(n_channels, lut_size);
    if (workspace_bytes < need || reinterpret_cast<uintptr_t>(workspace_dev) % 16 != 0)
        return fail(CLAIR_E_ARG, "clair_icrf_backward_theta: workspace too small or misaligned");
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_icrf_backward_theta: n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaError_t e = cudaMemsetAsync(workspace_dev, 0, need, s); e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(workspace)");
    CurveRows rows;
    fill_rows(rows, curve_row_base_host, n_channels, plane);
    dim3 grid(static_cast<unsigned>((plane + 255) / 256), static_cast<unsigned>(slabs));
    icrf_backward_theta_kernel<<<grid, 256, 0, s>>>(x_dev, grad_y_dev, static_cast<float *>(workspace_dev), plane, n_channels,
                                                    lut_size, rows);
    if (int rc = launched("icrf_backward_theta_kernel")) return rc;
    return finalize_grad(static_cast<const float *>(workspace_dev), grad_theta_dev, n_channels, lut_size, s);
}
