// Per-pixel stack kernels: ICRF forward, single-image linearisation and the fused HDR merge with
// first-order uncertainty.  sm_100a, HBM-bound: every input element is read once with 128-bit streaming
// loads, the ICRF table lives in shared memory, all per-pixel reductions stay in registers and the outputs
// are written once.  See DESIGN.md §3 for the derivation of the single-pass variance form.
#include "clair_common.cuh"
#include "clair_host.h"

namespace clair {

constexpr int kBlock = 256;
constexpr int kFrameChunk = 4;          // frames whose loads are in flight together (8 x LDG.128 / thread)
constexpr float kHdrNegScale = -30.0f;  // training/losses.py:193 default scale, used by inference/hdr_merge.py:95

// =====================================================================================================
// ICRF forward / linearise
// =====================================================================================================
struct ForwardParams {
    const float *x;
    const float *std;      // linearise only
    const float *theta;
    float *y;
    float *dydx;           // optional
    float *sigma;          // linearise only
    int64_t plane;         // H*W
    int n_channels;
    int lut;
    CurveRows rows;
};

// grid: (ceil(plane / (VEC*kBlock)), n_frames * C).  MODE: 0 = LINEAR forward, 1 = LOOKUP forward,
// 2 = linearise (LINEAR + sigma)
template <int VEC, int MODE>
__global__ void __launch_bounds__(kBlock) icrf_forward_kernel(const ForwardParams p) {
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    stage_curve_pairs(s_tab, p.theta, C, L);
    __syncthreads();

    const int64_t item = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
    const int64_t pix = item * VEC;
    if (pix >= p.plane) return;
    const int slab = blockIdx.y;                    // n * C + c
    const int c = slab % C;
    const int64_t off = static_cast<int64_t>(slab) * p.plane + pix;
    const float lm1 = static_cast<float>(L - 1);

    const Pack<VEC> xv = load_stream<VEC>(p.x + off);
    Pack<VEC> sv;
    if constexpr (MODE == 2) {
        if (p.std != nullptr) sv = load_stream<VEC>(p.std + off);
    }
    Pack<VEC> yv, dv, gv;
    int u = static_cast<int>((pix + p.rows.base(c)) % C);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        if constexpr (MODE == 1) {
            yv.v[k] = s_tab[c * L + icrf_lookup_index(xv.v[k], lm1)].x;
        } else {
            const IcrfTap t = icrf_linear(xv.v[k], s_tab + u * L, lm1);
            yv.v[k] = t.f;
            dv.v[k] = t.fp;
            if constexpr (MODE == 2) {
                // sqrt((f' * s)^2), inference/linearization.py:106,132
                const float g = (p.std != nullptr) ? __fmul_rn(t.fp, sv.v[k]) : 0.0f;
                gv.v[k] = sqrtf(__fmul_rn(g, g));
            }
            u = wrap_inc(u, C);
        }
    }
    store_stream<VEC>(p.y + off, yv);
    if constexpr (MODE == 0) {
        if (p.dydx != nullptr) store_stream<VEC>(p.dydx + off, dv);
    }
    if constexpr (MODE == 2) store_stream<VEC>(p.sigma + off, gv);
}

// =====================================================================================================
// HDR merge + uncertainty
// =====================================================================================================
struct HdrParams {
    const float *val;
    const float *std;
    const float *theta;       // nullptr = identity
    double *mean_state;
    float *wsum_state;
    float *var_state;
    void *radiance;
    float *sigma;
    int64_t plane;
    int n_frames;
    int n_channels;
    int lut;
    int gaussian;
    int is_first;
    int is_final;
    int radiance_f64;
    CurveRows rows;
    FrameScale scale;
};

// One thread owns VEC horizontally adjacent pixels of one channel and walks the N frames.
//
// Per frame n (fp32):  w = exp(-30 (x-.5)^2) | 1,  q = w'/w = -60 (x-.5) | 0,  v = f(x)/t,
//                      Wsum += w,  S += w v,
//                      R = s w (f'(x)/t + q v),   Q = s w q
// so that  s * d mean_new / d x  =  alpha R + gamma Q   with per-pixel constants
//      alpha = (W_B/W) / (W_B + 1e-6),   gamma = (W_A/W^2)(mean_B - mean_A) - alpha mean_B,
// and the variance update  sum_n (alpha R_n + gamma Q_n)^2 = alpha^2 SRR + 2 alpha gamma SRQ + gamma^2 SQQ
// needs only three running sums.  The expansion cancels (|gamma Q| can be ~14x the result), so the three
// sums and the final combination are float64; everything per element stays float32.
template <int VEC, bool HAS_STD>
__global__ void __launch_bounds__(kBlock) hdr_merge_kernel(const HdrParams p) {
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    if (has_model) {
        stage_curve_pairs(s_tab, p.theta, C, L);
        __syncthreads();
    }
    const int64_t item = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
    const int64_t pix = item * VEC;
    if (pix >= p.plane) return;
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const int64_t off = static_cast<int64_t>(c) * p.plane + pix;
    const float lm1 = static_cast<float>(L - 1);
    const int u0 = static_cast<int>((pix + p.rows.base(c)) % C);
    const bool gaussian = p.gaussian != 0;
    const int N = p.n_frames;

    float wsum[VEC], wv[VEC];
    double srr[VEC], srq[VEC], sqq[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        wsum[k] = 0.0f; wv[k] = 0.0f;
        srr[k] = 0.0; srq[k] = 0.0; sqq[k] = 0.0;
    }

    for (int n0 = 0; n0 < N; n0 += kFrameChunk) {
        Pack<VEC> xv[kFrameChunk], sv[kFrameChunk];
#pragma unroll
        for (int j = 0; j < kFrameChunk; ++j) {
            if (n0 + j < N) {
                const int64_t o = off + static_cast<int64_t>(n0 + j) * frame_stride;
                xv[j] = load_stream<VEC>(p.val + o);
                if constexpr (HAS_STD) sv[j] = load_stream<VEC>(p.std + o);
            }
        }
#pragma unroll
        for (int j = 0; j < kFrameChunk; ++j) {
            if (n0 + j < N) {
                const float it = p.scale.inv_t[n0 + j];
                int u = u0;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const float x = xv[j].v[k];
                    float f = x, fp = 1.0f;
                    if (has_model) {
                        const IcrfTap t = icrf_linear(x, s_tab + u * L, lm1);
                        f = t.f; fp = t.fp;
                        u = wrap_inc(u, C);
                    }
                    float w = 1.0f, q = 0.0f;
                    if (gaussian) {
                        float d;
                        w = gaussian_weight(x, kHdrNegScale, d);
                        q = -60.0f * d;
                    }
                    const float v = f * it;
                    wsum[k] += w;
                    wv[k] = fmaf(w, v, wv[k]);
                    if constexpr (HAS_STD) {
                        const float ws = w * sv[j].v[k];
                        const double R = static_cast<double>(ws * fmaf(q, v, fp * it));
                        const double Q = static_cast<double>(ws * q);
                        srr[k] = fma(R, R, srr[k]);
                        srq[k] = fma(R, Q, srq[k]);
                        sqq[k] = fma(Q, Q, sqq[k]);
                    }
                }
            }
        }
    }

    // ---- per-pixel merge with the running state (common/statistics.py:88-109) ----
    double mean_new[VEC];
    Pack<VEC> wtot, var_new;
    Pack<VEC> w_a, var_a;
    double mean_a[VEC];
    if (!p.is_first) {
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
        const float wbe = wsum[k] + 1e-6f;                        // statistics.py:76 (fp32 add)
        const double inv_wbe = 1.0 / static_cast<double>(wbe);
        const double mean_b = static_cast<double>(wv[k]) * inv_wbe;
        const float wt = w_a.v[k] + wsum[k];                      // statistics.py:104
        const float frac = wsum[k] / wt;                          // statistics.py:106 (0/0 = NaN as in the reference)
        const double dm = mean_b - mean_a[k];
        mean_new[k] = mean_a[k] + static_cast<double>(frac) * dm;
        wtot.v[k] = wt;
        if constexpr (HAS_STD) {
            const double alpha = static_cast<double>(frac) * inv_wbe;
            double gamma = -alpha * mean_b;
            if (!p.is_first) {
                const double wtd = static_cast<double>(wt);
                gamma += static_cast<double>(w_a.v[k]) / (wtd * wtd) * dm;
            }
            const double upd = alpha * alpha * srr[k] + 2.0 * alpha * gamma * srq[k] + gamma * gamma * sqq[k];
            var_new.v[k] = var_a.v[k] + static_cast<float>(fmax(upd, 0.0));
        }
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
            for (int k = 0; k < VEC; ++k) sg.v[k] = sqrtf(var_new.v[k]);
            store_stream<VEC>(p.sigma + off, sg);
        }
    } else {
        store_stream_f64<VEC>(p.mean_state + off, mean_new);
        store_stream<VEC>(p.wsum_state + off, wtot);
        if constexpr (HAS_STD) store_stream<VEC>(p.var_state + off, var_new);
    }
}

}  // namespace clair

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace clair;

namespace {

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    return 0;
}

int pick_vec(int64_t plane, std::initializer_list<const void *> ptrs) {
    // frame / channel slabs start at multiples of `plane` elements, so plane % VEC == 0 keeps every slab aligned
    int vec = (plane % 4 == 0) ? 4 : (plane % 2 == 0) ? 2 : 1;
    for (const void *q : ptrs) {
        if (q == nullptr) continue;
        const uintptr_t a = reinterpret_cast<uintptr_t>(q);
        while (vec > 1 && (a % (vec * sizeof(float))) != 0) vec >>= 1;
    }
    return vec;
}

}  // namespace

extern "C" int clair_icrf_forward(const float *x_dev, const float *theta_dev, float *y_dev, float *dydx_dev,
                                  int n_frames, int n_channels, int64_t plane, int lut_size, int interp_mode,
                                  const int32_t *curve_row_base_host, void *stream) {
    if (!x_dev || !theta_dev || !y_dev) return fail(CLAIR_E_ARG, "clair_icrf_forward: null buffer");
    if (int rc = check_geometry("clair_icrf_forward", n_frames, n_channels, plane, lut_size, /*limit_frames=*/false)) return rc;
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_LOOKUP)
        return fail(CLAIR_E_MODE, "clair_icrf_forward: interp_mode must be CLAIR_INTERP_LINEAR or CLAIR_INTERP_LOOKUP");
    if (interp_mode == CLAIR_INTERP_LOOKUP && dydx_dev) return fail(CLAIR_E_MODE, "clair_icrf_forward: LOOKUP has no derivative");
    ForwardParams p{};
    p.x = x_dev; p.theta = theta_dev; p.y = y_dev; p.dydx = dydx_dev;
    p.plane = plane; p.n_channels = n_channels; p.lut = lut_size;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane);
    const int vec = pick_vec(plane, {x_dev, y_dev, dydx_dev});
    const size_t smem = sizeof(float2) * n_channels * lut_size;
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_icrf_forward: n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_FWD(V, M)                                                                          \
    do {                                                                                          \
        if (int rc = ensure_smem(icrf_forward_kernel<V, M>, smem)) return rc;                     \
        dim3 grid(static_cast<unsigned>((plane / V + kBlock - 1) / kBlock), static_cast<unsigned>(slabs)); \
        icrf_forward_kernel<V, M><<<grid, kBlock, smem, s>>>(p);                                  \
    } while (0)
    const bool lookup = interp_mode == CLAIR_INTERP_LOOKUP;
    if (vec == 4) { if (lookup) LAUNCH_FWD(4, 1); else LAUNCH_FWD(4, 0); }
    else if (vec == 2) { if (lookup) LAUNCH_FWD(2, 1); else LAUNCH_FWD(2, 0); }
    else { if (lookup) LAUNCH_FWD(1, 1); else LAUNCH_FWD(1, 0); }
#undef LAUNCH_FWD
    return launched("icrf_forward_kernel");
}

extern "C" int clair_linearize(const float *val_dev, const float *std_dev, const float *theta_dev, float *lin_dev,
                               float *sigma_dev, int n_frames, int n_channels, int64_t plane, int lut_size,
                               const int32_t *curve_row_base_host, void *stream) {
    if (!val_dev || !theta_dev || !lin_dev || !sigma_dev) return fail(CLAIR_E_ARG, "clair_linearize: null buffer");
    if (int rc = check_geometry("clair_linearize", n_frames, n_channels, plane, lut_size, false)) return rc;
    ForwardParams p{};
    p.x = val_dev; p.std = std_dev; p.theta = theta_dev; p.y = lin_dev; p.sigma = sigma_dev;
    p.plane = plane; p.n_channels = n_channels; p.lut = lut_size;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane);
    const int vec = pick_vec(plane, {val_dev, std_dev, lin_dev, sigma_dev});
    const size_t smem = sizeof(float2) * n_channels * lut_size;
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_linearize: n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_LIN(V)                                                                             \
    do {                                                                                          \
        if (int rc = ensure_smem(icrf_forward_kernel<V, 2>, smem)) return rc;                     \
        dim3 grid(static_cast<unsigned>((plane / V + kBlock - 1) / kBlock), static_cast<unsigned>(slabs)); \
        icrf_forward_kernel<V, 2><<<grid, kBlock, smem, s>>>(p);                                  \
    } while (0)
    if (vec == 4) LAUNCH_LIN(4); else if (vec == 2) LAUNCH_LIN(2); else LAUNCH_LIN(1);
#undef LAUNCH_LIN
    return launched("icrf_forward_kernel<linearize>");
}

extern "C" int clair_hdr_merge_update(const float *val_dev, const float *std_dev, const double *exposure_host,
                                      int n_frames, const float *theta_dev, int n_channels, int lut_size,
                                      int64_t plane, const int32_t *curve_row_base_host, int gaussian_weights,
                                      double *mean_state_dev, float *wsum_state_dev, float *var_state_dev,
                                      int is_first, int is_final, void *radiance_dev, int radiance_f64,
                                      float *sigma_dev, void *stream) {
    if (!val_dev || !exposure_host) return fail(CLAIR_E_ARG, "clair_hdr_merge_update: null val/exposure");
    if (theta_dev == nullptr && lut_size <= 0) lut_size = 2;   // unused without a model
    if (int rc = check_geometry("clair_hdr_merge_update", n_frames, n_channels, plane, lut_size, true)) return rc;
    const bool need_state = !(is_first && is_final);
    if (need_state && (!mean_state_dev || !wsum_state_dev || (std_dev && !var_state_dev)))
        return fail(CLAIR_E_ARG, "clair_hdr_merge_update: running-state buffers required unless is_first && is_final");
    if (is_final && (!radiance_dev || (std_dev && !sigma_dev)))
        return fail(CLAIR_E_ARG, "clair_hdr_merge_update: output buffers required when is_final");
    HdrParams p{};
    p.val = val_dev; p.std = std_dev; p.theta = theta_dev;
    p.mean_state = mean_state_dev; p.wsum_state = wsum_state_dev; p.var_state = var_state_dev;
    p.radiance = radiance_dev; p.sigma = sigma_dev;
    p.plane = plane; p.n_frames = n_frames; p.n_channels = n_channels; p.lut = lut_size;
    p.gaussian = gaussian_weights; p.is_first = is_first; p.is_final = is_final; p.radiance_f64 = radiance_f64;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane);
    for (int n = 0; n < n_frames; ++n) p.scale.inv_t[n] = static_cast<float>(1.0 / exposure_host[n]);
    int vec = pick_vec(plane, {val_dev, std_dev, wsum_state_dev, var_state_dev, sigma_dev});
    // float64 buffers need twice the alignment for the paired 128-bit stores
    for (const void *q : {static_cast<const void *>(mean_state_dev), static_cast<const void *>(radiance_f64 ? radiance_dev : nullptr)}) {
        if (q && reinterpret_cast<uintptr_t>(q) % 16 != 0) vec = 1;
    }
    if (!radiance_f64 && radiance_dev) vec = std::min(vec, pick_vec(plane, {radiance_dev}));
    const size_t smem = theta_dev ? sizeof(float2) * n_channels * lut_size : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_HDR(V, S)                                                                          \
    do {                                                                                          \
        if (int rc = ensure_smem(hdr_merge_kernel<V, S>, smem)) return rc;                        \
        dim3 grid(static_cast<unsigned>((plane / V + kBlock - 1) / kBlock), static_cast<unsigned>(n_channels)); \
        hdr_merge_kernel<V, S><<<grid, kBlock, smem, s>>>(p);                                     \
    } while (0)
    const bool has_std = std_dev != nullptr;
    if (vec == 4) { if (has_std) LAUNCH_HDR(4, true); else LAUNCH_HDR(4, false); }
    else if (vec == 2) { if (has_std) LAUNCH_HDR(2, true); else LAUNCH_HDR(2, false); }
    else { if (has_std) LAUNCH_HDR(1, true); else LAUNCH_HDR(1, false); }
#undef LAUNCH_HDR
    return launched("hdr_merge_kernel");
}
