// Per-pixel stack kernels: ICRF forward, single-image linearisation and the fused HDR merge with
// first-order uncertainty.  sm_100a, HBM-bound: every input element is read once with 128-bit streaming
// loads, the ICRF table lives in shared memory, all per-pixel reductions stay in registers and the outputs
// are written once.  See DESIGN.md §3 for the derivation of the single-pass variance form.
#include "clair_merge.cuh"
#include <cmath>
#include <mutex>

namespace clair {

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
    int64_t plane;         // pixels per slab handled by this launch (H*W, or a band of it)
    int64_t stride;        // elements between consecutive (frame, channel) slabs in every buffer
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

    const int slab = blockIdx.y;                    // n * C + c
    const int c = slab % C;
    const float lm1 = static_cast<float>(L - 1);
    const int64_t slab_off = static_cast<int64_t>(slab) * p.stride;
    const int64_t n_items = (p.plane + VEC - 1) / VEC;
    const int64_t item_stride = static_cast<int64_t>(gridDim.x) * kBlock;
    const bool has_std = MODE == 2 && p.std != nullptr;

    auto finish = [&](int64_t pix, const Pack<VEC> &xv, const Pack<VEC> &sv) {
        const int64_t off = slab_off + pix;
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
                    const float g = has_std ? __fmul_rn(t.fp, sv.v[k]) : 0.0f;
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
        if constexpr (MODE == 1) {
            if (p.sigma != nullptr) {                   // linearise with a LOOKUP model and no std images: zeros (:97)
                Pack<VEC> zero;
#pragma unroll
                for (int k = 0; k < VEC; ++k) zero.v[k] = 0.0f;
                store_stream<VEC>(p.sigma + off, zero);
            }
        }
    };
    // Persistent blocks: the table is staged once per block, not once per 1024 pixels.
    for (int64_t item = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; item < n_items; item += item_stride) {
        const int64_t pix = item * VEC;
        const Pack<VEC> xv = load_stream<VEC>(p.x + slab_off + pix);
        Pack<VEC> sv;
        if (has_std) sv = load_stream<VEC>(p.std + slab_off + pix);
        finish(pix, xv, sv);
    }
}

// Blocks per (frame, channel) slab of the forward / linearise grid.  A block walks several 1024-pixel tiles (the 6 KB table
// is staged once per block), but the grid stays many times larger than what is resident: measured on 24 MP frames, one
// tile per block 0.210 ms (0.84 of the copy peak), exactly the resident blocks 0.225 ms, ~6 tiles per block 0.177 ms (0.99).
static unsigned forward_grid_x(int64_t plane, int vec, int64_t slabs) {
    const int64_t tiles = (plane / vec + kBlock - 1) / kBlock;
    if (g_tuning.fwd_blocks < 0) return static_cast<unsigned>(tiles);
    const int per_block = g_tuning.fwd_blocks > 0 ? g_tuning.fwd_blocks : (tiles * slabs >= 100000 ? 6 : 4);   // c1-sized stacks: 4 (0.94)
    const int64_t resident = std::max<int64_t>(1, (static_cast<int64_t>(device_sm_count()) * 8 + slabs - 1) / slabs);
    return static_cast<unsigned>(std::min(tiles, std::max(resident, (tiles + per_block - 1) / per_block)));
}

// Linearisation straight from the camera's integer codes (SURVEY.md 8(f) rank 2 for the lineariser): the reference's CPU
// transforms CastTo + Normalize (x = fl32(code) / fl32(code_max), an IEEE division: a 256-entry table of quotients for
// 8-bit codes, __fdiv_rn for 16-bit ones) and the missing-std synthesis of MultiFileMapDataset (std = x * m or a
// constant, clair_torch/datasets/base.py:128-133) happen in the load, so a frame crosses PCIe / HBM as 1 or 2 bytes per
// sample instead of 8.  Four codes per thread; results are bit-identical to feeding the CPU-transformed fp32 image.
struct LinearizeCodesParams {
    const void *codes;
    const float *std;         // kStdTensor only
    const float *theta;
    float *lin, *sigma;
    int64_t plane;            // H*W, a multiple of 4
    int n_channels, lut;
    int std_mode;             // kStdNone / kStdTensor / kStdMultiplier / kStdConstant
    float std_value, code_max;
    int mode;                 // CLAIR_INTERP_* of the model
    CurveRows rows;
};

template <int BYTES>
__global__ void __launch_bounds__(kBlock) linearize_codes_kernel(const LinearizeCodesParams p) {
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    float *s_x = reinterpret_cast<float *>(s_tab + C * L);           // 8-bit: code -> fl32(code) / code_max
    stage_curve_pairs(s_tab, p.theta, C, L);
    if constexpr (BYTES == 1) {
        for (int k = threadIdx.x; k < 256; k += blockDim.x) s_x[k] = __fdiv_rn(static_cast<float>(k), p.code_max);
    }
    __syncthreads();
    const int slab = blockIdx.y;                    // n * C + c
    const int c = slab % C;
    const float lm1 = static_cast<float>(L - 1);
    // a block walks several 1024-pixel tiles: the tables are staged once per block (forward_grid_x)
    for (int64_t pix = (static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x) * 4; pix < p.plane;
         pix += static_cast<int64_t>(gridDim.x) * kBlock * 4) {
    const int64_t off = static_cast<int64_t>(slab) * p.plane + pix;
    Pack<4> xv;
    if constexpr (BYTES == 1) {
        const uint32_t w = __ldcs(reinterpret_cast<const uint32_t *>(static_cast<const uint8_t *>(p.codes) + off));
#pragma unroll
        for (int k = 0; k < 4; ++k) xv.v[k] = s_x[(w >> (8 * k)) & 0xffu];
    } else {
        const uint2 w = __ldcs(reinterpret_cast<const uint2 *>(static_cast<const uint16_t *>(p.codes) + off));
        xv.v[0] = __fdiv_rn(static_cast<float>(w.x & 0xffffu), p.code_max);
        xv.v[1] = __fdiv_rn(static_cast<float>(w.x >> 16), p.code_max);
        xv.v[2] = __fdiv_rn(static_cast<float>(w.y & 0xffffu), p.code_max);
        xv.v[3] = __fdiv_rn(static_cast<float>(w.y >> 16), p.code_max);
    }
    Pack<4> sv;
    if (p.std_mode == kStdTensor) {
        sv = load_stream<4>(p.std + off);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            sv.v[k] = (p.std_mode == kStdMultiplier) ? __fmul_rn(xv.v[k], p.std_value) : (p.std_mode == kStdConstant ? p.std_value : 0.0f);
    }
    Pack<4> yv, gv;
    int u = static_cast<int>((pix + p.rows.base(c)) % C);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        IcrfTap t;
        if (p.mode == CLAIR_INTERP_LINEAR) t = icrf_linear(xv.v[k], s_tab + u * L, lm1);
        else icrf_mode_eval_rt(p.mode, xv.v[k], s_tab + (p.mode == CLAIR_INTERP_LOOKUP ? c : u) * L, L, lm1, t.f, t.fp);
        yv.v[k] = t.f;
        const float g = (p.std_mode != kStdNone) ? __fmul_rn(t.fp, sv.v[k]) : 0.0f;     // sqrt((f' * s)^2), linearization.py:106,132
        gv.v[k] = sqrtf(__fmul_rn(g, g));
        u = wrap_inc(u, C);
    }
    store_stream<4>(p.lin + off, yv);
    store_stream<4>(p.sigma + off, gv);
    }
}

// grid: (ceil(plane / kBlock), n_frames * C); one element per thread (CATMULL is off the hot path)
__global__ void __launch_bounds__(kBlock) icrf_catmull_kernel(const ForwardParams p) {
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    stage_curve_pairs(s_tab, p.theta, C, L);
    __syncthreads();
    const int64_t pix = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
    if (pix >= p.plane) return;
    const int slab = blockIdx.y, c = slab % C;
    const int64_t off = static_cast<int64_t>(slab) * p.stride + pix;
    const int u = static_cast<int>((pix + p.rows.base(c)) % C);
    const CatmullTaps t = catmull_taps(__ldcs(p.x + off), L);
    const float2 *row = s_tab + u * L;
    const float g0 = row[t.idx[0]].x, g1 = row[t.idx[1]].x, g2 = row[t.idx[2]].x, g3 = row[t.idx[3]].x;
    // stack([w_i g_i]).sum(dim=0): ((w0 g0 + w1 g1) + w2 g2) + w3 g3
    const float y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w[0], g0), __fmul_rn(t.w[1], g1)), __fmul_rn(t.w[2], g2)),
                              __fmul_rn(t.w[3], g3));
    __stcs(p.y + off, y);
    const float dydx = t.dw[0] * g0 + t.dw[1] * g1 + t.dw[2] * g2 + t.dw[3] * g3;
    if (p.dydx != nullptr) __stcs(p.dydx + off, dydx);
    if (p.sigma != nullptr) {                       // linearise: sqrt((f' * s)^2), inference/linearization.py:106,132
        const float g = (p.std != nullptr) ? __fmul_rn(dydx, __ldcs(p.std + off)) : 0.0f;
        __stcs(p.sigma + off, sqrtf(__fmul_rn(g, g)));
    }
}

// ---- LOOKUP / CATMULL models -------------------------------------------------------------------------------
// Off the fast path (the reference default is LINEAR): one pixel per thread, any input kind, two passes over the
// frames.  Pass 1 forms W_B and sum w v in float64; pass 2 re-reads
// the frames (L1 / L2 hits) and accumulates
//     g_n = s_n w_n [alpha f'_n / t_n + q_n (alpha (v_n - mean_B) + beta)],   beta = (W_A / W^2)(mean_B - mean_A)
// with the difference v_n - mean_B taken in float64: in LOOKUP mode f' = 0 (models/base.py:138-158 has no autograd
// edge to the image), so that difference is ALL of the uncertainty and it cancels to ~1e-3 of v_n.
struct FrameScale64 {
    double inv_t[CLAIR_MAX_FRAMES];
};

__device__ __forceinline__ float load_any_pixel(const HdrParams &p, int64_t o, int n, int c, int64_t pix) {
    if (p.src == kSrcF32) return __ldg(static_cast<const float *>(p.val) + o);
    if (p.hwc) o = (static_cast<int64_t>(n) * p.stride + pix) * 3 + (2 - c);
    if (p.src == kSrcU8) return __fdiv_rn(static_cast<float>(__ldg(static_cast<const uint8_t *>(p.val) + o)), p.code_max);
    return __fdiv_rn(static_cast<float>(__ldg(static_cast<const uint16_t *>(p.val) + o)), p.code_max);
}

__device__ __forceinline__ float load_any_std(const HdrParams &p, int64_t o, float x) {
    if (p.std_mode == kStdTensor) return __ldg(p.std + o);
    return (p.std_mode == kStdMultiplier) ? __fmul_rn(x, p.std_value) : p.std_value;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) hdr_merge_modes_kernel(const HdrParams p, const FrameScale64 scale) {
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut, N = p.n_frames;
    stage_curve_pairs(s_tab, p.theta, C, L);
    __syncthreads();
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.stride;
    const float lm1 = static_cast<float>(L - 1);
    const bool gaussian = p.gaussian != 0, has_std = p.std_mode != kStdNone, first = p.is_first != 0;
    const int64_t step = static_cast<int64_t>(gridDim.x) * kBlock;
    for (int64_t pix = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; pix < p.plane; pix += step) {
        const int64_t off = static_cast<int64_t>(c) * p.stride + pix;
        // LOOKUP reads the true channel row (base.py:148-158); CATMULL the k-mod-C row like LINEAR (:217-219)
        const int u = (MODE == CLAIR_INTERP_LOOKUP) ? c : static_cast<int>((pix + p.rows.base(c)) % C);
        const float2 *row = s_tab + u * L;
        double wsum = 0.0, wv = 0.0;
        for (int n = 0; n < N; ++n) {
            const float x = load_any_pixel(p, off + n * frame_stride, n, c, pix);
            float f, fp, d;
            icrf_mode_eval<MODE>(x, row, L, lm1, f, fp);
            const float w = gaussian ? gaussian_weight(x, kHdrNegScaleLog2e, d) : 1.0f;
            wsum += static_cast<double>(w);
            wv = fma(static_cast<double>(w), static_cast<double>(f) * scale.inv_t[n], wv);
        }
        const double wbe = wsum + 1e-6;                                       // statistics.py:76
        const double mean_b = wv / wbe;
        double mean_new, alpha, beta = 0.0, var = 0.0, wtot = wsum;
        if (first) {
            const double frac = (wsum != 0.0) ? 1.0 : static_cast<double>(__int_as_float(0x7fc00000));
            mean_new = frac * mean_b;
            alpha = frac / wbe;
        } else {
            const double w_a = static_cast<double>(__ldg(p.wsum_state + off));
            const double mean_a = __ldg(p.mean_state + off);
            wtot = w_a + wsum;                                                // statistics.py:104
            const double frac = wsum / wtot;                                  // :106
            const double dm = mean_b - mean_a;
            mean_new = mean_a + frac * dm;
            alpha = frac / wbe;
            beta = w_a / (wtot * wtot) * dm;
            if (has_std) var = static_cast<double>(__ldg(p.var_state + off));
        }
        if (has_std) {
            for (int n = 0; n < N; ++n) {
                const int64_t o = off + n * frame_stride;
                const float x = load_any_pixel(p, o, n, c, pix);
                const float sd = load_any_std(p, o, x);
                float f, fp, d = 0.0f;
                icrf_mode_eval<MODE>(x, row, L, lm1, f, fp);
                float w = 1.0f;
                double q = 0.0;
                if (gaussian) {
                    w = gaussian_weight(x, kHdrNegScaleLog2e, d);
                    q = -60.0 * static_cast<double>(d);
                }
                const double v = static_cast<double>(f) * scale.inv_t[n];
                const double g = static_cast<double>(sd) * w * (alpha * fp * scale.inv_t[n] + q * (alpha * (v - mean_b) + beta));
                var = fma(g, g, var);
            }
        }
        if (p.is_final) {
            if (p.radiance_f64) static_cast<double *>(p.radiance)[off] = mean_new;
            else static_cast<float *>(p.radiance)[off] = static_cast<float>(mean_new);
            if (has_std) p.sigma[off] = static_cast<float>(sqrt(var));
        } else {
            p.mean_state[off] = mean_new;
            p.wsum_state[off] = static_cast<float>(wtot);
            if (has_std) p.var_state[off] = static_cast<float>(var);
        }
    }
}

}  // namespace clair

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace clair;

namespace {

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
    NvtxRange nvtx_range_("clair_icrf_forward");
    if (!x_dev || !theta_dev || !y_dev) return fail(CLAIR_E_ARG, "clair_icrf_forward: null buffer");
    if (int rc = check_geometry("clair_icrf_forward", n_frames, n_channels, plane, lut_size, /*limit_frames=*/false)) return rc;
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_LOOKUP && interp_mode != CLAIR_INTERP_CATMULL)
        return fail(CLAIR_E_MODE, "clair_icrf_forward: interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL");
    if (interp_mode == CLAIR_INTERP_LOOKUP && dydx_dev) return fail(CLAIR_E_MODE, "clair_icrf_forward: LOOKUP has no derivative");
    ForwardParams p{};
    p.x = x_dev; p.theta = theta_dev; p.y = y_dev; p.dydx = dydx_dev;
    p.plane = plane; p.stride = plane; p.n_channels = n_channels; p.lut = lut_size;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane);
    const int vec = pick_vec(plane, {x_dev, y_dev, dydx_dev});
    const size_t smem = sizeof(float2) * n_channels * lut_size;
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_icrf_forward: n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (interp_mode == CLAIR_INTERP_CATMULL) {
        if (int rc = ensure_smem(icrf_catmull_kernel, smem)) return rc;
        dim3 grid(static_cast<unsigned>((plane + kBlock - 1) / kBlock), static_cast<unsigned>(slabs));
        icrf_catmull_kernel<<<grid, kBlock, smem, s>>>(p);
        return launched("icrf_catmull_kernel");
    }
#define LAUNCH_FWD(V, M)                                                                          \
    do {                                                                                          \
        if (int rc = ensure_smem(icrf_forward_kernel<V, M>, smem)) return rc;                     \
        dim3 grid(forward_grid_x(plane, V, slabs), static_cast<unsigned>(slabs));                 \
        icrf_forward_kernel<V, M><<<grid, kBlock, smem, s>>>(p);                                  \
    } while (0)
    const bool lookup = interp_mode == CLAIR_INTERP_LOOKUP;
    if (vec == 4) { if (lookup) LAUNCH_FWD(4, 1); else LAUNCH_FWD(4, 0); }
    else if (vec == 2) { if (lookup) LAUNCH_FWD(2, 1); else LAUNCH_FWD(2, 0); }
    else { if (lookup) LAUNCH_FWD(1, 1); else LAUNCH_FWD(1, 0); }
#undef LAUNCH_FWD
    return launched("icrf_forward_kernel");
}

namespace {

int linearize_impl(const char *fn, const float *val_dev, const float *std_dev, const float *theta_dev, float *lin_dev,
                   float *sigma_dev, int n_frames, int n_channels, int64_t plane, int64_t stride, int lut_size, int interp_mode,
                   const int32_t *curve_row_base_host, void *stream) {
    NvtxRange nvtx_range_(fn);
    char msg[160];
    auto bad = [&](int code, const char *what) {
        std::snprintf(msg, sizeof(msg), "%s: %s", fn, what);
        return fail(code, msg);
    };
    if (!val_dev || !theta_dev || !lin_dev || !sigma_dev) return bad(CLAIR_E_ARG, "null buffer");
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, false)) return rc;
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_LOOKUP && interp_mode != CLAIR_INTERP_CATMULL)
        return bad(CLAIR_E_MODE, "interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL");
    if (interp_mode == CLAIR_INTERP_LOOKUP && std_dev)
        return bad(CLAIR_E_MODE, "a LOOKUP model has no derivative to propagate std images through");
    ForwardParams p{};
    p.x = val_dev; p.std = std_dev; p.theta = theta_dev; p.y = lin_dev; p.sigma = sigma_dev;
    p.plane = plane; p.stride = stride; p.n_channels = n_channels; p.lut = lut_size;
    fill_rows(p.rows, curve_row_base_host, n_channels, stride);
    int vec = pick_vec(plane, {val_dev, std_dev, lin_dev, sigma_dev});
    while (vec > 1 && stride % vec != 0) vec >>= 1;
    const size_t smem = sizeof(float2) * n_channels * lut_size;
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return bad(CLAIR_E_LIMIT, "n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (interp_mode == CLAIR_INTERP_CATMULL) {
        if (int rc = ensure_smem(icrf_catmull_kernel, smem)) return rc;
        dim3 grid(static_cast<unsigned>((plane + kBlock - 1) / kBlock), static_cast<unsigned>(slabs));
        icrf_catmull_kernel<<<grid, kBlock, smem, s>>>(p);
        return launched("icrf_catmull_kernel<linearize>");
    }
#define LAUNCH_LIN(V, M)                                                                          \
    do {                                                                                          \
        if (int rc = ensure_smem(icrf_forward_kernel<V, M>, smem)) return rc;                     \
        dim3 grid(forward_grid_x(plane, V, slabs), static_cast<unsigned>(slabs));                 \
        icrf_forward_kernel<V, M><<<grid, kBlock, smem, s>>>(p);                                  \
    } while (0)
    if (interp_mode == CLAIR_INTERP_LOOKUP) {
        if (vec == 4) LAUNCH_LIN(4, 1); else if (vec == 2) LAUNCH_LIN(2, 1); else LAUNCH_LIN(1, 1);
    } else {
        if (vec == 4) LAUNCH_LIN(4, 2); else if (vec == 2) LAUNCH_LIN(2, 2); else LAUNCH_LIN(1, 2);
    }
#undef LAUNCH_LIN
    return launched("icrf_forward_kernel<linearize>");
}

}  // namespace

extern "C" int clair_linearize(const float *val_dev, const float *std_dev, const float *theta_dev, float *lin_dev,
                               float *sigma_dev, int n_frames, int n_channels, int64_t plane, int lut_size, int interp_mode,
                               const int32_t *curve_row_base_host, void *stream) {
    return linearize_impl("clair_linearize", val_dev, std_dev, theta_dev, lin_dev, sigma_dev, n_frames, n_channels, plane, plane,
                          lut_size, interp_mode, curve_row_base_host, stream);
}

extern "C" int clair_linearize_codes(const void *codes_dev, int code_bytes, float code_max, const float *std_dev, int std_mode,
                                     float std_value, const float *theta_dev, float *lin_dev, float *sigma_dev, int n_frames,
                                     int n_channels, int64_t plane, int lut_size, int interp_mode,
                                     const int32_t *curve_row_base_host, void *stream) {
    NvtxRange nvtx_range_("clair_linearize_codes");
    const char *fn = "clair_linearize_codes";
    if (!codes_dev || !theta_dev || !lin_dev || !sigma_dev) return fail(CLAIR_E_ARG, "clair_linearize_codes: null buffer");
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_LOOKUP && interp_mode != CLAIR_INTERP_CATMULL)
        return fail(CLAIR_E_MODE, "clair_linearize_codes: interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL");
    if (interp_mode == CLAIR_INTERP_LOOKUP && std_mode != kStdNone)
        return fail(CLAIR_E_MODE, "clair_linearize_codes: a LOOKUP model has no derivative to propagate std images through");
    if (code_bytes != 1 && code_bytes != 2) return fail(CLAIR_E_MODE, "clair_linearize_codes: code_bytes must be 1 (uint8) or 2 (uint16)");
    if (!(code_max > 0.0f)) return fail(CLAIR_E_ARG, "clair_linearize_codes: code_max must be positive");
    if (std_mode < kStdNone || std_mode > kStdConstant) return fail(CLAIR_E_MODE, "clair_linearize_codes: unknown std_mode");
    if (std_mode == kStdTensor && !std_dev) return fail(CLAIR_E_ARG, "clair_linearize_codes: std_mode = tensor needs std_dev");
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, false)) return rc;
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_linearize_codes: n_frames*n_channels exceeds 65535");
    const uintptr_t align = reinterpret_cast<uintptr_t>(codes_dev) % (4 * code_bytes) | reinterpret_cast<uintptr_t>(lin_dev) % 16 |
                            reinterpret_cast<uintptr_t>(sigma_dev) % 16 | (std_mode == kStdTensor ? reinterpret_cast<uintptr_t>(std_dev) % 16 : 0);
    if (plane % 4 != 0 || align != 0)
        return fail(CLAIR_E_ARG, "clair_linearize_codes: H*W must be a multiple of 4 and the buffers aligned for 4-sample accesses");
    LinearizeCodesParams p{};
    p.codes = codes_dev; p.std = std_dev; p.theta = theta_dev; p.lin = lin_dev; p.sigma = sigma_dev;
    p.plane = plane; p.n_channels = n_channels; p.lut = lut_size;
    p.std_mode = std_mode; p.std_value = std_value; p.code_max = code_max; p.mode = interp_mode;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane);
    const size_t smem = sizeof(float2) * n_channels * lut_size + (code_bytes == 1 ? 256 * sizeof(float) : 0);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid(forward_grid_x(plane, 4, slabs), static_cast<unsigned>(slabs));
    if (code_bytes == 1) {
        if (int rc = ensure_smem(linearize_codes_kernel<1>, smem)) return rc;
        linearize_codes_kernel<1><<<grid, kBlock, smem, s>>>(p);
    } else {
        if (int rc = ensure_smem(linearize_codes_kernel<2>, smem)) return rc;
        linearize_codes_kernel<2><<<grid, kBlock, smem, s>>>(p);
    }
    return launched("linearize_codes_kernel");
}

// Camera codes -> the fp32 value (and std) stacks the pair kernels take: CastTo + Normalize + missing-std synthesis on the
// device, so that a 16-bit stack crosses PCIe as 2 bytes per sample instead of 8 (SURVEY.md 8(f) rank 2 for
// measure_linearity / train_icrf, whose kernels are instruction-bound and re-read nothing: expanding once in HBM costs
// ~0.6 ms for the c3 stack against 45 ms less on the link).  4 codes per thread, IEEE division like the CPU transform.
namespace clair {
template <int BYTES>
__global__ void __launch_bounds__(kBlock) expand_codes_kernel(const void *__restrict__ codes, float code_max, int std_mode, float std_value,
                                                             int64_t n_quads, float *__restrict__ val, float *__restrict__ std) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kBlock;
    for (int64_t q = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x; q < n_quads; q += stride) {
        uint32_t c[4];
        if constexpr (BYTES == 1) {
            const uint32_t w = __ldcs(static_cast<const uint32_t *>(codes) + q);
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = (w >> (8 * k)) & 0xffu;
        } else {
            const uint2 w = __ldcs(static_cast<const uint2 *>(codes) + q);
            c[0] = w.x & 0xffffu; c[1] = w.x >> 16; c[2] = w.y & 0xffffu; c[3] = w.y >> 16;
        }
        Pack<4> x, s;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x.v[k] = __fdiv_rn(static_cast<float>(c[k]), code_max);
            s.v[k] = (std_mode == kStdMultiplier) ? __fmul_rn(x.v[k], std_value) : std_value;
        }
        store_stream<4>(val + 4 * q, x);
        if (std != nullptr) store_stream<4>(std + 4 * q, s);
    }
}
}  // namespace clair

extern "C" int clair_copy_band_h2d(void *dst_dev, const void *src_host, int64_t n_slabs, int64_t slab_stride_bytes,
                                   int64_t band_offset_bytes, int64_t band_bytes, void *stream) {
    if (!dst_dev || !src_host) return fail(CLAIR_E_ARG, "clair_copy_band_h2d: null buffer");
    if (n_slabs <= 0 || band_bytes <= 0 || band_offset_bytes < 0 || band_offset_bytes + band_bytes > slab_stride_bytes)
        return fail(CLAIR_E_ARG, "clair_copy_band_h2d: the band must lie inside a slab");
    const cudaError_t e = cudaMemcpy2DAsync(dst_dev, static_cast<size_t>(band_bytes), static_cast<const char *>(src_host) + band_offset_bytes,
                                            static_cast<size_t>(slab_stride_bytes), static_cast<size_t>(band_bytes),
                                            static_cast<size_t>(n_slabs), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpy2DAsync(band)");
    return 0;
}

extern "C" int clair_expand_codes(const void *codes_dev, int code_bytes, float code_max, int std_mode, float std_value,
                                  int64_t n_elements, float *val_dev, float *std_dev, void *stream) {
    NvtxRange nvtx_range_("clair_expand_codes");
    if (!codes_dev || !val_dev) return fail(CLAIR_E_ARG, "clair_expand_codes: null buffer");
    if (code_bytes != 1 && code_bytes != 2) return fail(CLAIR_E_MODE, "clair_expand_codes: code_bytes must be 1 (uint8) or 2 (uint16)");
    if (!(code_max > 0.0f)) return fail(CLAIR_E_ARG, "clair_expand_codes: code_max must be positive");
    if (std_mode != kStdNone && std_mode != kStdMultiplier && std_mode != kStdConstant)
        return fail(CLAIR_E_MODE, "clair_expand_codes: std_mode must be 0 (none), 2 (multiplier) or 3 (constant)");
    if ((std_mode != kStdNone) != (std_dev != nullptr)) return fail(CLAIR_E_ARG, "clair_expand_codes: std_dev goes with std_mode 2 / 3");
    if (n_elements <= 0 || n_elements % 4 != 0 || reinterpret_cast<uintptr_t>(codes_dev) % (4 * code_bytes) != 0 ||
        reinterpret_cast<uintptr_t>(val_dev) % 16 != 0 || reinterpret_cast<uintptr_t>(std_dev) % 16 != 0)
        return fail(CLAIR_E_ARG, "clair_expand_codes: the element count must be a multiple of 4 and the buffers aligned for 4-sample accesses");
    const int64_t quads = n_elements / 4;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>((quads + kBlock - 1) / kBlock, static_cast<int64_t>(device_sm_count()) * 8 * (g_tuning.aux_waves > 0 ? g_tuning.aux_waves : 32)));   // c3: 724 us at 2 waves, 637 us at 32
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (code_bytes == 1) expand_codes_kernel<1><<<grid, kBlock, 0, s>>>(codes_dev, code_max, std_mode, std_value, quads, val_dev, std_dev);
    else expand_codes_kernel<2><<<grid, kBlock, 0, s>>>(codes_dev, code_max, std_mode, std_value, quads, val_dev, std_dev);
    return launched("expand_codes_kernel");
}

// Host in, host out: band b+1 travels to the device (copy engine, in_stream) and band b-1 back to the host (second copy
// engine, out_stream) while the kernel linearises band b.  Both PCIe directions run at once (49.9 GB/s each way measured
// with the two copy engines; a kernel reading and writing pinned host memory itself reaches 39 GB/s each way).
extern "C" int clair_linearize_staged(const float *val_host, const float *std_host, float *lin_host, float *sigma_host,
                                      float *val_stage_dev, float *std_stage_dev, float *lin_stage_dev, float *sigma_stage_dev,
                                      const float *theta_dev, int n_frames, int n_channels, int64_t plane, int lut_size,
                                      int interp_mode, const int32_t *curve_row_base_host, int n_bands, void *in_stream,
                                      void *out_stream, void *stream) {
    const char *fn = "clair_linearize_staged";
    if (!val_host || !lin_host || !sigma_host || !val_stage_dev || !lin_stage_dev || !sigma_stage_dev)
        return fail(CLAIR_E_ARG, "clair_linearize_staged: null host / staging buffer");
    if (std_host && !std_stage_dev) return fail(CLAIR_E_ARG, "clair_linearize_staged: std_host without a std staging buffer");
    if (in_stream == stream || out_stream == stream || in_stream == out_stream)
        return fail(CLAIR_E_ARG, "clair_linearize_staged: in_stream, out_stream and stream must be three different streams");
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, false)) return rc;
    constexpr int64_t kGranule = 1024;
    const int64_t granules = (plane + kGranule - 1) / kGranule;
    const int bands = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(n_bands, 64), granules)));
    const int C = n_channels;
    const size_t slabs = static_cast<size_t>(n_frames) * C, pitch = static_cast<size_t>(plane) * sizeof(float);
    cudaStream_t is = static_cast<cudaStream_t>(in_stream), os = static_cast<cudaStream_t>(out_stream), ks = static_cast<cudaStream_t>(stream);
    cudaEvent_t ev[2 * 64 + 2];
    int n_ev = 0;
    auto cleanup = [&]() { for (int k = 0; k < n_ev; ++k) cudaEventDestroy(ev[k]); };
    auto new_event = [&](cudaEvent_t &e) -> cudaError_t {
        const cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (err == cudaSuccess) ev[n_ev++] = e;
        return err;
    };
#define STAGED_CUDA(call)                                                        \
    do {                                                                         \
        const cudaError_t err_ = (call);                                         \
        if (err_ != cudaSuccess) { cleanup(); return fail_cuda(err_, #call); }   \
    } while (0)
    cudaEvent_t free_ev;                            // staging buffers may still be in use by work queued on `stream`
    STAGED_CUDA(new_event(free_ev));
    STAGED_CUDA(cudaEventRecord(free_ev, ks));
    STAGED_CUDA(cudaStreamWaitEvent(is, free_ev, 0));
    STAGED_CUDA(cudaStreamWaitEvent(os, free_ev, 0));
    int32_t band_base[CLAIR_MAX_CHANNELS];
    int rc = 0;
    for (int b = 0; b < bands && rc == 0; ++b) {
        const int64_t p0 = std::min<int64_t>(plane, granules * b / bands * kGranule);
        const int64_t p1 = (b + 1 == bands) ? plane : std::min<int64_t>(plane, granules * (b + 1) / bands * kGranule);
        if (p1 <= p0) continue;
        const size_t width = static_cast<size_t>(p1 - p0) * sizeof(float);
        STAGED_CUDA(cudaMemcpy2DAsync(val_stage_dev + p0, pitch, val_host + p0, pitch, width, slabs, cudaMemcpyHostToDevice, is));
        if (std_host) STAGED_CUDA(cudaMemcpy2DAsync(std_stage_dev + p0, pitch, std_host + p0, pitch, width, slabs, cudaMemcpyHostToDevice, is));
        cudaEvent_t ready, done;
        STAGED_CUDA(new_event(ready));
        STAGED_CUDA(cudaEventRecord(ready, is));
        STAGED_CUDA(cudaStreamWaitEvent(ks, ready, 0));
        for (int c = 0; c < C; ++c) {
            const int64_t full = curve_row_base_host ? curve_row_base_host[c] : (static_cast<int64_t>(c) * plane) % C;
            band_base[c] = static_cast<int32_t>((full + p0) % C);
        }
        rc = linearize_impl(fn, val_stage_dev + p0, std_host ? std_stage_dev + p0 : nullptr, theta_dev, lin_stage_dev + p0,
                            sigma_stage_dev + p0, n_frames, C, p1 - p0, plane, lut_size, interp_mode, band_base, stream);
        if (rc) break;
        STAGED_CUDA(new_event(done));
        STAGED_CUDA(cudaEventRecord(done, ks));
        STAGED_CUDA(cudaStreamWaitEvent(os, done, 0));
        STAGED_CUDA(cudaMemcpy2DAsync(lin_host + p0, pitch, lin_stage_dev + p0, pitch, width, slabs, cudaMemcpyDeviceToHost, os));
        STAGED_CUDA(cudaMemcpy2DAsync(sigma_host + p0, pitch, sigma_stage_dev + p0, pitch, width, slabs, cudaMemcpyDeviceToHost, os));
    }
    if (rc == 0) {                                  // results are complete once `stream` is: it waits for the last D2H copy
        cudaEvent_t all_out;
        STAGED_CUDA(new_event(all_out));
        STAGED_CUDA(cudaEventRecord(all_out, os));
        STAGED_CUDA(cudaStreamWaitEvent(ks, all_out, 0));
    }
#undef STAGED_CUDA
    cleanup();
    return rc;
}

namespace {

// Shared implementation of clair_hdr_merge_update (src = kSrcF32) and clair_hdr_merge_codes (kSrcU8 / kSrcU16).
// interp_mode: CLAIR_INTERP_LINEAR takes the fused fast kernels; LOOKUP / CATMULL the all-modes kernel.
// plane_stride: elements between channel planes in EVERY buffer (0 = plane); > plane when the call covers a band of rows
// of larger planes (all pointers then address the band's first pixel).
struct DarkOptions {          // fused dark-field mix: dark == nullptr switches it off
    const float *dark = nullptr, *dark_std = nullptr;
    int height = 0, width = 0;
    float threshold = 0.0f, alpha = 0.0f;
    int hwc = 0;              // integer codes in the interleaved (H, W, 3) BGR camera layout
};

// normalise_code16 (clair_merge.cuh) divides by code_max as q0 = a * rcp, q = fma(fma(-q0, d, a), rcp, q0).  Returns rcp when
// that equals the IEEE quotient a / d for EVERY 16-bit code a (checked here, once per code_max: 65536 divisions), else 0 and
// the kernel keeps __fdiv_rn.  65535, 4095, 1023 pass; the check is what makes an arbitrary caller-supplied code_max safe.
float exact_code_reciprocal(float code_max) {
    static std::mutex mu;
    static float cached_max = 0.0f, cached_rcp = 0.0f;
    std::lock_guard<std::mutex> lock(mu);
    if (code_max == cached_max) return cached_rcp;
    const volatile float d = code_max;
    const float rcp = 1.0f / d;
    bool exact = std::isfinite(rcp) && rcp > 0.0f;
    for (uint32_t k = 0; exact && k < 65536u; ++k) {
        const volatile float a = static_cast<float>(k);
        const float q0 = a * rcp;
        const float q = std::fmaf(std::fmaf(-q0, d, a), rcp, q0);
        const float want = a / d;
        exact = (q == want) && !(q == 0.0f && k != 0);
    }
    cached_max = code_max;
    cached_rcp = exact ? rcp : 0.0f;
    return cached_rcp;
}

int hdr_merge_impl(const char *fn, const void *val_dev, int src, float code_max, const float *std_dev, int std_mode,
                   float std_value, const double *exposure_host, int n_frames, const float *theta_dev, int n_channels,
                   int lut_size, int interp_mode, int64_t plane, int64_t plane_stride, const int32_t *curve_row_base_host,
                   int gaussian_weights, double *mean_state_dev, float *wsum_state_dev, float *var_state_dev, int is_first,
                   int is_final, void *radiance_dev, int radiance_f64, float *sigma_dev, void *stream,
                   const DarkOptions &dark = DarkOptions()) {
    NvtxRange nvtx_range_(fn);
    char msg[200];
    auto bad = [&](int code, const char *what) {
        std::snprintf(msg, sizeof(msg), "%s: %s", fn, what);
        return fail(code, msg);
    };
    if (!val_dev || !exposure_host) return bad(CLAIR_E_ARG, "null val/exposure");
    if (theta_dev == nullptr && lut_size <= 0) lut_size = 2;   // unused without a model
    if (int rc = check_geometry(fn, n_frames, n_channels, plane, lut_size, true)) return rc;
    if (std_mode < kStdNone || std_mode > kStdConstant) return bad(CLAIR_E_MODE, "unknown std_mode");
    if (plane_stride == 0) plane_stride = plane;
    if (plane_stride < plane) return bad(CLAIR_E_ARG, "plane_stride must be 0 or >= plane");
    const bool all_modes = theta_dev != nullptr && interp_mode != CLAIR_INTERP_LINEAR;
    if (dark.hwc) {
        if (src == kSrcF32 || n_channels != 3)
            return bad(CLAIR_E_MODE, "the interleaved BGR layout is for uint8 / uint16 codes of 3-channel frames");
        if (dark.dark) return bad(CLAIR_E_MODE, "the fused dark-field mix takes planar fp32 images");
    }
    if (all_modes && interp_mode != CLAIR_INTERP_LOOKUP && interp_mode != CLAIR_INTERP_CATMULL)
        return bad(CLAIR_E_MODE, "interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL");
    if (std_mode == kStdTensor && !std_dev) return bad(CLAIR_E_ARG, "std_mode = tensor needs std_dev");
    const bool has_std = std_mode != kStdNone;
    const bool need_state = !(is_first && is_final);
    if (need_state && (!mean_state_dev || !wsum_state_dev || (has_std && !var_state_dev)))
        return bad(CLAIR_E_ARG, "running-state buffers required unless is_first && is_final");
    if (is_final && (!radiance_dev || (has_std && !sigma_dev))) return bad(CLAIR_E_ARG, "output buffers required when is_final");
    HdrParams p{};
    p.val = val_dev; p.std = std_dev; p.theta = theta_dev;
    p.mean_state = mean_state_dev; p.wsum_state = wsum_state_dev; p.var_state = var_state_dev;
    p.radiance = radiance_dev; p.sigma = sigma_dev;
    p.plane = plane; p.stride = plane_stride; p.n_frames = n_frames; p.n_channels = n_channels; p.lut = lut_size; p.src = src;
    p.hwc = dark.hwc;
    // camera-layout register kernels: next trip's codes requested into L2 for 16-bit codes (measured: c4 1.20 -> 1.10 ms, 16 frames
    // 2.64 -> 2.27 ms; L1 and L2 requests time the same), off for 8-bit codes (c1 46.6 -> 47.1 us, 13 x 24 MP 1.30 -> 1.33 ms)
    p.prefetch = g_tuning.hdr_prefetch > 0 ? g_tuning.hdr_prefetch : ((g_tuning.hdr_prefetch == 0 && src == kSrcU16) ? 2 : 0);
    p.gaussian = gaussian_weights; p.is_first = is_first; p.is_final = is_final; p.radiance_f64 = radiance_f64;
    p.std_mode = std_mode; p.std_value = std_value; p.code_max = code_max;
    p.code_rcp = src == kSrcU16 ? exact_code_reciprocal(code_max) : 0.0f;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane_stride);
    for (int n = 0; n < n_frames; ++n) p.scale.inv_t[n] = static_cast<float>(1.0 / exposure_host[n]);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dark.dark != nullptr) {
        // the fused form covers the configuration the drivers produce; everything else goes through the
        // clair_dark_field_mix pre-pass
        const bool ok = src == kSrcF32 && std_mode == kStdTensor && dark.dark_std != nullptr && !all_modes && n_frames <= kMaxFixedFrames &&
                        dark.width >= 2 && dark.height >= 2 && dark.width % 2 == 0 &&
                        static_cast<int64_t>(dark.height) * dark.width == plane && plane_stride == plane;
        if (!ok)
            return bad(CLAIR_E_MODE, "fused dark-field mix needs fp32 images with std and dark std, a LINEAR model, <= 8 frames, "
                                     "an even width and whole dense planes; use clair_dark_field_mix otherwise");
        for (const void *q : {val_dev, static_cast<const void *>(std_dev), static_cast<const void *>(dark.dark), static_cast<const void *>(dark.dark_std),
                              static_cast<const void *>(wsum_state_dev), static_cast<const void *>(var_state_dev), static_cast<const void *>(sigma_dev),
                              static_cast<const void *>(radiance_dev), static_cast<const void *>(mean_state_dev)})
            if (q && reinterpret_cast<uintptr_t>(q) % 16 != 0) return bad(CLAIR_E_ARG, "fused dark-field mix needs 16-byte aligned buffers");
        p.dark = dark.dark; p.dark_std = dark.dark_std;
        p.dg = DarkGeometry{dark.height, dark.width, dark.threshold, dark.alpha, -dark.alpha * 1.4426950408889634f};
        const size_t smem = theta_dev ? sizeof(float2) * n_channels * lut_size : 0;
        const bool single = is_first && is_final;
        // several frames: a warp walks down a strip of 60 pixels, band by band (hdr_merge_dark_strip_kernel; dark_strip = -1: the
        // grid-stride form with chunked loads)
        // (2..5 frames with the next row's inputs in a second register set; 6..8 frames without it — there the windows would spill)
        const bool strip = n_frames > 1 && g_tuning.dark_strip >= 0;
        auto launch_dark = [&](auto kernel) -> int {
            if (int rc = ensure_smem(kernel, smem)) return rc;
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem);
            int64_t want_blocks = (plane / 2 + kBlock - 1) / kBlock;
            if (strip) {
                // Rows per band: a warp's task is one (strip, band); the launch takes ceil(tasks / resident warps) rounds of
                // R + 1 row steps (the band's first step loads three value rows), so R is chosen to fill the last round
                // (measured at c1 size: 45 rows = 2304 tasks on 2368 warps 116 us; 16 rows = 2.76 rounds 145 us; 12 rows 126 us)
                const int64_t strips = (dark.width + kStripPix - 1) / kStripPix;
                const int64_t slots = std::max<int64_t>(1, resident_blocks_per_channel(std::max(per_sm, 1), n_channels)) * (kBlock / 32);
                int best_rows = 8;
                int64_t best_cost = -1;
                for (int rows = 6; rows <= 128; ++rows) {
                    const int64_t tasks = strips * ((dark.height + rows - 1) / rows);
                    const int64_t cost = ((tasks + slots - 1) / slots) * (rows + 1);
                    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; }
                }
                p.dark_rows = g_tuning.dark_rows > 0 ? g_tuning.dark_rows : best_rows;
                const int64_t tasks = strips * ((dark.height + p.dark_rows - 1) / p.dark_rows);
                want_blocks = (tasks + kBlock / 32 - 1) / (kBlock / 32);
            }
            per_sm = std::max(per_sm, 1) * (g_tuning.hdr_waves > 0 ? g_tuning.hdr_waves : 2);
            const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(want_blocks, resident_blocks_per_channel(per_sm, n_channels)));
            kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels)), kBlock, smem, s>>>(p);
            return 0;
        };
        int rc = 0;
#define DARK_NF(NF) rc = single ? launch_dark(hdr_merge_dark_kernel<NF, true>) : launch_dark(hdr_merge_dark_kernel<NF, false>)
#define DARK_STRIP(NF) \
    if (strip) rc = single ? launch_dark(hdr_merge_dark_strip_kernel<NF, true>) : launch_dark(hdr_merge_dark_strip_kernel<NF, false>); \
    else DARK_NF(NF)
        switch (n_frames) {
            case 1: DARK_NF(1); break;
            case 2: DARK_STRIP(2); break;
            case 3: DARK_STRIP(3); break;
            case 4: DARK_STRIP(4); break;
            case 5: DARK_STRIP(5); break;
            case 6: DARK_STRIP(6); break;
            case 7: DARK_STRIP(7); break;
            default: DARK_STRIP(8); break;
        }
#undef DARK_STRIP
#undef DARK_NF
        if (rc) return rc;
        return launched("hdr_merge_dark_kernel");
    }
    if (all_modes) {
        FrameScale64 scale{};
        for (int n = 0; n < n_frames; ++n) scale.inv_t[n] = 1.0 / exposure_host[n];
        const size_t smem = sizeof(float2) * n_channels * lut_size;
        auto launch_modes = [&](auto kernel) -> int {
            if (int rc = ensure_smem(kernel, smem)) return rc;
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem);
            const int64_t want = (plane + kBlock - 1) / kBlock;
            const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(want, resident_blocks_per_channel(std::max(per_sm, 1) * 2, n_channels)));
            kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels)), kBlock, smem, s>>>(p, scale);
            return 0;
        };
        const int rc = interp_mode == CLAIR_INTERP_LOOKUP ? launch_modes(hdr_merge_modes_kernel<CLAIR_INTERP_LOOKUP>)
                                                           : launch_modes(hdr_merge_modes_kernel<CLAIR_INTERP_CATMULL>);
        if (rc) return rc;
        return launched("hdr_merge_modes_kernel");
    }
    // every slab starts at a multiple of plane_stride elements: the stride has to keep the vector alignment too
    int vec = pick_vec(plane, {src == kSrcF32 ? val_dev : nullptr, std_dev, wsum_state_dev, var_state_dev, sigma_dev});
    while (vec > 1 && plane_stride % vec != 0) vec >>= 1;
    // float64 buffers need twice the alignment for the paired 128-bit stores
    for (const void *q : {static_cast<const void *>(mean_state_dev), static_cast<const void *>(radiance_f64 ? radiance_dev : nullptr)}) {
        if (q && reinterpret_cast<uintptr_t>(q) % 16 != 0) vec = 1;
    }
    if (!radiance_f64 && radiance_dev) vec = std::min(vec, pick_vec(plane, {radiance_dev}));
    if (src != kSrcF32) {
        // four codes per 32- / 64-bit load
        const size_t code_bytes = src == kSrcU8 ? 1 : 2;
        if (dark.hwc && plane_stride % 4 != 0) vec = 0;      // 12 codes per thread: every frame must start 4-pixel aligned
        if (vec != 4 || reinterpret_cast<uintptr_t>(val_dev) % (4 * code_bytes) != 0)
            return bad(CLAIR_E_ARG, "integer ingest needs H*W % 4 == 0 and 16-byte aligned output / std / state buffers");
    }
    size_t smem = (theta_dev ? sizeof(float2) * n_channels * lut_size : 0) + (src == kSrcU8 ? 256 * sizeof(float) : 0);
    const bool single = is_first && is_final;
    // integer ingest takes 4 codes per load, which leaves no registers for more than 8 frames of (R_n, Q_n): 9..16 frames
    // take 2 codes per load (`wide`), like the fp32 stacks
    const int fixed_max = g_tuning.hdr_fixed_max > 0 ? std::min(g_tuning.hdr_fixed_max, kMaxFixedFramesF32) : kMaxFixedFramesF32;
    // the register kernels normalise 16-bit codes with the short division: only when it is exact for this code_max
    const bool fixed = has_std && !g_tuning.hdr_force_dynamic && n_frames <= fixed_max && (src != kSrcU16 || p.code_rcp != 0.0f);
    const bool wide = src != kSrcF32 && fixed && n_frames > kMaxFixedFrames;
    if (wide) vec = 2;
    // measured on B200 (profiles/): with fp32 input 2 pixels per thread keep the fixed-N kernel at 40 registers
    // (6 blocks/SM); 4 pixels per thread need 64.  Integer ingest takes 4 codes per load up to 8 frames, 2 from 9 to 16.
    // 9 .. ~40 frames: (R_n, Q_n) parked in shared memory, 2 pixels per thread (1 when H*W is odd)
    bool parked = false;
    if (src != kSrcF32) {
        const size_t park_bytes = sizeof(float) * 2 * 4 * kBlock * static_cast<size_t>(n_frames);
        parked = has_std && !fixed && !g_tuning.hdr_force_dynamic && smem + park_bytes <= 110 * 1024;
        if (parked) smem += park_bytes;
    }
    if (src == kSrcF32) {
        const int park_vec = std::min(vec, 2);
        const size_t park_bytes = sizeof(float) * 2 * park_vec * kBlock * static_cast<size_t>(n_frames);
        parked = has_std && !fixed && !g_tuning.hdr_force_dynamic && smem + park_bytes <= 110 * 1024;   // >= 2 blocks per SM
        // the register kernel beyond 8 frames exists for 1 and 2 pixels per thread only
        const int vec_cap = (fixed && n_frames > kMaxFixedFrames) ? 2 : (g_tuning.hdr_vec > 0 ? g_tuning.hdr_vec : ((fixed || parked) ? 2 : 4));
        if (vec_cap < vec) vec = vec_cap;
        if (parked) smem += sizeof(float) * 2 * vec * kBlock * static_cast<size_t>(n_frames);
    }
    // camera layout, 9..16 frames: the block's contiguous codes come in by bulk copies (two-stage ring in shared memory)
    // when every copy is 16-byte granular: H*W, the plane stride and the base address
    const bool tma = wide && dark.hwc && g_tuning.hdr_tma >= 0 && plane % 16 == 0 && plane_stride % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(val_dev) % 16 == 0;
    if (tma) smem += 128 + 2 * static_cast<size_t>(n_frames) * kBlock * 2 * 3 * (src == kSrcU8 ? 1 : 2) + 16;
    MergeLaunch m{};
    m.p = p; m.smem = smem; m.want_blocks = (plane / vec + kBlock - 1) / kBlock;
    m.fixed = fixed; m.parked = parked; m.single = single;
    m.std_mode = std_mode; m.n_frames = n_frames; m.n_channels = n_channels; m.stream = s;
    int rc;
    if (wide && tma) rc = launch_merge_codes_hwc_wide_tma(m, src == kSrcU8);
    else if (wide) rc = dark.hwc ? launch_merge_codes_hwc_wide(m, src == kSrcU8) : launch_merge_codes_planar_wide(m, src == kSrcU8);
    else if (src != kSrcF32) rc = dark.hwc ? launch_merge_codes_hwc(m, src == kSrcU8) : launch_merge_codes_planar(m, src == kSrcU8);
    else if (vec == 4) rc = launch_merge_by_std<4, kSrcF32>(m);
    else if (vec == 2) rc = launch_merge_by_std<2, kSrcF32>(m);
    else rc = launch_merge_by_std<1, kSrcF32>(m);
    if (rc) return rc;
    return launched("hdr_merge_kernel");
}

}  // namespace

extern "C" int clair_hdr_merge_update(const float *val_dev, const float *std_dev, const double *exposure_host,
                                      int n_frames, const float *theta_dev, int n_channels, int lut_size, int64_t plane,
                                      const int32_t *curve_row_base_host, int gaussian_weights, double *mean_state_dev,
                                      float *wsum_state_dev, float *var_state_dev, int is_first, int is_final,
                                      void *radiance_dev, int radiance_f64, float *sigma_dev, void *stream) {
    return hdr_merge_impl("clair_hdr_merge_update", val_dev, kSrcF32, 1.0f, std_dev, std_dev ? kStdTensor : kStdNone,
                          0.0f, exposure_host, n_frames, theta_dev, n_channels, lut_size, CLAIR_INTERP_LINEAR, plane, 0,
                          curve_row_base_host, gaussian_weights, mean_state_dev, wsum_state_dev, var_state_dev, is_first,
                          is_final, radiance_dev, radiance_f64, sigma_dev, stream);
}

extern "C" int clair_hdr_merge_codes(const void *codes_dev, int code_bytes, float code_max, const float *std_dev, int std_mode,
                                     float std_value, const double *exposure_host, int n_frames, const float *theta_dev,
                                     int n_channels, int lut_size, int64_t plane, const int32_t *curve_row_base_host,
                                     int gaussian_weights, double *mean_state_dev, float *wsum_state_dev,
                                     float *var_state_dev, int is_first, int is_final, void *radiance_dev, int radiance_f64,
                                     float *sigma_dev, void *stream) {
    if (code_bytes != 1 && code_bytes != 2) return fail(CLAIR_E_MODE, "clair_hdr_merge_codes: code_bytes must be 1 (uint8) or 2 (uint16)");
    if (!(code_max > 0.0f)) return fail(CLAIR_E_ARG, "clair_hdr_merge_codes: code_max must be positive");
    return hdr_merge_impl("clair_hdr_merge_codes", codes_dev, code_bytes == 1 ? kSrcU8 : kSrcU16, code_max, std_dev, std_mode,
                          std_value, exposure_host, n_frames, theta_dev, n_channels, lut_size, CLAIR_INTERP_LINEAR, plane, 0,
                          curve_row_base_host, gaussian_weights, mean_state_dev, wsum_state_dev, var_state_dev, is_first, is_final,
                          radiance_dev, radiance_f64, sigma_dev, stream);
}

namespace {

int check_desc(const char *fn, const clair_merge_desc *d, int &src) {
    char msg[160];
    if (!d) { std::snprintf(msg, sizeof(msg), "%s: null descriptor", fn); return fail(CLAIR_E_ARG, msg); }
    if (d->struct_bytes != sizeof(clair_merge_desc)) {
        std::snprintf(msg, sizeof(msg), "%s: descriptor is %u bytes, this library expects %zu", fn, d->struct_bytes, sizeof(clair_merge_desc));
        return fail(CLAIR_E_ARG, msg);
    }
    if (d->code_bytes != 0 && d->code_bytes != 1 && d->code_bytes != 2) {
        std::snprintf(msg, sizeof(msg), "%s: code_bytes must be 0 (fp32), 1 (uint8) or 2 (uint16)", fn);
        return fail(CLAIR_E_MODE, msg);
    }
    if (d->code_bytes != 0 && !(d->code_max > 0.0f)) {
        std::snprintf(msg, sizeof(msg), "%s: code_max must be positive", fn);
        return fail(CLAIR_E_ARG, msg);
    }
    src = d->code_bytes == 0 ? kSrcF32 : d->code_bytes == 1 ? kSrcU8 : kSrcU16;
    return 0;
}

int std_mode_of(const clair_merge_desc *d) {
    return d->code_bytes == 0 ? (d->std_dev ? kStdTensor : kStdNone) : d->std_mode;
}

}  // namespace

extern "C" int clair_hdr_merge(const clair_merge_desc *d, void *stream) {
    int src = kSrcF32;
    if (int rc = check_desc("clair_hdr_merge", d, src)) return rc;
    DarkOptions dark;
    dark.dark = d->dark_dev; dark.dark_std = d->dark_std_dev; dark.height = d->height; dark.width = d->width;
    dark.threshold = d->dark_threshold; dark.alpha = d->dark_alpha;
    if (d->code_layout != CLAIR_CODES_PLANAR && d->code_layout != CLAIR_CODES_HWC_BGR)
        return fail(CLAIR_E_MODE, "clair_hdr_merge: code_layout must be CLAIR_CODES_PLANAR or CLAIR_CODES_HWC_BGR");
    dark.hwc = d->code_layout == CLAIR_CODES_HWC_BGR;
    return hdr_merge_impl("clair_hdr_merge", d->val_dev, src, d->code_max, d->std_dev, std_mode_of(d), d->std_value, d->exposure_host,
                          d->n_frames, d->theta_dev, d->n_channels, d->lut_size, d->interp_mode, d->plane, d->plane_stride,
                          d->curve_row_base_host, d->gaussian_weights, d->mean_state_dev, d->wsum_state_dev, d->var_state_dev,
                          d->is_first, d->is_final, d->radiance_dev, d->radiance_f64, d->sigma_dev, stream, dark);
}

// Host-resident stack: the copy engine moves band b+1 of every frame plane into the device staging buffers while the
// kernel merges band b, and the kernel writes its outputs wherever radiance_dev / sigma_dev point (pinned host memory
// makes the whole call host-to-host).  The copy engine reads host memory at the full PCIe rate (55 GB/s measured here
// against 51 GB/s for loads issued by the SMs), and only the last band's kernel is not hidden behind a copy.
extern "C" int clair_hdr_merge_staged(const clair_merge_desc *d, const void *val_host, const float *std_host, int n_bands,
                                      void *copy_stream, void *stream) {
    const char *fn = "clair_hdr_merge_staged";
    int src = kSrcF32;
    if (int rc = check_desc(fn, d, src)) return rc;
    if (!val_host || !d->val_dev) return fail(CLAIR_E_ARG, "clair_hdr_merge_staged: null host / staging buffer");
    if (d->plane_stride != 0 && d->plane_stride != d->plane) return fail(CLAIR_E_ARG, "clair_hdr_merge_staged: the stack must be dense (plane_stride = 0)");
    if (int rc = check_geometry(fn, d->n_frames, d->n_channels, d->plane, d->lut_size > 0 ? d->lut_size : 2, true)) return rc;
    const int std_mode = std_mode_of(d);
    if (std_mode == kStdTensor && !std_host) return fail(CLAIR_E_ARG, "clair_hdr_merge_staged: std staging buffer without std_host");
    if (copy_stream == stream) return fail(CLAIR_E_ARG, "clair_hdr_merge_staged: copy_stream must differ from stream");
    if (d->dark_dev) return fail(CLAIR_E_MODE, "clair_hdr_merge_staged: the dark-field mix needs whole planes on the device (clair_hdr_merge)");
    if (d->code_layout != CLAIR_CODES_PLANAR && d->code_layout != CLAIR_CODES_HWC_BGR)
        return fail(CLAIR_E_MODE, "clair_hdr_merge_staged: code_layout must be CLAIR_CODES_PLANAR or CLAIR_CODES_HWC_BGR");
    const bool hwc = d->code_layout == CLAIR_CODES_HWC_BGR;
    // bands of whole 1024-pixel blocks keep every band base 16-byte aligned for all element sizes
    constexpr int64_t kGranule = 1024;
    const int64_t granules = (d->plane + kGranule - 1) / kGranule;
    const int bands = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(n_bands, 64), granules)));
    const size_t esz = src == kSrcF32 ? 4 : src == kSrcU8 ? 1 : 2;
    const int C = d->n_channels;
    const size_t slabs = static_cast<size_t>(d->n_frames) * C;
    cudaStream_t cs = static_cast<cudaStream_t>(copy_stream), ks = static_cast<cudaStream_t>(stream);
    cudaEvent_t ev[65];
    int n_ev = 0;
    auto cleanup = [&]() { for (int k = 0; k < n_ev; ++k) cudaEventDestroy(ev[k]); };
    auto new_event = [&](cudaEvent_t &e) -> cudaError_t {
        const cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (err == cudaSuccess) ev[n_ev++] = e;
        return err;
    };
#define STAGED_CUDA(call)                                                        \
    do {                                                                         \
        const cudaError_t err_ = (call);                                         \
        if (err_ != cudaSuccess) { cleanup(); return fail_cuda(err_, #call); }   \
    } while (0)
    // the staging buffers may still be read by kernels queued earlier on the compute stream
    cudaEvent_t free_ev;
    STAGED_CUDA(new_event(free_ev));
    STAGED_CUDA(cudaEventRecord(free_ev, ks));
    STAGED_CUDA(cudaStreamWaitEvent(cs, free_ev, 0));
    DarkOptions band_opts;
    band_opts.hwc = hwc ? 1 : 0;
    int32_t full_base[CLAIR_MAX_CHANNELS], band_base[CLAIR_MAX_CHANNELS];
    for (int c = 0; c < C; ++c)
        full_base[c] = d->curve_row_base_host ? d->curve_row_base_host[c] : static_cast<int32_t>((static_cast<int64_t>(c) * d->plane) % C);
    int rc = 0;
    for (int b = 0; b < bands && rc == 0; ++b) {
        const int64_t p0 = std::min<int64_t>(d->plane, granules * b / bands * kGranule);
        const int64_t p1 = (b + 1 == bands) ? d->plane : std::min<int64_t>(d->plane, granules * (b + 1) / bands * kGranule);
        if (p1 <= p0) continue;
        // planar: one row per (frame, channel) plane; interleaved: one row per frame, three codes per pixel
        const size_t px_bytes = hwc ? 3 * esz : esz;
        const size_t pitch = static_cast<size_t>(d->plane) * px_bytes, width = static_cast<size_t>(p1 - p0) * px_bytes;
        char *val_stage = static_cast<char *>(const_cast<void *>(d->val_dev)) + p0 * px_bytes;
        STAGED_CUDA(cudaMemcpy2DAsync(val_stage, pitch, static_cast<const char *>(val_host) + p0 * px_bytes, pitch, width,
                                      hwc ? static_cast<size_t>(d->n_frames) : slabs, cudaMemcpyHostToDevice, cs));
        if (std_mode == kStdTensor)
            STAGED_CUDA(cudaMemcpy2DAsync(const_cast<float *>(d->std_dev) + p0, static_cast<size_t>(d->plane) * 4, std_host + p0,
                                          static_cast<size_t>(d->plane) * 4, static_cast<size_t>(p1 - p0) * 4, slabs,
                                          cudaMemcpyHostToDevice, cs));
        cudaEvent_t ready;
        STAGED_CUDA(new_event(ready));
        STAGED_CUDA(cudaEventRecord(ready, cs));
        STAGED_CUDA(cudaStreamWaitEvent(ks, ready, 0));
        for (int c = 0; c < C; ++c) band_base[c] = static_cast<int32_t>((full_base[c] + p0) % C);
        auto at = [&](auto *q) { return q ? q + p0 : q; };
        void *rad = d->radiance_dev ? static_cast<char *>(d->radiance_dev) + p0 * (d->radiance_f64 ? 8 : 4) : nullptr;
        rc = hdr_merge_impl(fn, val_stage, src, d->code_max, at(d->std_dev), std_mode, d->std_value, d->exposure_host, d->n_frames,
                            d->theta_dev, C, d->lut_size, d->interp_mode, p1 - p0, d->plane, band_base, d->gaussian_weights,
                            at(d->mean_state_dev), at(d->wsum_state_dev), at(d->var_state_dev), d->is_first, d->is_final, rad,
                            d->radiance_f64, at(d->sigma_dev), stream, band_opts);
    }
#undef STAGED_CUDA
    cleanup();      // events are released once the work recorded on them has completed
    return rc;
}

// =====================================================================================================
// Streaming weighted mean / variance over frames — WBOMeanVar.update_values + _update_internal_values
// (clair_torch/common/statistics.py:209-259), as used by compute_video_mean_and_std
// (clair_torch/inference/inferential_statistics.py:19-49).  SURVEY.md §8(f) rank 3.
//   batch:  W_B = sum w,  W2_B = sum w^2,  mean_B = sum w v / (W_B + 1e-6)   [weights]   |   mean(v), W_B = W2_B = N   [none]
//           M_B = sum w (v - mean_B)^2
//   merge:  W = W_A + W_B,  M = M_A + M_B + (W_A W_B / W)(mean_B - mean_A)^2,  mean = mean_A + (W_B/W)(mean_B - mean_A)
// One pass over the N frames: the second moment is accumulated around a pivot (the first frame's value), so
// sum w (v - mean_B)^2 = S2 - 2 (mean_B - k) S1 + (mean_B - k)^2 W_B never cancels in fp32.
// =====================================================================================================
namespace clair {

struct FrameStatsParams {
    const float *val;
    const float *weights;     // nullptr = unweighted
    const float *theta;       // nullptr = no linearisation
    float *mean, *m2, *wsum, *wsq;
    int64_t plane;
    int n_frames, n_channels, lut, is_first;
    int mode;                 // CLAIR_INTERP_* of the model
    CurveRows rows;
};

__device__ __forceinline__ float rcp_1ulp(float x) {     // MUFU.RCP + one Newton step
    const float r = rcp_approx(x);
    return fmaf(fmaf(-x, r, 1.0f), r, r);
}

// WEIGHTED / MODES are compile-time so the common case (unweighted video frames, LINEAR model or none) carries no weight
// accumulators and no mode dispatch: 94 -> 40 registers, and the four running-state loads are issued with the frame
// loads instead of after them.  Quotients go through a 1-ulp reciprocal (the IEEE divisions were ~60 of the epilogue's
// instructions per pixel).
#ifndef STATS_CHUNK
#define STATS_CHUNK 6
#endif
template <int VEC, bool WEIGHTED, bool MODES>
__global__ void __launch_bounds__(kBlock, WEIGHTED ? 2 : 4) frame_stats_kernel(const FrameStatsParams p) {
    extern __shared__ float2 s_tab[];
    const int C = p.n_channels, L = p.lut;
    const bool has_model = p.theta != nullptr;
    if (has_model) {
        stage_curve_pairs(s_tab, p.theta, C, L);
        __syncthreads();
    }
    const int c = blockIdx.y;
    const int64_t frame_stride = static_cast<int64_t>(C) * p.plane;
    const float lm1 = static_cast<float>(L - 1);
    const uint32_t n_items = static_cast<uint32_t>(p.plane / VEC);
    const uint32_t item_stride = gridDim.x * kBlock;
    const uint32_t first_item = blockIdx.x * kBlock + threadIdx.x;
    RowCursor<VEC> cur(first_item * VEC, item_stride * VEC, static_cast<uint32_t>(p.rows.base(c)), static_cast<uint32_t>(C));
    const uint32_t tab_bias = curve_row_bias(s_tab);
    const uint32_t row_bytes = static_cast<uint32_t>(L) * 8u;
    const int N = p.n_frames;
    const bool first = p.is_first != 0;
    const float n_f = static_cast<float>(N);
    // frames whose loads are in flight together: unweighted batches of up to 6 frames (a video burst, c1's 5) in ONE chunk
    constexpr int kChunk = WEIGHTED ? kFrameChunk : STATS_CHUNK;

    for (uint32_t item = first_item; item < n_items; item += item_stride, cur.advance()) {
        const uint32_t pix = item * VEC;
        const int64_t off = static_cast<int64_t>(c) * p.plane + pix;
        uint32_t bias[VEC];
        {
            uint32_t u = cur.u0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) { bias[k] = tab_bias + u * row_bytes; u = (u + 1 == cur.C) ? 0u : u + 1; }
        }
        // running state of the pixel: requested first, needed last
        Pack<VEC> a_mean, a_m2, a_w, a_wsq;
        if (!first) {
            a_mean = load_stream<VEC>(p.mean + off); a_m2 = load_stream<VEC>(p.m2 + off);
            a_w = load_stream<VEC>(p.wsum + off); a_wsq = load_stream<VEC>(p.wsq + off);
        }
        float w0[VEC], w2[VEC], s1[VEC], s2[VEC], pivot[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { w0[k] = 0.0f; w2[k] = 0.0f; s1[k] = 0.0f; s2[k] = 0.0f; pivot[k] = 0.0f; }
        for (int n0 = 0; n0 < N; n0 += kChunk) {
            Pack<VEC> xv[kChunk], wv[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (n0 + j < N) {
                    const int64_t o = off + static_cast<int64_t>(n0 + j) * frame_stride;
                    xv[j] = load_stream<VEC>(p.val + o);
                    if constexpr (WEIGHTED) wv[j] = load_stream<VEC>(p.weights + o);
                }
            }
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                if (n0 + j < N) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        float v = xv[j].v[k];
                        if (has_model) {
                            float fp;
                            if constexpr (MODES) {
                                // LOOKUP reads the true channel row, CATMULL the k-mod-C row like LINEAR
                                const uint32_t u = (bias[k] - tab_bias) / row_bytes;
                                icrf_mode_eval_rt(p.mode, v, s_tab + (p.mode == CLAIR_INTERP_LOOKUP ? static_cast<uint32_t>(c) : u) * L, L, lm1, v, fp);
                            } else {
                                icrf_linear_biased(v, bias[k], lm1, v, fp);
                            }
                        }
                        if (n0 + j == 0) pivot[k] = v;
                        const float d = v - pivot[k];
                        if constexpr (WEIGHTED) {
                            const float w = wv[j].v[k];
                            w0[k] += w;
                            w2[k] = fmaf(w, w, w2[k]);
                            s1[k] = fmaf(w, d, s1[k]);
                            s2[k] = fmaf(w * d, d, s2[k]);
                        } else {
                            s1[k] += d;
                            s2[k] = fmaf(d, d, s2[k]);
                        }
                    }
                }
            }
        }
        Pack<VEC> o_mean, o_m2, o_w, o_wsq;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float wb = WEIGHTED ? w0[k] : n_f;
            // weighted: sum w v / (W_B + 1e-6) (statistics.py:223-224); unweighted: plain mean (:227)
            float db;                                                       // mean_B - pivot
            if constexpr (WEIGHTED) {
                const float r = rcp_1ulp(w0[k] + 1e-6f);
                db = s1[k] * r - pivot[k] * (1e-6f * r);
            } else {
                db = s1[k] * rcp_1ulp(wb);
            }
            const float mean_b = pivot[k] + db;
            const float m2_b = fmaxf(s2[k] - 2.0f * db * s1[k] + db * db * wb, 0.0f);
            const float wsq_b = WEIGHTED ? w2[k] : wb;
            if (first) {
                // W_A = 0, mean_A = 0, M_A = 0 (python floats in the reference): mean = (W_B / W_B) mean_B
                o_w.v[k] = wb; o_mean.v[k] = (wb != 0.0f) ? mean_b : __int_as_float(0x7fc00000); o_m2.v[k] = m2_b; o_wsq.v[k] = wsq_b;
            } else {
                const float wt = a_w.v[k] + wb;
                const float rt = rcp_1ulp(wt);
                const float dm = mean_b - a_mean.v[k];
                o_m2.v[k] = a_m2.v[k] + m2_b + (a_w.v[k] * wb * rt) * dm * dm;
                o_mean.v[k] = a_mean.v[k] + (wb * rt) * dm;
                o_w.v[k] = wt;
                o_wsq.v[k] = a_wsq.v[k] + wsq_b;
            }
        }
        store_stream<VEC>(p.mean + off, o_mean); store_stream<VEC>(p.m2 + off, o_m2);
        store_stream<VEC>(p.wsum + off, o_w); store_stream<VEC>(p.wsq + off, o_wsq);
    }
}

}  // namespace clair

extern "C" int clair_frame_stats_update(const float *val_dev, const float *weights_dev, const float *theta_dev, int n_frames,
                                        int n_channels, int64_t plane, int lut_size, int interp_mode,
                                        const int32_t *curve_row_base_host,
                                        float *mean_state_dev, float *m2_state_dev, float *wsum_state_dev,
                                        float *wsq_state_dev, int is_first, void *stream) {
    NvtxRange nvtx_range_("clair_frame_stats_update");
    if (!val_dev || !mean_state_dev || !m2_state_dev || !wsum_state_dev || !wsq_state_dev)
        return fail(CLAIR_E_ARG, "clair_frame_stats_update: null buffer");
    if (theta_dev == nullptr && lut_size <= 0) lut_size = 2;
    if (theta_dev == nullptr) interp_mode = CLAIR_INTERP_LINEAR;
    if (interp_mode != CLAIR_INTERP_LINEAR && interp_mode != CLAIR_INTERP_LOOKUP && interp_mode != CLAIR_INTERP_CATMULL)
        return fail(CLAIR_E_MODE, "clair_frame_stats_update: interp_mode must be CLAIR_INTERP_LOOKUP, _LINEAR or _CATMULL");
    if (int rc = check_geometry("clair_frame_stats_update", n_frames, n_channels, plane, lut_size, false)) return rc;
    FrameStatsParams p{};
    p.mode = interp_mode;
    p.val = val_dev; p.weights = weights_dev; p.theta = theta_dev;
    p.mean = mean_state_dev; p.m2 = m2_state_dev; p.wsum = wsum_state_dev; p.wsq = wsq_state_dev;
    p.plane = plane; p.n_frames = n_frames; p.n_channels = n_channels; p.lut = lut_size; p.is_first = is_first;
    fill_rows(p.rows, curve_row_base_host, n_channels, plane);
    const int vec = pick_vec(plane, {val_dev, weights_dev, mean_state_dev, m2_state_dev, wsum_state_dev, wsq_state_dev});
    const size_t smem = theta_dev ? sizeof(float2) * n_channels * lut_size : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto launch = [&](auto kernel) -> int {
        if (int rc = ensure_smem(kernel, smem)) return rc;
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem);
        const int64_t want = (plane / vec + kBlock - 1) / kBlock;
        const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(want, resident_blocks_per_channel(std::max(per_sm, 1) * (g_tuning.aux_waves > 0 ? g_tuning.aux_waves : 2), n_channels)));
        kernel<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n_channels)), kBlock, smem, s>>>(p);
        return 0;
    };
    const bool weighted = weights_dev != nullptr, modes = theta_dev != nullptr && interp_mode != CLAIR_INTERP_LINEAR;
    int rc;
#define STATS_VEC(W, M) (vec == 4 ? launch(frame_stats_kernel<4, W, M>) : vec == 2 ? launch(frame_stats_kernel<2, W, M>) : launch(frame_stats_kernel<1, W, M>))
    if (weighted) rc = modes ? STATS_VEC(true, true) : STATS_VEC(true, false);
    else rc = modes ? STATS_VEC(false, true) : STATS_VEC(false, false);
#undef STATS_VEC
    if (rc) return rc;
    return launched("frame_stats_kernel");
}
