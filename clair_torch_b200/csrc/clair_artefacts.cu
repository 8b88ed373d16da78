// Dark-field and flat-field corrections with their first-order variance terms.
//
// Dark field (inference/hdr_merge.py:76-92,117-126; inference/linearization.py:73-91,108-116): hot pixels of the dark
// frame D select a 3x3 Gaussian-blurred copy of the image through a soft mask,
//     m = sigmoid(alpha (D - threshold)),   x' = m B(x) + (1 - m) x          (common/general_functions.py:440-486)
// and both drivers then differentiate with respect to the MIXED image x' (the name `images` is reassigned before
// autograd.grad is called), so the two variance terms collapse into one effective std per element:
//     (g' s)^2 + (g' dx'/dD s_D)^2 = (g' s_eff)^2,   s_eff = sqrt(s^2 + ((B(x) - x) alpha m (1 - m) s_D)^2).
// This pre-pass writes (x', s_eff); the merge / linearise kernels then run unchanged on them.
//
// Flat field (inference/hdr_merge.py:131-153; inference/linearization.py:50-57,118-130):
//     y = v / (F + 1e-6) * mu,  mu = mean_{H,W}(F)                            (common/general_functions.py:182-238)
//     var += (dy/dF s_F)^2,  dy/dF = -v mu / (F + 1e-6)^2  [+ (1/HW) sum_q v_q / (F_q + 1e-6) when mu is inside the graph]
// (hdr_merge differentiates through mu, linearization treats it as a constant — SURVEY.md Q11; the existing variance is
// not rescaled by the gain).
#include "clair_common.cuh"
#include "clair_dark.cuh"
#include "clair_host.h"

namespace clair {

__device__ __forceinline__ int reflect_index(int i, int n) {   // 'reflect' padding by one pixel
    return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

// Row-group form: persistent blocks, VEC adjacent pixels of a row per thread (W % VEC == 0), grid (blocks, N*C).
template <int VEC, bool HAS_STD>
__global__ void __launch_bounds__(256) dark_mix_rows_kernel(const float *__restrict__ val, const float *__restrict__ std,
                                                            const float *__restrict__ dark, const float *__restrict__ dark_std,
                                                            const DarkGeometry g, float *__restrict__ val_out,
                                                            float *__restrict__ std_out) {
    const int64_t slab = static_cast<int64_t>(blockIdx.y) * g.H * g.W;
    const uint32_t n_items = static_cast<uint32_t>(static_cast<int64_t>(g.H) * g.W / VEC);
    const uint32_t groups_per_row = static_cast<uint32_t>(g.W / VEC);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    // the loop bound is warp-uniform (the shuffles need all 32 lanes); lanes past the end are inactive
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item - lane < n_items; item += stride) {
        const bool active = item < n_items;
        const uint32_t row = active ? item / groups_per_row : 0u;
        const uint32_t grp = item - row * groups_per_row;
        const int col = static_cast<int>(grp) * VEC;
        const bool chained_left = lane > 0 && grp > 0;
        const bool chained_right = lane < 31 && grp + 1 < groups_per_row;
        RowGroup<VEC> rg;
        row_group_load<VEC>(val + slab, static_cast<int>(row), col, g, active, chained_left, chained_right, rg);
        const int64_t o = slab + static_cast<int64_t>(active ? item : 0u) * VEC;
        const Pack<VEC> dk = load_stream<VEC>(dark + o);
        Pack<VEC> sv, ds, xo, so;
        if constexpr (HAS_STD) {
            sv = load_stream<VEC>(std + o);
            ds = load_stream<VEC>(dark_std + o);
        }
        float x[VEC], blur[VEC];
        row_group_blur<VEC>(rg, col, g, chained_left, chained_right, x, blur);
        if (!active) continue;
#pragma unroll
        for (int k = 0; k < VEC; k += 2) {      // pixel pairs on packed fp32x2
            f32x2 xm, sm = 0ull;
            dark_mix_value2<HAS_STD>(pack2(x[k], x[k + 1]), pack2(blur[k], blur[k + 1]),
                                     HAS_STD ? pack2(sv.v[k], sv.v[k + 1]) : 0ull, pack2(dk.v[k], dk.v[k + 1]),
                                     HAS_STD ? pack2(ds.v[k], ds.v[k + 1]) : 0ull, g, xm, sm);
            unpack2(xm, xo.v[k], xo.v[k + 1]);
            if constexpr (HAS_STD) unpack2(sm, so.v[k], so.v[k + 1]);
        }
        store_stream<VEC>(val_out + o, xo);
        if constexpr (HAS_STD) store_stream<VEC>(std_out + o, so);
    }
}

// Scalar form for odd widths.  grid: (ceil(W/32), ceil(H/8), N*C)
__global__ void __launch_bounds__(256) dark_mix_kernel(const float *__restrict__ val, const float *__restrict__ std,
                                                       const float *__restrict__ dark, const float *__restrict__ dark_std,
                                                       int H, int W, float threshold, float alpha, float *__restrict__ val_out,
                                                       float *__restrict__ std_out) {
    const int w = blockIdx.x * 32 + threadIdx.x, h = blockIdx.y * 8 + threadIdx.y;
    if (w >= W || h >= H) return;
    const int64_t slab = static_cast<int64_t>(blockIdx.z) * H * W;
    const float *img = val + slab;
    float blur = 0.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int hh = reflect_index(h + dy, H);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int ww = reflect_index(w + dx, W);
            blur = fmaf(blur_tap(dy + 1) * blur_tap(dx + 1), __ldg(img + static_cast<int64_t>(hh) * W + ww), blur);
        }
    }
    const int64_t o = slab + static_cast<int64_t>(h) * W + w;
    const float x = __ldg(val + o);
    const float m = 1.0f / (1.0f + expf(-(__ldg(dark + o) - threshold) * alpha));
    val_out[o] = fmaf(m, blur, (1.0f - m) * x);
    if (std_out != nullptr) {
        const float s = __ldg(std + o);
        const float e = (blur - x) * alpha * m * (1.0f - m) * __ldg(dark_std + o);
        std_out[o] = sqrtf(fmaf(s, s, e * e));
    }
}

// ---- flat field -------------------------------------------------------------------------------------------------
// 1 / a in float64 from the fp32 MUFU seed and two Newton steps (2^-22 -> 2^-44 -> rounding): a handful of DFMAs
// instead of the IEEE division's slow path; anything outside the fp32 exponent range takes the division.
__device__ __forceinline__ double rcp_f64(double a) {
    const double mag = fabs(a);
    if (!(mag > 1e-30 && mag < 1e30)) return 1.0 / a;
    double r = static_cast<double>(rcp_approx(static_cast<float>(a)));
    r = fma(fma(-a, r, 1.0), r, r);
    r = fma(fma(-a, r, 1.0), r, r);
    return r;
}

template <typename V, int VEC>
__device__ __forceinline__ void load_values(const V *p, double (&v)[VEC]) {
    if constexpr (sizeof(V) == 4) {
        const Pack<VEC> t = load_stream<VEC>(reinterpret_cast<const float *>(p));
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = static_cast<double>(t.v[k]);
    } else if constexpr (VEC == 4) {
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(p)), b = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
        v[0] = __ldcs(reinterpret_cast<const double *>(p));
    }
}

template <typename V, int VEC>
__device__ __forceinline__ void store_values(V *p, const double (&v)[VEC]) {
    if constexpr (sizeof(V) == 4) {
        Pack<VEC> t;
#pragma unroll
        for (int k = 0; k < VEC; ++k) t.v[k] = static_cast<float>(v[k]);
        store_stream<VEC>(reinterpret_cast<float *>(p), t);
    } else {
        store_stream_f64<VEC>(reinterpret_cast<double *>(p), v);
    }
}

// plain (L2-allocating) vector loads: the reduce pass leaves value and flat in the 126 MB L2 for the apply pass that
// follows it on the stream (c1: 50 MB), so only sigma and flat_std of the second pass come from DRAM
template <int VEC>
__device__ __forceinline__ Pack<VEC> load_keep(const float *p) {
    Pack<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = *p;
    }
    return r;
}

__device__ __forceinline__ float rcp_newton(float x) {     // MUFU.RCP + one Newton step: <= 1 ulp
    const float r = rcp_approx(x);
    return fmaf(fmaf(-x, r, 1.0f), r, r);
}

// per channel: sums[c*2] = sum F, sums[c*2+1] = sum v / (F + 1e-6) over the plane.  Persistent blocks, VEC elements
// per thread and trip, two trips in flight, one pair of float64 atomics per block.  float64 values: float64 partials;
// fp32 values: fp32 per-thread partials (a thread sums a few dozen terms) widened once at the warp reduction — the
// float64 form spent its time converting (F2F runs on the 16-lane XU pipe: 56 % busy, 17 us for 50 MB).
template <typename V, int VEC>
__global__ void __launch_bounds__(256) flat_reduce_kernel(const V *__restrict__ value, const float *__restrict__ flat, int64_t plane,
                                                          double *sums) {
    const int c = blockIdx.y;
    const float *f_c = flat + c * plane;
    const V *v_c = value ? value + c * plane : nullptr;
    const int64_t n_items = plane / VEC, stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    double sf = 0.0, sv = 0.0;
    if constexpr (sizeof(V) == 4) {
        float pf = 0.0f, pv = 0.0f;
        int terms = 0;
        const float *vf = reinterpret_cast<const float *>(v_c);
        for (int64_t item = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; item < n_items; item += 2 * stride) {
            const bool two = item + stride < n_items;
            const Pack<VEC> f0 = load_keep<VEC>(f_c + item * VEC);
            Pack<VEC> f1, x0, x1;
#pragma unroll
            for (int k = 0; k < VEC; ++k) { f1.v[k] = 1.0f; x0.v[k] = 0.0f; x1.v[k] = 0.0f; }
            if (two) f1 = load_keep<VEC>(f_c + (item + stride) * VEC);
            if (vf) {
                x0 = load_keep<VEC>(vf + item * VEC);
                if (two) x1 = load_keep<VEC>(vf + (item + stride) * VEC);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                pf += f0.v[k] + (two ? f1.v[k] : 0.0f);
                if (vf) pv = fmaf(x0.v[k], rcp_newton(f0.v[k] + 1e-6f), fmaf(x1.v[k], rcp_newton(f1.v[k] + 1e-6f), pv));
            }
            if (++terms == 16) { sf += static_cast<double>(pf); sv += static_cast<double>(pv); pf = 0.0f; pv = 0.0f; terms = 0; }
        }
        sf += static_cast<double>(pf);
        sv += static_cast<double>(pv);
    } else {
        for (int64_t item = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; item < n_items; item += 2 * stride) {
            const bool two = item + stride < n_items;
            const Pack<VEC> f0 = load_stream<VEC>(f_c + item * VEC);
            Pack<VEC> f1;
            if (two) f1 = load_stream<VEC>(f_c + (item + stride) * VEC);
            double v0[VEC], v1[VEC];
            if (v_c) {
                load_values<V, VEC>(v_c + item * VEC, v0);
                if (two) load_values<V, VEC>(v_c + (item + stride) * VEC, v1);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const double f = static_cast<double>(f0.v[k]);
                sf += f;
                if (v_c) sv = fma(v0[k], rcp_f64(f + 1e-6), sv);
            }
            if (two) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const double f = static_cast<double>(f1.v[k]);
                    sf += f;
                    if (v_c) sv = fma(v1[k], rcp_f64(f + 1e-6), sv);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sf += __shfl_xor_sync(0xffffffffu, sf, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    __shared__ double part[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { part[0][warp] = sf; part[1][warp] = sv; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[threadIdx.x][w];
        atomicAdd(sums + 2 * c + threadIdx.x, t);
    }
}

// value (n_images, C, plane) corrected in place; sigma (same shape) updated in place when flat_std is given.
// grid (ceil(plane / (VEC*256)), C, n_images).  fp32 values are corrected in fp32 (reciprocal to 1 ulp), float64 ones in
// float64 as the reference's promoted arithmetic does.
template <typename V, int VEC>
__global__ void __launch_bounds__(256) flat_apply_kernel(V *value, float *sigma, const float *__restrict__ flat,
                                                         const float *__restrict__ flat_std, int64_t plane, int C,
                                                         int mean_in_graph, const double *__restrict__ sums) {
    const int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
    if (p >= plane) return;
    const int c = blockIdx.y;
    const int64_t o = (static_cast<int64_t>(blockIdx.z) * C + c) * plane + p;
    const bool with_sigma = flat_std != nullptr && sigma != nullptr;
    const Pack<VEC> f = load_stream<VEC>(flat + c * plane + p);
    const double inv_plane = 1.0 / static_cast<double>(plane);
    const double mu = sums[2 * c] * inv_plane;
    const double g_mean = mean_in_graph ? sums[2 * c + 1] * inv_plane : 0.0;
    if constexpr (sizeof(V) == 4) {
        float *vp = reinterpret_cast<float *>(value) + o;
        Pack<VEC> x = load_stream<VEC>(vp);
        Pack<VEC> fs, sg;
        if (with_sigma) {
            fs = load_stream<VEC>(flat_std + c * plane + p);
            sg = load_stream<VEC>(sigma + o);
        }
        const float muf = static_cast<float>(mu), gmf = static_cast<float>(g_mean);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float r = rcp_newton(f.v[k] + 1e-6f);
            const float y = x.v[k] * r * muf;
            x.v[k] = y;
            if (with_sigma) {
                const float gs = (gmf - y * r) * fs.v[k];                           // dy/dF = -v mu / (F + eps)^2 [+ mean term]
                sg.v[k] = sqrtf(fmaf(sg.v[k], sg.v[k], gs * gs));
            }
        }
        store_stream<VEC>(vp, x);
        if (with_sigma) store_stream<VEC>(sigma + o, sg);
    } else {
        double v[VEC];
        load_values<V, VEC>(value + o, v);
        Pack<VEC> fs, sg;
        if (with_sigma) {
            fs = load_stream<VEC>(flat_std + c * plane + p);
            sg = load_stream<VEC>(sigma + o);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const double r = rcp_f64(static_cast<double>(f.v[k]) + 1e-6);
            const double y = v[k] * r * mu;
            v[k] = y;
            if (with_sigma) {
                const double gs = (g_mean - y * r) * static_cast<double>(fs.v[k]);   // dy/dF = -v mu / (F + eps)^2 [+ mean term]
                const double s = static_cast<double>(sg.v[k]);
                sg.v[k] = sqrtf(static_cast<float>(fma(s, s, gs * gs)));
            }
        }
        store_values<V, VEC>(value + o, v);
        if (with_sigma) store_stream<VEC>(sigma + o, sg);
    }
}

}  // namespace clair

using namespace clair;

extern "C" int clair_dark_field_mix(const float *val_dev, const float *std_dev, const float *dark_dev, const float *dark_std_dev,
                                    int n_frames, int n_channels, int height, int width, float threshold, float alpha,
                                    float *val_out_dev, float *std_out_dev, void *stream) {
    NvtxRange nvtx_range_("clair_dark_field_mix");
    if (!val_dev || !dark_dev || !val_out_dev) return fail(CLAIR_E_ARG, "clair_dark_field_mix: null buffer");
    if ((std_out_dev != nullptr) && (!std_dev || !dark_std_dev))
        return fail(CLAIR_E_ARG, "clair_dark_field_mix: std and dark_std are required when an effective std is requested");
    if (n_frames <= 0 || n_channels <= 0 || height < 2 || width < 2)
        return fail(CLAIR_E_ARG, "clair_dark_field_mix: need positive sizes and at least 2x2 pixels (reflect padding)");
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_dark_field_mix: n_frames*n_channels exceeds 65535");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int vec = (width % 4 == 0) ? 4 : (width % 2 == 0) ? 2 : 1;
    for (const void *q : {static_cast<const void *>(val_dev), static_cast<const void *>(std_dev), static_cast<const void *>(dark_dev),
                          static_cast<const void *>(dark_std_dev), static_cast<const void *>(val_out_dev), static_cast<const void *>(std_out_dev)})
        while (q && vec > 1 && reinterpret_cast<uintptr_t>(q) % (vec * sizeof(float)) != 0) vec >>= 1;
    const int64_t plane = static_cast<int64_t>(height) * width;
    if (vec > 1 && plane / vec < (1ll << 31)) {
        DarkGeometry g{height, width, threshold, alpha, -alpha * 1.4426950408889634f};
        const int64_t want = (plane / vec + 255) / 256;
        // 16 x the resident blocks: a grid of exactly the resident blocks ran 130 us at c1 size, this one 120 us (the blocks that
        // happen to get the slower memory partitions no longer hold the whole launch back)
        const unsigned gx = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, resident_blocks_per_channel(8 * (g_tuning.aux_waves > 0 ? g_tuning.aux_waves : 16), static_cast<int>(slabs)))));
        dim3 grid(gx, static_cast<unsigned>(slabs));
        const bool has_std = std_out_dev != nullptr;
        if (vec == 4) {
            if (has_std) dark_mix_rows_kernel<4, true><<<grid, 256, 0, s>>>(val_dev, std_dev, dark_dev, dark_std_dev, g, val_out_dev, std_out_dev);
            else dark_mix_rows_kernel<4, false><<<grid, 256, 0, s>>>(val_dev, std_dev, dark_dev, dark_std_dev, g, val_out_dev, std_out_dev);
        } else {
            if (has_std) dark_mix_rows_kernel<2, true><<<grid, 256, 0, s>>>(val_dev, std_dev, dark_dev, dark_std_dev, g, val_out_dev, std_out_dev);
            else dark_mix_rows_kernel<2, false><<<grid, 256, 0, s>>>(val_dev, std_dev, dark_dev, dark_std_dev, g, val_out_dev, std_out_dev);
        }
        return launched("dark_mix_rows_kernel");
    }
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8, static_cast<unsigned>(slabs));
    if (grid.y > 65535) return fail(CLAIR_E_LIMIT, "clair_dark_field_mix: image too tall");
    dark_mix_kernel<<<grid, block, 0, s>>>(val_dev, std_dev, dark_dev, dark_std_dev, height, width,
                                                                           threshold, alpha, val_out_dev, std_out_dev);
    return launched("dark_mix_kernel");
}

extern "C" int clair_flat_field_correct(void *value_dev, int value_f64, float *sigma_dev, const float *flat_dev,
                                        const float *flat_std_dev, int n_images, int n_channels, int64_t plane, int mean_in_graph,
                                        double *scratch_dev, void *stream) {
    NvtxRange nvtx_range_("clair_flat_field_correct");
    if (!value_dev || !flat_dev || !scratch_dev) return fail(CLAIR_E_ARG, "clair_flat_field_correct: null buffer");
    if (n_images <= 0 || n_channels <= 0 || n_channels > CLAIR_MAX_CHANNELS || plane <= 0)
        return fail(CLAIR_E_ARG, "clair_flat_field_correct: bad geometry");
    if (mean_in_graph && n_images != 1)
        return fail(CLAIR_E_MODE, "clair_flat_field_correct: the in-graph mean term is defined for a single image (hdr_merge)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaError_t e = cudaMemsetAsync(scratch_dev, 0, sizeof(double) * 2 * n_channels, s); e != cudaSuccess)
        return fail_cuda(e, "cudaMemsetAsync(scratch)");
    // 4 elements per thread when every plane stays 16-byte (float) / 32-byte (double) aligned
    bool vec4 = plane % 4 == 0;
    for (const void *q : {static_cast<const void *>(value_dev), static_cast<const void *>(sigma_dev), static_cast<const void *>(flat_dev),
                          static_cast<const void *>(flat_std_dev)})
        if (q && reinterpret_cast<uintptr_t>(q) % 16 != 0) vec4 = false;
    const int vec = vec4 ? 4 : 1;
    const int64_t items = plane / vec;
    const unsigned rblocks = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((items + 511) / 512, resident_blocks_per_channel(8 * (g_tuning.aux_waves > 0 ? g_tuning.aux_waves : 1), n_channels))));
    dim3 rgrid(rblocks, n_channels), agrid(static_cast<unsigned>((items + 255) / 256), n_channels, n_images);
#define FLAT_LAUNCH(V, VEC)                                                                                                    \
    do {                                                                                                                       \
        flat_reduce_kernel<V, VEC><<<rgrid, 256, 0, s>>>(mean_in_graph ? static_cast<const V *>(value_dev) : nullptr, flat_dev, plane, \
                                                         scratch_dev);                                                       \
        flat_apply_kernel<V, VEC><<<agrid, 256, 0, s>>>(static_cast<V *>(value_dev), sigma_dev, flat_dev, flat_std_dev, plane, \
                                                        n_channels, mean_in_graph, scratch_dev);                              \
    } while (0)
    if (value_f64) { if (vec4) FLAT_LAUNCH(double, 4); else FLAT_LAUNCH(double, 1); }
    else { if (vec4) FLAT_LAUNCH(float, 4); else FLAT_LAUNCH(float, 1); }
#undef FLAT_LAUNCH
    if (int rc = launched("flat_reduce_kernel")) return rc;
    return launched("flat_apply_kernel");
}
