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
#include "clair_host.h"

namespace clair {

// torchvision GaussianBlur(kernel_size=3, sigma=1): exp(-0.5 d^2) / sum, in fp32
__constant__ float kBlurTaps[3] = {0.27406862f, 0.45186276f, 0.27406862f};

__device__ __forceinline__ int reflect_index(int i, int n) {   // 'reflect' padding by one pixel
    return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

// grid: (ceil(W/32), ceil(H/8), N*C)
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
            blur = fmaf(kBlurTaps[dy + 1] * kBlurTaps[dx + 1], __ldg(img + static_cast<int64_t>(hh) * W + ww), blur);
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

// per channel: sums[c*2] = sum F, sums[c*2+1] = sum v / (F + 1e-6) over the plane (float64 atomics)
template <typename V>
__global__ void __launch_bounds__(256) flat_reduce_kernel(const V *__restrict__ value, const float *__restrict__ flat, int64_t plane,
                                                          double *sums) {
    const int c = blockIdx.y;
    double sf = 0.0, sv = 0.0;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < plane; p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double f = static_cast<double>(flat[c * plane + p]);
        sf += f;
        if (value != nullptr) sv += static_cast<double>(value[c * plane + p]) / (f + 1e-6);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sf += __shfl_xor_sync(0xffffffffu, sf, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sums + 2 * c, sf);
        atomicAdd(sums + 2 * c + 1, sv);
    }
}

// value (n_images, C, plane) corrected in place; sigma (same shape) updated in place when flat_std is given
template <typename V>
__global__ void __launch_bounds__(256) flat_apply_kernel(V *value, float *sigma, const float *__restrict__ flat,
                                                         const float *__restrict__ flat_std, int64_t plane, int C,
                                                         int mean_in_graph, const double *__restrict__ sums) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= plane) return;
    const int c = blockIdx.y;
    const int64_t o = (static_cast<int64_t>(blockIdx.z) * C + c) * plane + p;
    const double mu = sums[2 * c] / static_cast<double>(plane);
    const double f = static_cast<double>(flat[c * plane + p]) + 1e-6;
    const double v = static_cast<double>(value[o]);
    value[o] = static_cast<V>(v / f * mu);
    if (flat_std != nullptr && sigma != nullptr) {
        double g = -v * mu / (f * f);
        if (mean_in_graph) g += sums[2 * c + 1] / static_cast<double>(plane);
        const double gs = g * static_cast<double>(flat_std[c * plane + p]);
        const double s = static_cast<double>(sigma[o]);
        sigma[o] = static_cast<float>(sqrt(s * s + gs * gs));
    }
}

}  // namespace clair

using namespace clair;

extern "C" int clair_dark_field_mix(const float *val_dev, const float *std_dev, const float *dark_dev, const float *dark_std_dev,
                                    int n_frames, int n_channels, int height, int width, float threshold, float alpha,
                                    float *val_out_dev, float *std_out_dev, void *stream) {
    if (!val_dev || !dark_dev || !val_out_dev) return fail(CLAIR_E_ARG, "clair_dark_field_mix: null buffer");
    if ((std_out_dev != nullptr) && (!std_dev || !dark_std_dev))
        return fail(CLAIR_E_ARG, "clair_dark_field_mix: std and dark_std are required when an effective std is requested");
    if (n_frames <= 0 || n_channels <= 0 || height < 2 || width < 2)
        return fail(CLAIR_E_ARG, "clair_dark_field_mix: need positive sizes and at least 2x2 pixels (reflect padding)");
    const int64_t slabs = static_cast<int64_t>(n_frames) * n_channels;
    if (slabs > 65535) return fail(CLAIR_E_LIMIT, "clair_dark_field_mix: n_frames*n_channels exceeds 65535");
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8, static_cast<unsigned>(slabs));
    if (grid.y > 65535) return fail(CLAIR_E_LIMIT, "clair_dark_field_mix: image too tall");
    dark_mix_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(val_dev, std_dev, dark_dev, dark_std_dev, height, width,
                                                                           threshold, alpha, val_out_dev, std_out_dev);
    return launched("dark_mix_kernel");
}

extern "C" int clair_flat_field_correct(void *value_dev, int value_f64, float *sigma_dev, const float *flat_dev,
                                        const float *flat_std_dev, int n_images, int n_channels, int64_t plane, int mean_in_graph,
                                        double *scratch_dev, void *stream) {
    if (!value_dev || !flat_dev || !scratch_dev) return fail(CLAIR_E_ARG, "clair_flat_field_correct: null buffer");
    if (n_images <= 0 || n_channels <= 0 || n_channels > CLAIR_MAX_CHANNELS || plane <= 0)
        return fail(CLAIR_E_ARG, "clair_flat_field_correct: bad geometry");
    if (mean_in_graph && n_images != 1)
        return fail(CLAIR_E_MODE, "clair_flat_field_correct: the in-graph mean term is defined for a single image (hdr_merge)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaError_t e = cudaMemsetAsync(scratch_dev, 0, sizeof(double) * 2 * n_channels, s); e != cudaSuccess)
        return fail_cuda(e, "cudaMemsetAsync(scratch)");
    const unsigned rblocks = static_cast<unsigned>(std::min<int64_t>((plane + 255) / 256, 4 * device_sm_count()));
    dim3 rgrid(rblocks, n_channels), agrid(static_cast<unsigned>((plane + 255) / 256), n_channels, n_images);
    if (value_f64) {
        flat_reduce_kernel<double><<<rgrid, 256, 0, s>>>(mean_in_graph ? static_cast<const double *>(value_dev) : nullptr, flat_dev, plane, scratch_dev);
        flat_apply_kernel<double><<<agrid, 256, 0, s>>>(static_cast<double *>(value_dev), sigma_dev, flat_dev, flat_std_dev, plane,
                                                         n_channels, mean_in_graph, scratch_dev);
    } else {
        flat_reduce_kernel<float><<<rgrid, 256, 0, s>>>(mean_in_graph ? static_cast<const float *>(value_dev) : nullptr, flat_dev, plane, scratch_dev);
        flat_apply_kernel<float><<<agrid, 256, 0, s>>>(static_cast<float *>(value_dev), sigma_dev, flat_dev, flat_std_dev, plane,
                                                        n_channels, mean_in_graph, scratch_dev);
    }
    if (int rc = launched("flat_reduce_kernel")) return rc;
    return launched("flat_apply_kernel");
}
