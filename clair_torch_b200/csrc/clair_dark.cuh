// Dark-field mix of one frame-channel plane, evaluated for VEC adjacent pixels of one row per thread: the load stage
// shared by the stand-alone pre-pass (clair_artefacts.cu) and the fused HDR merge (clair_stack.cu).
//     m = sigmoid(alpha (D - threshold)),  x' = m B(x) + (1 - m) x,  s_eff = sqrt(s^2 + ((B(x) - x) alpha m (1 - m) s_D)^2)
// (common/general_functions.py:440-486 as the drivers use it, see clair_artefacts.cu).  B is torchvision's 3x3 Gaussian
// blur with reflect padding.  Threads of a warp hold consecutive pixel groups, so the left / right neighbour columns
// come from the adjacent lanes by shuffle; only a warp's two end lanes (and row ends) touch memory for them, and the
// rows above / below are the only extra loads: 3 vector loads per VEC pixels instead of 9 scalar ones per pixel.
#pragma once

#include "clair_common.cuh"

namespace clair {

// torchvision GaussianBlur(kernel_size=3, sigma=1): exp(-0.5 d^2) / sum, in fp32
__device__ __forceinline__ constexpr float blur_tap(int i) { return i == 1 ? 0.45186276f : 0.27406862f; }

struct DarkGeometry {
    int H, W;
    float threshold, alpha;
    float neg_alpha_log2e;      // -alpha * log2(e)
};

// All 32 lanes of the warp must call this together (`active` = this lane owns real pixels).
//   plane: first element of the (H, W) plane;  row, col: position of the lane's first pixel (col % VEC == 0, W % VEC == 0)
//   chained: the lane to the left / right holds the adjacent pixels of the same row
template <int VEC>
__device__ __forceinline__ void blur3_row_group(const float *__restrict__ plane, int row, int col, const DarkGeometry &g,
                                                bool active, bool chained_left, bool chained_right, float (&center)[VEC],
                                                float (&blur)[VEC]) {
    static_assert(VEC == 2 || VEC == 4, "row groups are 2 or 4 pixels wide");
    const int rows[3] = {row == 0 ? 1 : row - 1, row, row == g.H - 1 ? g.H - 2 : row + 1};     // reflect padding
    Pack<VEC> seg[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (active) {
            seg[j] = load_stream<VEC>(plane + static_cast<int64_t>(rows[j]) * g.W + col);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) seg[j].v[k] = 0.0f;
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) { center[k] = seg[1].v[k]; blur[k] = 0.0f; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float left = __shfl_up_sync(0xffffffffu, seg[j].v[VEC - 1], 1);
        float right = __shfl_down_sync(0xffffffffu, seg[j].v[0], 1);
        if (!chained_left) {
            left = seg[j].v[1];                                                        // col == 0: x[-1] = x[1]
            if (active && col > 0) left = __ldg(plane + static_cast<int64_t>(rows[j]) * g.W + col - 1);
        }
        if (!chained_right) {
            right = seg[j].v[VEC - 2];                                                 // row end: x[W] = x[W-2]
            if (active && col + VEC < g.W) right = __ldg(plane + static_cast<int64_t>(rows[j]) * g.W + col + VEC);
        }
        float ext[VEC + 2];
        ext[0] = left;
#pragma unroll
        for (int k = 0; k < VEC; ++k) ext[k + 1] = seg[j].v[k];
        ext[VEC + 1] = right;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) blur[k] = fmaf(blur_tap(j) * blur_tap(dx), ext[k + dx], blur[k]);
        }
    }
}

// (x', s_eff) from (x, B(x), s, D, s_D); s_eff is only formed when HAS_STD
template <bool HAS_STD>
__device__ __forceinline__ void dark_mix_value(float x, float blur, float s, float dark, float dark_std, const DarkGeometry &g,
                                               float &x_out, float &s_out) {
    const float e = exp2f_approx(__fmul_rn(g.neg_alpha_log2e, __fsub_rn(dark, g.threshold)));
    const float d = 1.0f + e;
    float m = rcp_approx(d);
    m = fmaf(fmaf(-d, m, 1.0f), m, m);
    if (!(e < 3.0e38f)) m = 0.0f;                          // exp overflow: the sigmoid is 0
    x_out = fmaf(m, blur, (1.0f - m) * x);
    if constexpr (HAS_STD) {
        const float t = (blur - x) * g.alpha * m * (1.0f - m) * dark_std;
        s_out = sqrtf(fmaf(s, s, t * t));
    }
}

}  // namespace clair
