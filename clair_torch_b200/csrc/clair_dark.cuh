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

// Two phases, so that a caller can put the loads of several planes in flight before the first shuffle (the compiler
// does not move loads across warp-synchronous instructions):
//   row_group_load   three row segments of VEC pixels (rows above / at / below, reflect padding) plus, for a warp's end
//                    lanes only, the single neighbour pixels the adjacent lanes cannot supply (predicated loads)
//   row_group_blur   neighbours by shuffle, 9-tap blur.  All 32 lanes of the warp must call it together.
//   plane: first element of the (H, W) plane;  row, col: position of the lane's first pixel (col % VEC == 0, W % VEC == 0)
//   active: the lane owns real pixels;  chained_*: the lane to the left / right holds the adjacent pixels of the same row
template <int VEC>
struct RowGroup {
    Pack<VEC> seg[3];
    float edge_l[3], edge_r[3];
};

template <int VEC>
__device__ __forceinline__ void row_group_load(const float *__restrict__ plane, int row, int col, const DarkGeometry &g, bool active,
                                               bool chained_left, bool chained_right, RowGroup<VEC> &rg) {
    static_assert(VEC == 2 || VEC == 4, "row groups are 2 or 4 pixels wide");
    const int rows[3] = {row == 0 ? 1 : row - 1, row, row == g.H - 1 ? g.H - 2 : row + 1};     // reflect padding
    const bool need_l = active && !chained_left && col > 0, need_r = active && !chained_right && col + VEC < g.W;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float *seg = plane + static_cast<int64_t>(rows[j]) * g.W + col;
        if (active) {
            rg.seg[j] = load_stream<VEC>(seg);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) rg.seg[j].v[k] = 0.0f;
        }
        rg.edge_l[j] = need_l ? __ldg(seg - 1) : 0.0f;
        rg.edge_r[j] = need_r ? __ldg(seg + VEC) : 0.0f;
    }
}

template <int VEC>
__device__ __forceinline__ void row_group_blur(const RowGroup<VEC> &rg, int col, const DarkGeometry &g, bool chained_left,
                                               bool chained_right, float (&center)[VEC], float (&blur)[VEC]) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) { center[k] = rg.seg[1].v[k]; blur[k] = 0.0f; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float left = __shfl_up_sync(0xffffffffu, rg.seg[j].v[VEC - 1], 1);
        float right = __shfl_down_sync(0xffffffffu, rg.seg[j].v[0], 1);
        if (!chained_left) left = (col > 0) ? rg.edge_l[j] : rg.seg[j].v[1];                     // col == 0: x[-1] = x[1]
        if (!chained_right) right = (col + VEC < g.W) ? rg.edge_r[j] : rg.seg[j].v[VEC - 2];     // row end: x[W] = x[W-2]
        float ext[VEC + 2];
        ext[0] = left;
#pragma unroll
        for (int k = 0; k < VEC; ++k) ext[k + 1] = rg.seg[j].v[k];
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
        s_out = sqrt_approx(fmaf(s, s, t * t));
    }
}

}  // namespace clair
