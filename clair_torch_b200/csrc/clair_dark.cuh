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
//   row_group_blur   neighbour column sums by shuffle, separable blur.  All 32 lanes of the warp must call it together.
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

// The blur is evaluated in its separable form: the three rows are combined first (per column of the lane's group, and
// for the two single neighbour pixels of an unchained end lane), the left / right neighbour COLUMN SUMS come from the
// adjacent lanes (two shuffles instead of six), then the three columns are combined: 3 (VEC + 2) + 3 VEC operations
// instead of 9 VEC multiply-adds and 6 shuffles.  Reflection commutes with the column sums.
template <int VEC>
__device__ __forceinline__ void row_group_blur(const RowGroup<VEC> &rg, int col, const DarkGeometry &g, bool chained_left,
                                               bool chained_right, float (&center)[VEC], float (&blur)[VEC]) {
    constexpr float t0 = blur_tap(0), t1 = blur_tap(1);
    float v[VEC + 2];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        center[k] = rg.seg[1].v[k];
        v[k + 1] = fmaf(t0, rg.seg[0].v[k] + rg.seg[2].v[k], t1 * rg.seg[1].v[k]);
    }
    const float vl = fmaf(t0, rg.edge_l[0] + rg.edge_l[2], t1 * rg.edge_l[1]);
    const float vr = fmaf(t0, rg.edge_r[0] + rg.edge_r[2], t1 * rg.edge_r[1]);
    float left = __shfl_up_sync(0xffffffffu, v[VEC], 1);
    float right = __shfl_down_sync(0xffffffffu, v[1], 1);
    if (!chained_left) left = (col > 0) ? vl : v[2];                           // col == 0: x[-1] = x[1]
    if (!chained_right) right = (col + VEC < g.W) ? vr : v[VEC - 1];           // row end: x[W] = x[W-2]
    v[0] = left;
    v[VEC + 1] = right;
#pragma unroll
    for (int k = 0; k < VEC; ++k) blur[k] = fmaf(t0, v[k] + v[k + 2], t1 * v[k + 1]);
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

__device__ __forceinline__ float min_keep_nan(float a, float b) {     // a NaN dark value stays a NaN, as through torch.sigmoid
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// The same for two adjacent pixels on packed fp32x2 arithmetic (the fused merge kernel is instruction-bound).  The
// exponent is clamped at 126 instead of testing for overflow afterwards: the sigmoid is then 2^-126, which is 0 for
// every use made of it.
template <bool HAS_STD>
__device__ __forceinline__ void dark_mix_value2(f32x2 x, f32x2 blur, f32x2 s, f32x2 dark, f32x2 dark_std, const DarkGeometry &g,
                                                f32x2 &x_out, f32x2 &s_out) {
    float a0, a1;
    unpack2(mul2(splat2(g.neg_alpha_log2e), sub2(dark, splat2(g.threshold))), a0, a1);
    const f32x2 e = pack2(exp2f_approx(min_keep_nan(a0, 126.0f)), exp2f_approx(min_keep_nan(a1, 126.0f)));
    const f32x2 d = add2(e, splat2(1.0f));
    const f32x2 nd = fma2(e, splat2(-1.0f), splat2(-1.0f));                    // -(1 + e), same rounding
    float d0, d1;
    unpack2(d, d0, d1);
    f32x2 m = pack2(rcp_approx(d0), rcp_approx(d1));
    m = fma2(fma2(nd, m, splat2(1.0f)), m, m);
    const f32x2 om = sub2(splat2(1.0f), m);
    x_out = fma2(m, blur, mul2(om, x));
    if constexpr (HAS_STD) {
        const f32x2 t = mul2(mul2(mul2(mul2(sub2(blur, x), splat2(g.alpha)), m), om), dark_std);
        float q0, q1;
        unpack2(fma2(s, s, mul2(t, t)), q0, q1);
        s_out = pack2(sqrt_approx(q0), sqrt_approx(q1));
    }
}

}  // namespace clair
