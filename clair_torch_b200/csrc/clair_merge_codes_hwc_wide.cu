// As clair_merge_codes_wide.cu, for the interleaved (n_frames, H, W, 3) BGR camera layout.
#include "clair_merge.cuh"

namespace clair {

int launch_merge_codes_hwc_wide(const MergeLaunch &m, bool u8) {
    return u8 ? launch_merge_codes_wide<kSrcU8Hwc>(m) : launch_merge_codes_wide<kSrcU16Hwc>(m);
}

// the same with the block's codes staged in shared memory by bulk copies (clair_merge.cuh: TMA)
int launch_merge_codes_hwc_wide_tma(const MergeLaunch &m, bool u8) {
    return u8 ? launch_merge_codes_wide<kSrcU8HwcTma>(m) : launch_merge_codes_wide<kSrcU16HwcTma>(m);
}

}  // namespace clair
