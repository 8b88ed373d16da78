// Integer-ingest instantiations of the HDR-merge kernels for the interleaved (n_frames, H, W, 3) BGR camera layout: the
// reference's CvToTorch transform (common/general_functions.py:315-336) folded into the load.
#include "clair_merge.cuh"

namespace clair {

int launch_merge_codes_hwc(const MergeLaunch &m, bool u8) {
    return u8 ? launch_merge_by_std<4, kSrcU8Hwc>(m) : launch_merge_by_std<4, kSrcU16Hwc>(m);
}

}  // namespace clair
