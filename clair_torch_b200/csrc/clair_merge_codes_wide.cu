// Integer-ingest instantiations of the register HDR-merge kernel for 9..16 frames, planar codes, 2 codes per thread
// (BASELINE config c4 handed over as uint16 codes).  A translation unit of its own: 64 unrolled kernels.
#include "clair_merge.cuh"

namespace clair {

int launch_merge_codes_planar_wide(const MergeLaunch &m, bool u8) {
    return u8 ? launch_merge_codes_wide<kSrcU8>(m) : launch_merge_codes_wide<kSrcU16>(m);
}

}  // namespace clair
