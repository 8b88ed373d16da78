// Integer-ingest instantiations of the HDR-merge kernels, planar (n_frames, C, plane) codes — SURVEY.md 8(f) rank 2.
// A translation unit of its own so that it compiles in parallel with clair_stack.cu.
#include "clair_merge.cuh"

namespace clair {

int launch_merge_codes_planar(const MergeLaunch &m, bool u8) {
    return u8 ? launch_merge_by_std<4, kSrcU8>(m) : launch_merge_by_std<4, kSrcU16>(m);
}

}  // namespace clair
