// Host-side helpers shared by the C-ABI entry points (argument checks, error reporting, launch accounting).
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include <algorithm>
#include <initializer_list>

#include "../../include/clair_b200.h"
#include "clair_common.cuh"

namespace clair {

// NVTX range around the host side of an entry point (header-only NVTX 3: a no-op unless a profiler injects itself), so a
// timeline shows merge / statistics / gradient calls by name
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// records `msg` as the calling thread's last error and returns `code`
int fail(int code, const char *msg);
int fail_cuda(cudaError_t e, const char *what);
// checks cudaGetLastError() after a launch and bumps the launch counter
int launched(const char *kernel_name);

int check_geometry(const char *fn, int n_frames, int n_channels, int64_t plane, int lut_size, bool limit_frames);

// curve_row_base_host == NULL: whole image, row of element (c, p) is (c*plane + p) mod C
void fill_rows(CurveRows &rows, const int32_t *curve_row_base_host, int n_channels, int64_t plane);

int device_sm_count();

// Grid x-extent of a persistent kernel launched as (x, n_channels) blocks with `blocks_per_sm` co-resident blocks per
// SM: the largest x whose x * n_channels blocks are ALL resident at once.  Rounded DOWN on purpose — one block more
// than fits waits for a free slot and then runs its full share of the work alone, doubling the kernel's duration.
inline int64_t resident_blocks_per_channel(int blocks_per_sm, int n_channels) {
    return std::max<int64_t>(1, static_cast<int64_t>(device_sm_count()) * std::max(blocks_per_sm, 1) / std::max(n_channels, 1));
}

// Developer tuning knobs (clair_set_tuning); 0 = library default.
struct Tuning {
    int hdr_vec = 0;            // cap the pixels per thread of the HDR-merge kernel (1, 2, 4)
    int hdr_waves = 0;          // resident waves per persistent grid
    int hdr_fixed_max = 0;      // largest N that takes the register kernel (fp32 stacks)
    int hdr_force_dynamic = 0;  // use the N-dynamic float64-sum kernel even for N <= 8
    int fwd_blocks = 0;         // forward / linearise: 1024-pixel tiles per block (-1 = one tile per block, no loop bound sharing)
    int aux_waves = 0;          // dark mix / flat reduce / frame statistics / code expansion: grid = resident blocks x this
    int dark_strip = 0, dark_rows = 0;   // fused dark merge: -1 = grid-stride form instead of the strip walk; rows per band
    int hdr_tma = 0;            // camera-layout 9..16-frame kernels: -1 = per-thread loads instead of the bulk-copy staging
    int hdr_prefetch = 0;       // camera-layout register kernels: prefetch the next trip's codes (1 = into L1, 2 = into L2, -1 = off)
    int stats_waves = 0, grad_waves = 0;   // pair kernels: grid = resident blocks x this
    int stats_blocks_per_sm = 0;
    int stats_warps = 0;        // block shape of the statistics kernel (both must be set)
    int stats_slots = 0;
    int stats_buffers = 0;      // tile buffers of the packed statistics kernel (1 or 2)
    int grad_blocks_per_sm = 0;
    int grad_pix = 0;           // 1 forces one pixel per lane in the gradient kernel
    int grad_warps = 0;         // warps per block of the gradient kernel
    int grad_copies = 0;        // replicated gradient tables the pair-gradient kernel spreads its reductions over
};
extern Tuning g_tuning;

}  // namespace clair
