// Error reporting, argument validation and launch accounting for libclair_b200.so.
#include "clair_host.h"

#include <atomic>
#include <cstdio>
#include <string>

namespace clair {

namespace {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

int fail(int code, const char *msg) {
    std::snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
    return code;
}

int fail_cuda(cudaError_t e, const char *what) {
    std::snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return static_cast<int>(e);
}

int launched(const char *kernel_name) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(e, kernel_name);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int check_geometry(const char *fn, int n_frames, int n_channels, int64_t plane, int lut_size, bool limit_frames) {
    char buf[256];
    if (n_frames <= 0 || n_channels <= 0 || plane <= 0 || lut_size < 2) {
        std::snprintf(buf, sizeof(buf), "%s: n_frames, n_channels, plane must be positive and lut_size >= 2", fn);
        return fail(CLAIR_E_ARG, buf);
    }
    if (n_channels > CLAIR_MAX_CHANNELS || lut_size > CLAIR_MAX_LUT || (limit_frames && n_frames > CLAIR_MAX_FRAMES)) {
        std::snprintf(buf, sizeof(buf), "%s: limit exceeded (channels <= %d, lut_size <= %d, frames per batch <= %d)", fn,
                      CLAIR_MAX_CHANNELS, CLAIR_MAX_LUT, CLAIR_MAX_FRAMES);
        return fail(CLAIR_E_LIMIT, buf);
    }
    if (plane >= (int64_t(1) << 31)) {
        std::snprintf(buf, sizeof(buf), "%s: H*W must be below 2^31", fn);
        return fail(CLAIR_E_LIMIT, buf);
    }
    return 0;
}

void fill_rows(CurveRows &rows, const int32_t *curve_row_base_host, int n_channels, int64_t plane) {
    static_assert(CLAIR_MAX_CHANNELS <= 8, "CurveRows packs 4 bits per channel into 32 bits");
    rows.packed = 0;
    for (int c = 0; c < n_channels; ++c) {
        uint32_t b;
        if (curve_row_base_host != nullptr) {
            b = static_cast<uint32_t>(((curve_row_base_host[c] % n_channels) + n_channels) % n_channels);
        } else {
            b = static_cast<uint32_t>((static_cast<int64_t>(c) * plane) % n_channels);
        }
        rows.packed |= b << (4 * c);
    }
}

Tuning g_tuning;

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

}  // namespace clair

extern "C" int clair_set_tuning(const char *key, int value) {
    using clair::g_tuning;
    const std::string k(key ? key : "");
    if (k == "hdr_vec") g_tuning.hdr_vec = value;
    else if (k == "hdr_waves") g_tuning.hdr_waves = value;
    else if (k == "hdr_force_dynamic") g_tuning.hdr_force_dynamic = value;
    else if (k == "hdr_prefetch") g_tuning.hdr_prefetch = value;
    else if (k == "stats_blocks_per_sm") g_tuning.stats_blocks_per_sm = value;
    else if (k == "grad_blocks_per_sm") g_tuning.grad_blocks_per_sm = value;
    else if (k == "grad_pix") g_tuning.grad_pix = value;
    else if (k == "hdr_fixed_max") g_tuning.hdr_fixed_max = value;
    else if (k == "grad_warps") g_tuning.grad_warps = value;
    else if (k == "grad_copies") g_tuning.grad_copies = value;
    else if (k == "hdr_tma") g_tuning.hdr_tma = value;
    else if (k == "fwd_blocks") g_tuning.fwd_blocks = value;
    else if (k == "dark_strip") g_tuning.dark_strip = value;
    else if (k == "dark_rows") g_tuning.dark_rows = value;
    else if (k == "aux_waves") g_tuning.aux_waves = value;
    else if (k == "stats_waves") g_tuning.stats_waves = value;
    else if (k == "grad_waves") g_tuning.grad_waves = value;
    else if (k == "stats_warps") g_tuning.stats_warps = value;
    else if (k == "stats_buffers") g_tuning.stats_buffers = value;
    else if (k == "stats_slots") g_tuning.stats_slots = value;
    else return clair::fail(CLAIR_E_ARG, "clair_set_tuning: unknown key");
    return 0;
}

extern "C" int clair_abi_version(void) { return CLAIR_ABI_VERSION; }
extern "C" const char *clair_last_error(void) { return clair::g_last_error; }
extern "C" uint64_t clair_launch_count(void) { return clair::g_launches.load(std::memory_order_relaxed); }
