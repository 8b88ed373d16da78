// Micro-benchmark: throughput of shared-memory accumulation variants on sm_100a (random bins, 768 x 2 floats).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 d; asm("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint32_t rng(uint32_t &s){ s = s * 1664525u + 1013904223u; return s >> 8; }
constexpr int BINS = 774;   // pairs
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters, int spread) {
    __shared__ __align__(8) float h[2 * BINS];
    for (int i = threadIdx.x; i < 2 * BINS; i += blockDim.x) h[i] = 0.f;
    __syncthreads();
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    for (int it = 0; it < iters; ++it) {
        const uint32_t bin = rng(s) % spread;
        const float a = 1.0f, b = 0.5f;
        if (MODE == 0) {            // two fp32 atomicAdd (compiler's CAST.SPIN loops)
            atomicAdd(&h[2 * bin], a); atomicAdd(&h[2 * bin + 1], b);
        } else if (MODE == 1) {     // one 64-bit CAS loop
            u64 *p = reinterpret_cast<u64 *>(&h[2 * bin]);
            u64 old = *p, assumed; const u64 inc = pack2(a, b);
            do { assumed = old; old = atomicCAS(p, assumed, add2(assumed, inc)); } while (old != assumed);
        } else if (MODE == 2) {     // native int32 add x2
            atomicAdd(reinterpret_cast<int *>(&h[2 * bin]), 3); atomicAdd(reinterpret_cast<int *>(&h[2 * bin + 1]), 5);
        } else if (MODE == 3) {     // one fp32 atomicAdd
            atomicAdd(&h[2 * bin], a);
        } else if (MODE == 4) {     // u64 atomicAdd (CAST.SPIN.64)
            atomicAdd(reinterpret_cast<u64 *>(&h[2 * bin]), 0x0000000100000001ull);
        } else if (MODE == 6) {     // 64-bit fixed point: returning 32-bit add on the low word, carry / borrow into the high word
            unsigned *w = reinterpret_cast<unsigned *>(h);
            const float v = (rng(s) & 1) ? -a * 1000.5f : a * 777.25f;
            const long long x = static_cast<long long>(__float2ll_rn(v * 1048576.0f));
            const unsigned lo = static_cast<unsigned>(x);
            const unsigned old = atomicAdd(&w[2 * bin], lo);
            const int hi = static_cast<int>(x >> 32) + ((old + lo < old) ? 1 : 0);
            if (hi != 0) atomicAdd(reinterpret_cast<int *>(&w[2 * bin + 1]), hi);
        } else if (MODE == 5) {     // global red v2
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(out + 2 * ((bin + blockIdx.x % 64 * BINS))), "f"(a), "f"(b));
        }
    }
    __syncthreads();
    if (MODE != 5) for (int i = threadIdx.x; i < 2 * BINS; i += blockDim.x) if (h[i] != 0.f) atomicAdd(out + i, h[i]);
}
template <int MODE> void run(const char *name, float *out, int spread) {
    const int blocks = 148 * 8, iters = 2000;
    k<MODE><<<blocks, 128>>>(out, 10, spread);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, 128>>>(out, iters, spread);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s spread %4d: %8.3f ms  %8.1f G updates/s  (%s)\n", name, spread, ms, double(blocks) * 128 * iters / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out; cudaMalloc(&out, sizeof(float) * 2 * BINS * 64); cudaMemset(out, 0, sizeof(float) * 2 * BINS * 64);
    for (int spread : {768, 64, 16}) {
        run<0>("2x atomicAdd(float) smem", out, spread);
        run<1>("CAS.64 loop smem", out, spread);
        run<2>("2x atomicAdd(int) smem", out, spread);
        run<3>("1x atomicAdd(float) smem", out, spread);
        run<4>("atomicAdd(u64) smem", out, spread);
        run<5>("red.global.add.v2.f32", out, spread);
        run<6>("fixed-point lo/hi int smem", out, spread);
    }
    return 0;
}
