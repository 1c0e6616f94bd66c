// Microbenchmark: what does a random small gather cost on B200?  (tuning aid, not part of the product)
// Random permutation index -> gather G bytes per element from a 2 GiB buffer -> coalesced 16-byte write.
// Run plain for timings, under ncu for dram__bytes / lts sectors per variant (kernel names carry the variant).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint4 *p) {
    uint4 r;
    if (MODE == 0) return __ldg(p);
    if (MODE == 1) return __ldcg(p);
    if (MODE == 2) return __ldcs(p);
    if (MODE == 3) { asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
    if (MODE == 4) { asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(0x12F0000000000000ull)); return r; }
    return *p;
}

// G = bytes gathered per element (16, 32, 64, 128), contiguous and G-aligned
template <int MODE, int G>
__global__ void gather_kernel(const uint4 *__restrict__ src, const unsigned *__restrict__ idx, uint4 *__restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned r = idx[i];
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < G / 16; ++k) {
            uint4 v = ld16<MODE>(src + (size_t)r * (G / 16) + k);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        dst[i] = acc;
    }
}

template <int MODE, int G>
void run(const char *name, const uint4 *src, const unsigned *idx, uint4 *dst, size_t n) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<MODE, G><<<148 * 8, 256>>>(src, idx, dst, n);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int it = 0; it < 3; ++it) gather_kernel<MODE, G><<<148 * 8, 256>>>(src, idx, dst, n);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("%-28s G=%3d  %.3f ms  %.2f Ggather/s  useful %.0f GB/s\n", name, G, ms, n / ms * 1e-6, n * (double)(G + 16 + 4) / ms * 1e-6);
}

int main(int argc, char **argv) {
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit: %zu\n", g);
    const size_t n = 1u << 27;            // 128 Mi elements
    const size_t src_bytes = n * 16;      // 2 GiB
    uint4 *src, *dst; unsigned *idx;
    CK(cudaMalloc(&src, src_bytes)); CK(cudaMalloc(&dst, n * 16)); CK(cudaMalloc(&idx, n * 4));
    CK(cudaMemset(src, 1, src_bytes));
    std::vector<unsigned> h(n);
    // random permutation by LCG-based multiplicative hashing (full period over 2^27)
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i * 2654435761ull + 12345) & (n - 1));
    // multiplicative by odd constant mod 2^27 is a bijection
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 16>("ldg(nc) 16B", src, idx, dst, n);
    run<1, 16>("ldcg 16B", src, idx, dst, n);
    run<2, 16>("ldcs 16B", src, idx, dst, n);
    run<3, 16>("nc.L1::no_allocate 16B", src, idx, dst, n);
    run<5, 16>("plain 16B", src, idx, dst, n);
    // wider gathers: indices reduced so that r*(G/16) stays in range
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i * 2654435761ull + 12345) & (n / 2 - 1));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 32>("ldg(nc) 32B", src, idx, dst, n / 2);
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i * 2654435761ull + 12345) & (n / 4 - 1));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 64>("ldg(nc) 64B", src, idx, dst, n / 4);
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned)((i * 2654435761ull + 12345) & (n / 8 - 1));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 128>("ldg(nc) 128B", src, idx, dst, n / 8);
    // sequential reference
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned)i;
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 16>("sequential 16B", src, idx, dst, n);
    return 0;
}
