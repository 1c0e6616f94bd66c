// Microbenchmark 3: tile-shaped random gathers (a CTA owns 256*U consecutive gathers, thread owns k = u*256 + tid,
// all U loads issued before use) inside L2-sized windows; load flavour and U swept.
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
    if (MODE == 3) { asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
    return *p;
}

template <int U, int MODE>
__global__ void tile_gather(const uint4 *__restrict__ src, const unsigned *__restrict__ idx, uint4 *__restrict__ dst, size_t n) {
    const size_t base = (size_t)blockIdx.x * 256 * U;
    unsigned r[U]; uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = idx[base + u * 256 + threadIdx.x];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld16<MODE>(src + r[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) dst[base + u * 256 + threadIdx.x] = v[u];
}

template <int U, int MODE>
void run(const char *nm, const uint4 *src, const unsigned *idx, uint4 *dst, size_t n, int wmb) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const unsigned grid = (unsigned)(n / (256 * U));
    tile_gather<U, MODE><<<grid, 256>>>(src, idx, dst, n); CK(cudaDeviceSynchronize());
    cudaEventRecord(a); for (int it = 0; it < 3; ++it) tile_gather<U, MODE><<<grid, 256>>>(src, idx, dst, n); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("window %4d MiB  U=%d %-12s %.3f ms  %.2f Ggather/s\n", wmb, U, nm, ms, n / ms * 1e-6);
}

int main() {
    const size_t n = 1u << 27;
    uint4 *src, *dst; unsigned *idx;
    CK(cudaMalloc(&src, (size_t)2048 << 20)); CK(cudaMalloc(&dst, n * 16)); CK(cudaMalloc(&idx, n * 4));
    CK(cudaMemset(src, 1, (size_t)2048 << 20));
    std::vector<unsigned> h(n);
    for (int wmb : {16, 32, 2048}) {
        const size_t slots = ((size_t)wmb << 20) / 16, nwin = (((size_t)2048 << 20) / 16) / slots;
        for (size_t i = 0; i < n; ++i) h[i] = (unsigned)(((i / slots) % nwin) * slots + ((i * 2654435761ull + 12345) & (slots - 1)));
        CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
        run<1, 0>("ldg", src, idx, dst, n, wmb); run<2, 0>("ldg", src, idx, dst, n, wmb); run<4, 0>("ldg", src, idx, dst, n, wmb); run<8, 0>("ldg", src, idx, dst, n, wmb);
        run<4, 1>("ldcg", src, idx, dst, n, wmb); run<4, 3>("nc.noalloc", src, idx, dst, n, wmb); run<4, 5>("plain", src, idx, dst, n, wmb);
        run<1, 1>("ldcg", src, idx, dst, n, wmb); run<8, 1>("ldcg", src, idx, dst, n, wmb);
    }
    return 0;
}
