// Microbenchmark 4: random 16-byte SCATTER stores against random 16-byte GATHER loads over the same window, the question
// behind the destination-major message layout (in-messages contiguous, out-messages scattered).  n accesses to a window
// of W MiB through a random permutation (every slot exactly once, like a BP sweep), warm L2 and flushed L2.
//   gather : dst[k] = src[idx[k]]      (idx streamed, 16-byte coalesced store)
//   scatter: dst[idx[k]] = src[k]      (idx streamed, 16-byte coalesced load)
//   copy   : dst[k] = src[k]           (the streaming floor)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int U, int MODE>
__global__ void kern(const uint4 *__restrict__ src, const unsigned *__restrict__ idx, uint4 *__restrict__ dst, size_t n) {
    const size_t base = (size_t)blockIdx.x * 256 * U;
    unsigned r[U]; uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = idx[base + u * 256 + threadIdx.x];
    if (MODE == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(src + r[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) dst[base + u * 256 + threadIdx.x] = v[u];
    } else if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(src + base + u * 256 + threadIdx.x);
#pragma unroll
        for (int u = 0; u < U; ++u) dst[r[u]] = v[u];
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(src + base + u * 256 + threadIdx.x);
#pragma unroll
        for (int u = 0; u < U; ++u) { v[u].x += r[u]; dst[base + u * 256 + threadIdx.x] = v[u]; }
    }
}

template <int U, int MODE>
void run(const char *nm, const uint4 *src, const unsigned *idx, uint4 *dst, size_t n, int wmb, void *flush, size_t flush_bytes) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const unsigned grid = (unsigned)(n / (256 * U));
    kern<U, MODE><<<grid, 256>>>(src, idx, dst, n); CK(cudaDeviceSynchronize());
    float warm = 0, cold = 0;
    for (int it = 0; it < 5; ++it) { cudaEventRecord(a); kern<U, MODE><<<grid, 256>>>(src, idx, dst, n); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); warm += ms / 5; }
    for (int it = 0; it < 5; ++it) { CK(cudaMemsetAsync(flush, it, flush_bytes)); cudaEventRecord(a); kern<U, MODE><<<grid, 256>>>(src, idx, dst, n); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); cold += ms / 5; }
    printf("window %4d MiB  U=%d %-8s warm %.2f us %.1f G/s   cold %.2f us %.1f G/s\n", wmb, U, nm, warm * 1e3, n / warm * 1e-6, cold * 1e3, n / cold * 1e-6);
}

int main() {
    void *flush; const size_t flush_bytes = (size_t)256 << 20; CK(cudaMalloc(&flush, flush_bytes));
    for (int wmb : {16, 48, 192}) {
        const size_t n = ((size_t)wmb << 20) / 16;
        uint4 *src, *dst; unsigned *idx;
        CK(cudaMalloc(&src, n * 16)); CK(cudaMalloc(&dst, n * 16)); CK(cudaMalloc(&idx, n * 4));
        CK(cudaMemset(src, 1, n * 16)); CK(cudaMemset(dst, 0, n * 16));
        std::vector<unsigned> h(n);
        for (size_t i = 0; i < n; ++i) h[i] = (unsigned)i;
        std::mt19937 g(1); std::shuffle(h.begin(), h.end(), g);
        CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
        run<4, 0>("gather", src, idx, dst, n, wmb, flush, flush_bytes);
        run<4, 1>("scatter", src, idx, dst, n, wmb, flush, flush_bytes);
        run<4, 2>("copy", src, idx, dst, n, wmb, flush, flush_bytes);
        run<8, 0>("gather", src, idx, dst, n, wmb, flush, flush_bytes);
        run<8, 1>("scatter", src, idx, dst, n, wmb, flush, flush_bytes);
        // locality: permutation only within blocks of 2^k slots (a bucketed layout), k = 20 (16 MiB)
        for (size_t i = 0; i < n; ++i) h[i] = (unsigned)i;
        for (size_t b = 0; b < n; b += (1u << 20)) std::shuffle(h.begin() + b, h.begin() + std::min(n, b + (1u << 20)), g);
        CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
        run<4, 0>("gather16M", src, idx, dst, n, wmb, flush, flush_bytes);
        run<4, 1>("scatt16M", src, idx, dst, n, wmb, flush, flush_bytes);
        cudaFree(src); cudaFree(dst); cudaFree(idx);
    }
    return 0;
}
