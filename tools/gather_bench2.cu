// Microbenchmark 2: random 16-byte gathers confined to a window of W MiB (L2-resident when W is small),
// streaming index reads and coalesced 16-byte writes as in the BP sweep.  MLP per thread = U.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int U>
__global__ void gather_kernel(const uint4 *__restrict__ src, const unsigned *__restrict__ idx, uint4 *__restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride * U) {
        unsigned r[U]; uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = (i + u * stride < n) ? idx[i + u * stride] : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(src + r[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n) dst[i + u * stride] = v[u];
    }
}

int main(int argc, char **argv) {
    const size_t n = 1u << 27;  // gathers
    uint4 *src, *dst; unsigned *idx;
    CK(cudaMalloc(&src, (size_t)2048 << 20)); CK(cudaMalloc(&dst, n * 16)); CK(cudaMalloc(&idx, n * 4));
    CK(cudaMemset(src, 1, (size_t)2048 << 20));
    std::vector<unsigned> h(n);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int wmb : {8, 16, 32, 64, 128, 512, 2048}) {
        const size_t slots = ((size_t)wmb << 20) / 16;
        // window-local random positions: consecutive gathers walk through the buffer window by window
        // (like the bucketed BP layout: all gathers of a time window land in one W-MiB region)
        const size_t per_window = slots;  // as many gathers as slots per window
        for (size_t i = 0; i < n; ++i) {
            const size_t w = (i / per_window) % (((size_t)2048 << 20) / 16 / slots);
            h[i] = (unsigned)(w * slots + ((i * 2654435761ull + 12345) & (slots - 1)));
        }
        CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
        for (int U : {1, 4}) {
            auto launch = [&]() { if (U == 1) gather_kernel<1><<<148 * 8, 256>>>(src, idx, dst, n); else gather_kernel<4><<<148 * 8, 256>>>(src, idx, dst, n); };
            launch(); CK(cudaDeviceSynchronize());
            cudaEventRecord(a); for (int it = 0; it < 3; ++it) launch(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
            printf("window %4d MiB  U=%d  %.3f ms  %.2f Ggather/s\n", wmb, U, ms, n / ms * 1e-6);
        }
    }
    return 0;
}
