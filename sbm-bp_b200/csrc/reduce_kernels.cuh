// Non-template reduction kernels of the engine (included by engine.cu only): fixed-order column sums, node
// statistics (EM expectations, overlap confusion matrix), the non-edge pair terms and the moment tensors.
#pragma once
#include "bp_device.cuh"

namespace sbmbp {

// out[c] = sum_b partial[b][c] in a fixed order; one CTA per column
__global__ void __launch_bounds__(kThreads) reduce_cols_kernel(const double *__restrict__ partial, unsigned nrows,
                                                               unsigned ncols, double *__restrict__ out) {
    __shared__ double sred[kThreads / 32];
    const unsigned c = blockIdx.x;
    double p = 0.0;
    for (unsigned b = threadIdx.x; b < nrows; b += kThreads) p += partial[size_t(b) * ncols + c];
    const double v = block_sum(p, sred);
    if (threadIdx.x == 0) out[c] = v;
}

// ---- node-only reductions.  Row layout of partial: [na(Q) | nna(Q) | conf(Q*Q)], stride 2*kMaxQ + kMaxQ*kMaxQ.
// na_q = sum_i psi_i^q, nna_q = sum_i d_i psi_i^q (:428-440); conf[t][q] = sum_{i: true_i = t} psi_i^q, from which
// the host takes the maximum over label permutations (:775-811).
constexpr int kNodeCols = 2 * kMaxQ + kMaxQ * kMaxQ;

__global__ void __launch_bounds__(kThreads) node_stats_kernel(const double *__restrict__ marg,
                                                              const unsigned long long *__restrict__ row_ptr,
                                                              const unsigned *__restrict__ true_conf, unsigned N,
                                                              unsigned Q, double *__restrict__ partial) {
    extern __shared__ double sconf[];  // [kThreads/32][Q*Q] warp-private confusion sums
    __shared__ double sred[kThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (unsigned i = tid; i < (kThreads / 32) * Q * Q; i += kThreads) sconf[i] = 0.0;
    __syncthreads();
    double na[kMaxQ], nna[kMaxQ];
    for (unsigned q = 0; q < Q; ++q) na[q] = nna[q] = 0.0;
    const unsigned per = (N + gridDim.x - 1) / gridDim.x;
    const unsigned lo = blockIdx.x * per, hi = min(N, lo + per);
    for (unsigned base = lo; base < hi; base += kThreads) {
        const unsigned i = base + tid;
        const bool live = i < hi;
        const double d = live ? double(row_ptr[i + 1] - row_ptr[i]) : 0.0;
        const unsigned t = (live && true_conf) ? true_conf[i] : 0xffffffffu;
        for (unsigned q = 0; q < Q; ++q) {
            const double p = live ? marg[size_t(i) * Q + q] : 0.0;
            na[q] += p;
            nna[q] += d * p;
            if (true_conf) {
                // warp-private, lane-serialised accumulation keeps the order fixed
                for (unsigned tt = 0; tt < Q; ++tt) {
                    const double v = warp_sum((t == tt) ? p : 0.0);
                    if (lane == 0) sconf[(warp * Q + tt) * Q + q] += v;
                }
            }
        }
    }
    double *row = partial + size_t(blockIdx.x) * kNodeCols;
    for (unsigned q = 0; q < Q; ++q) {
        double r = block_sum(na[q], sred);
        if (tid == 0) row[q] = r;
        r = block_sum(nna[q], sred);
        if (tid == 0) row[kMaxQ + q] = r;
    }
    __syncthreads();
    for (unsigned i = tid; i < Q * Q; i += kThreads) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += sconf[w * Q * Q + i];
        row[2 * kMaxQ + (i / Q) * kMaxQ + (i % Q)] = s;
    }
}

// ---- non-edge terms, exact: sums over ALL ordered pairs (i, l), l == i included.  The pairs that ARE edges
// are subtracted afterwards by nonedge_edges_kernel (the reference skips them with std::find, :680).
//   mode 0 (free energy, :685-699): log(psi_i^T A psi_l) with A = (1 - c/N)^beta, exact zeros skipped as :697 does
//   mode 1 (series form of mode 0): log1p(-psi_i^T A psi_l) with A = 1 - (1 - c/N)^beta   [edges kernel only]
//   mode 2 (entropy, :719-731): (psi_i^T A psi_l) / (psi_i^T B psi_l), A = (c/N) log c, B = 1 - c/N, skipped when
//          the product of the two is 0 as :730 does
// grid.x tiles i, grid.y tiles l; one partial per CTA.
constexpr int kPairTile = 256;

__device__ __forceinline__ double pair_term(int mode, double fa, double fb) {
    if (mode == 0) return (fa != 0.0) ? log(fa) : 0.0;
    if (mode == 1) return log1p(-fa);
    return (fa * fb != 0.0) ? fa / fb : 0.0;
}

__global__ void __launch_bounds__(kThreads) nonedge_pairs_kernel(const double *__restrict__ marg, unsigned N,
                                                                 unsigned Q, const double *__restrict__ A,
                                                                 const double *__restrict__ B, int mode,
                                                                 double *__restrict__ partial) {
    extern __shared__ double sm[];  // psi_l tile [kPairTile][Q], then A [Q*Q], B [Q*Q]
    __shared__ double sred[kThreads / 32];
    double *spsi = sm;
    double *sA = sm + size_t(kPairTile) * Q;
    double *sB = sA + Q * Q;
    const int tid = threadIdx.x;
    for (unsigned i = tid; i < Q * Q; i += kThreads) {
        sA[i] = A[(i / Q) * kMaxQ + (i % Q)];
        sB[i] = B ? B[(i / Q) * kMaxQ + (i % Q)] : 0.0;
    }
    const unsigned l0 = blockIdx.y * kPairTile;
    for (unsigned i = tid; i < kPairTile * Q; i += kThreads) {
        const unsigned l = l0 + i / Q;
        spsi[i] = (l < N) ? marg[size_t(l) * Q + i % Q] : 0.0;
    }
    __syncthreads();
    const unsigned i = blockIdx.x * kThreads + tid;
    double acc = 0.0;
    if (i < N) {
        double ua[kMaxQ], ub[kMaxQ];  // u[q2] = sum_q1 A[q1][q2] psi_i[q1]
        for (unsigned q2 = 0; q2 < Q; ++q2) {
            double s = 0.0, t = 0.0;
            for (unsigned q1 = 0; q1 < Q; ++q1) {
                const double p = marg[size_t(i) * Q + q1];
                s += sA[q1 * Q + q2] * p;
                t += sB[q1 * Q + q2] * p;
            }
            ua[q2] = s;
            ub[q2] = t;
        }
        const unsigned lim = min(unsigned(kPairTile), N - l0);
        for (unsigned l = 0; l < lim; ++l) {
            double fa = 0.0, fb = 0.0;
            for (unsigned q2 = 0; q2 < Q; ++q2) {
                fa += ua[q2] * spsi[l * Q + q2];
                fb += ub[q2] * spsi[l * Q + q2];
            }
            acc += pair_term(mode, fa, fb);
        }
    }
    const double r = block_sum(acc, sred);
    if (tid == 0) partial[size_t(blockIdx.y) * gridDim.x + blockIdx.x] = r;
}

// the same pair term summed over the directed edges (i, l = col[e]) only
// marg: marginals of the rows (local nodes); marg_nb: marginals indexed by col[] (the same array on one GPU, the
// all-gathered global array on several)
__global__ void __launch_bounds__(kThreads) nonedge_edges_kernel(const double *__restrict__ marg,
                                                                 const double *__restrict__ marg_nb,
                                                                 const unsigned long long *__restrict__ row_ptr,
                                                                 const unsigned *__restrict__ col, unsigned N,
                                                                 unsigned Q, const double *__restrict__ A,
                                                                 const double *__restrict__ B, int mode,
                                                                 double *__restrict__ partial) {
    extern __shared__ double sm[];
    __shared__ double sred[kThreads / 32];
    double *sA = sm;
    double *sB = sm + Q * Q;
    const int tid = threadIdx.x;
    for (unsigned i = tid; i < Q * Q; i += kThreads) {
        sA[i] = A[(i / Q) * kMaxQ + (i % Q)];
        sB[i] = B ? B[(i / Q) * kMaxQ + (i % Q)] : 0.0;
    }
    __syncthreads();
    double acc = 0.0;
    const unsigned per = (N + gridDim.x - 1) / gridDim.x;
    const unsigned lo = blockIdx.x * per, hi = min(N, lo + per);
    for (unsigned i = lo + tid; i < hi; i += kThreads) {
        double ua[kMaxQ], ub[kMaxQ];
        for (unsigned q2 = 0; q2 < Q; ++q2) {
            double s = 0.0, t = 0.0;
            for (unsigned q1 = 0; q1 < Q; ++q1) {
                const double p = marg[size_t(i) * Q + q1];
                s += sA[q1 * Q + q2] * p;
                t += sB[q1 * Q + q2] * p;
            }
            ua[q2] = s;
            ub[q2] = t;
        }
        for (unsigned long long e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
            const unsigned l = col[e];
            double fa = 0.0, fb = 0.0;
            for (unsigned q2 = 0; q2 < Q; ++q2) {
                const double p = marg_nb[size_t(l) * Q + q2];
                fa += ua[q2] * p;
                fb += ub[q2] * p;
            }
            acc += pair_term(mode, fa, fb);
        }
    }
    const double r = block_sum(acc, sred);
    if (tid == 0) partial[blockIdx.x] = r;
}

// sum_i log(s_i), s_i = sum_q psi_i^q: what the marginals' normalisation defect (s_i = 1 + O(ulp)) contributes to the
// pair sum  sum_{i,l} log(psi_i^T W psi_l) = sum_{i,l} log(s_i s_l - y_il):  2 N sum_i log s_i.  The moment series
// expands log(1 - y) and would drop it; it is systematic per node, hence O(N ulp) after the sum over partners.
// s_i - 1 is formed without rounding (TwoSum accumulation, then an exact subtraction).
__global__ void __launch_bounds__(kThreads) lognorm_kernel(const double *__restrict__ marg, unsigned N, unsigned Q,
                                                           double *__restrict__ partial) {
    __shared__ double sred[kThreads / 32];
    double acc = 0.0;
    for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < N; i += gridDim.x * kThreads) {
        double s = marg[size_t(i) * Q], c = 0.0;
        for (unsigned q = 1; q < Q; ++q) {
            const double x = marg[size_t(i) * Q + q];
            const double t = __dadd_rn(s, x);
            const double bp = __dsub_rn(t, s);
            c = __dadd_rn(c, __dadd_rn(__dsub_rn(s, __dsub_rn(t, bp)), __dsub_rn(x, bp)));  // TwoSum error term
            s = t;
        }
        acc += log1p(__dadd_rn(__dsub_rn(s, 1.0), c));
    }
    const double r = block_sum(acc, sred);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// moment tensors T_k[idx] = sum_i prod_j psi_i[digit_j(idx)], idx in [0, Q^k), for the series of the non-edge term.
// Each thread owns tensor entries [idx0, idx0+len), len <= 16*kThreads per launch; nodes stream through smem.
// partial: [gridDim.x][len].
__global__ void __launch_bounds__(kThreads) moments_kernel(const double *__restrict__ marg, unsigned N, unsigned Q,
                                                           unsigned order, unsigned idx0, unsigned len,
                                                           double *__restrict__ partial) {
    extern __shared__ double spsi[];  // [kThreads][Q]
    const int tid = threadIdx.x;
    const unsigned per = (N + gridDim.x - 1) / gridDim.x;
    const unsigned lo = blockIdx.x * per, hi = min(N, lo + per);
    constexpr int kOwn = 16;  // len <= kOwn * kThreads
    double acc[kOwn];
#pragma unroll
    for (int j = 0; j < kOwn; ++j) acc[j] = 0.0;
    for (unsigned base = lo; base < hi; base += kThreads) {
        __syncthreads();
        const unsigned cnt = min(unsigned(kThreads), hi - base);
        for (unsigned i = tid; i < cnt * Q; i += kThreads) spsi[i] = marg[size_t(base) * Q + i];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kOwn; ++j) {
            const unsigned idx = tid + j * kThreads;
            if (idx < len) {
                unsigned dig[8];
                unsigned r = idx0 + idx;
                for (unsigned o = 0; o < order; ++o) {
                    dig[o] = r % Q;
                    r /= Q;
                }
                double s = 0.0;
                for (unsigned n = 0; n < cnt; ++n) {
                    double p = 1.0;
                    for (unsigned o = 0; o < order; ++o) p *= spsi[n * Q + dig[o]];
                    s += p;
                }
                acc[j] += s;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
        const unsigned idx = tid + j * kThreads;
        if (idx < len) partial[size_t(blockIdx.x) * len + idx] = acc[j];
    }
}

}  // namespace sbmbp
