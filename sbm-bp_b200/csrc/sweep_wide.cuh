// Wide-Q sweep kernel (Q = 32, BASELINE configs[4]): one WARP per node, lane = group index q.
//
// At Q = 32 a message is 256 bytes (FP64) -- two full lines -- so the gather needs no bucketing: messages stay in
// slot order (pos = identity), the old out-messages and the new ones of a node are one contiguous block.  What is
// different from the small-Q kernels is the arithmetic: every in-edge costs a 32 x 32 contraction
//     b_e[q] = sum_t K[t][q] psi_e[t]                                   (belief_propagation.cpp:1002-1012)
// i.e. 1024 FMAs against 788 bytes of traffic.  It is evaluated for 8 in-edges at a time as an 8 x 32 x 32 GEMM:
//   * FP64: on the tensor cores, 32 mma.sync.m8n8k4.f64 per 8 edges (DMMA is the only FP64 tensor path sm_100a has;
//     tcgen05 carries no FP64).  Each lane loads 64 bytes of one edge's message straight into its A fragments -- the
//     k index of the product is permuted (t = 8 (j >> 1) + 2 c + (j & 1) for fragment j of quad lane c) so that those
//     are four fully used 16-byte loads -- and the B fragments of K sit in shared memory in the same permutation.
//   * FP32 storage: FFMA on the CUDA cores, lane = q with K's column in registers and psi broadcast from shared memory
//     (TF32 tensor cores would miss the 1e-5 bar; the FP32 pipe is 4x under the HBM time anyway).
// The b_e of a node (degree <= 32 here) are parked in the warp's shared-memory slab; then, lane = q:
//   node   tot[q] = prod_e b_e[q] * eta_q * F_q, marginal = tot / sum_q (warp shuffle), coalesced 256-byte write
//   edge   cavity[q] = marginal[q] / b_e[q], normalise over q (warp shuffle), max |old - new|, damped write --
//          both the old and the new message are the node's own contiguous rows.
// Nodes of degree > 32 (none to speak of on the Poisson graphs this path is for) go through bp_sweep_fast_kernel
// on a tile list of their own, launched just before; the last CTA of this kernel closes the sweep over all rows.
#pragma once
#include "bp_device.cuh"
#include "sweep_tile.cuh"

namespace sbmbp {

#ifndef SBMBP_WIDE_MINB
#define SBMBP_WIDE_MINB 2
#endif
#ifndef SBMBP_WIDE_DFMA
#define SBMBP_WIDE_DFMA 0  // 1: FP64 contraction on the CUDA cores (DFMA, the FP32 path's structure) -- the comparison build
#endif

constexpr int kWideQ = 32;
constexpr int kWideMaxDeg = 32;   // degrees handled here (product domain: < 50 by construction)
constexpr int kWideRow = 34;      // row stride of the b slab in elements: 16-byte aligned rows, fragments spread over banks

template <typename T>
struct WideSweepArgs {
    const unsigned long long *row_ptr;
    const unsigned *rev;    // slot order: position (= slot) of the message INTO row(e) along e
    const unsigned *nodes;  // nodes of degree <= kWideMaxDeg
    unsigned nnodes;
    T *S[2];
    double *marg;
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    double *partial;       // [rows_before + gridDim.x][33]
    unsigned rows_before;  // rows left by the big-node launch, stored first
    unsigned dc;
    double damping;
};

template <typename T>
struct WideSmem {
    static constexpr int NW = kThreads / 32;
    static constexpr size_t slab = sizeof(T) * kWideMaxDeg * kWideRow;                     // b_e of one node
    static constexpr size_t off_slab = 0;                                                  // [NW][slab]
    static constexpr size_t off_stage = off_slab + NW * slab;                              // CUDA-core path: T[NW][8][32]
    static constexpr bool kDmma = sizeof(T) == 8 && !SBMBP_WIDE_DFMA;
    static constexpr size_t off_kf = off_stage + (kDmma ? 0 : NW * 8 * 32 * sizeof(T));    // DMMA path: double[8][4][32]
    static constexpr size_t bytes = off_kf + 8 * 4 * 32 * sizeof(double);
};

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, SBMBP_WIDE_MINB) bp_sweep_wide_kernel(const WideSweepArgs<T> a) {
    constexpr int QT = kWideQ, NW = kThreads / 32;
    constexpr bool kF64 = sizeof(T) == 8 && !SBMBP_WIDE_DFMA;  // true: the DMMA path
    using Lay = WideSmem<T>;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double s_rows[NW][QT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;  // fragment row (edge of the block) / quad lane
    T *sb = reinterpret_cast<T *>(smem + Lay::off_slab + size_t(warp) * Lay::slab);
    T *stage = reinterpret_cast<T *>(smem + Lay::off_stage) + warp * 8 * 32;
    double *sKf = reinterpret_cast<double *>(smem + Lay::off_kf);

    Ctl *ctl = a.ctl;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;  // uniform over the grid
    const int par = int(sweeps_done & 1u);
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const Field *fld = par ? a.field[1] : a.field[0];
    const bool dc = a.dc != 0;
    const double Nd = a.prm->N;
    const T damp = T(a.damping), keep = T(1.0 - a.damping);
    const double eta_q = a.prm->eta[lane], h_q = fld->h[lane], exph_q = fld->exph[lane];

    // K for the contraction.  FP64: B fragments in shared memory, sKf[(j * 4 + qb) * 32 + lane] =
    // K[t(j, lane & 3)][8 qb + (lane >> 2)] with t(j, c) = 8 (j >> 1) + 2 c + (j & 1).  FP32: column q = lane in registers.
    T kcol[kF64 ? 1 : QT];
    if constexpr (kF64) {
        for (int i = tid; i < 8 * 4 * 32; i += kThreads) {
            const int l = i & 31, qb = (i >> 5) & 3, j = i >> 7;
            const int t = 8 * (j >> 1) + 2 * (l & 3) + (j & 1), q = 8 * qb + (l >> 2);
            sKf[i] = a.prm->Ks[t * kMaxQ + q];
        }
    } else {
#pragma unroll
        for (int t = 0; t < QT; ++t) kcol[t] = T(a.prm->Ks[t * kMaxQ + lane]);
    }
    __syncthreads();

    double wsum = 0.0, mydiff = 0.0;
    const unsigned nwt = gridDim.x * NW;

    // FP64: A fragments of the 8-edge block starting at slot k0 (zeros past the node's degree)
    auto load_afrag = [&](unsigned k0, unsigned d, unsigned g, double (&af)[8]) {
        const unsigned k = k0 + fr;
        const unsigned gk = __shfl_sync(0xffffffffu, g, int(k & 31u));
        if (k < d) {
            const uint4 *src = reinterpret_cast<const uint4 *>(Sold + size_t(gk) * QT + 2 * fc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // 16 bytes at t = 8 i + 2 fc, + 1
                const uint4 v = __ldg(src + 4 * i);
                af[2 * i] = __hiloint2double(int(v.y), int(v.x));
                af[2 * i + 1] = __hiloint2double(int(v.w), int(v.z));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) af[j] = 0.0;
        }
    };

    // software pipeline over the warp's nodes: (node, e0, d, gather word) of the next node are fetched while this one
    // is being worked on, the node id one further ahead
    unsigned ni = blockIdx.x * NW + warp;
    unsigned node = 0, d = 0, g = 0, node_n = 0;
    unsigned long long e0 = 0;
    if (ni < a.nnodes) {
        node = __ldg(a.nodes + ni);
        e0 = __ldg(a.row_ptr + node);
        d = unsigned(__ldg(a.row_ptr + node + 1) - e0);
        g = (unsigned(lane) < d) ? __ldg(a.rev + e0 + lane) : 0u;
    }
    if (ni + nwt < a.nnodes) node_n = __ldg(a.nodes + ni + nwt);

    for (; ni < a.nnodes; ni += nwt) {
        const bool have_n = ni + nwt < a.nnodes;
        unsigned long long e0_n = 0, e1_n = 0;
        if (have_n) {
            e0_n = __ldg(a.row_ptr + node_n);
            e1_n = __ldg(a.row_ptr + node_n + 1);
        }
        const unsigned node_nn = (ni + 2u * nwt < a.nnodes) ? __ldg(a.nodes + ni + 2u * nwt) : 0u;

        // ---- contraction, 8 in-edges at a time -> b slab (rows = the node's slots)
        if constexpr (kF64) {
            double af[8], afn[8];
            load_afrag(0, d, g, af);
            for (unsigned k0 = 0; k0 < d; k0 += 8) {
                if (k0 + 8 < d) load_afrag(k0 + 8, d, g, afn);  // next block's loads fly during this block's DMMAs
                double cf[4][2];
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) cf[qb][0] = cf[qb][1] = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma_m8n8k4(cf[qb][0], cf[qb][1], af[j], sKf[(j * 4 + qb) * 32 + lane]);
                }
                const unsigned k = k0 + fr;
                if (k < d) {
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb)
                        *reinterpret_cast<double2 *>(reinterpret_cast<double *>(sb) + size_t(k) * kWideRow + 8 * qb + 2 * fc) =
                            make_double2(cf[qb][0], cf[qb][1]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) af[j] = afn[j];
            }
        } else {
            constexpr int VW = 16 / int(sizeof(T));  // elements per 16-byte broadcast load
            for (unsigned k0 = 0; k0 < d; k0 += 8) {
                T psi[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const unsigned k = k0 + r;
                    const unsigned gk = __shfl_sync(0xffffffffu, g, int(k & 31u));
                    psi[r] = (k < d) ? __ldg(Sold + size_t(gk) * QT + lane) : T(0);
                }
                __syncwarp();  // the previous block's broadcasts are done
#pragma unroll
                for (int r = 0; r < 8; ++r) stage[r * 32 + lane] = psi[r];
                __syncwarp();
                T acc[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = T(0);
#pragma unroll
                for (int tv = 0; tv < QT / VW; ++tv) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        T p[VW];
                        *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(stage + r * 32 + VW * tv);  // broadcast
#pragma unroll
                        for (int j = 0; j < VW; ++j) acc[r] += kcol[VW * tv + j] * p[j];
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (k0 + r < d) sb[size_t(k0 + r) * kWideRow + lane] = acc[r];
            }
        }
        // first batch of old out-messages, and the next node's gather words (its row offsets have arrived by now)
        T oldv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) oldv[u] = (unsigned(u) < d) ? __ldg(Sold + size_t(e0 + u) * QT + lane) : T(0);
        const unsigned d_n = unsigned(e1_n - e0_n);
        const unsigned g_n = (have_n && unsigned(lane) < d_n) ? __ldg(a.rev + e0_n + lane) : 0u;
        __syncwarp();

        // ---- node: product over the in-edges (slot order), marginal
        double tot = 1.0;
        for (unsigned k = 0; k < d; ++k) tot *= double(sb[size_t(k) * kWideRow + lane]);
        const double F = dc ? exp(-1.0 * double(d) * h_q / Nd) : exph_q;
        tot = tot * eta_q * F;
        const double sum = warp_sum(tot);
        const double mg = tot / sum;
        a.marg[size_t(node) * QT + lane] = mg;
        wsum += (dc ? double(d) : 1.0) * mg;

        // ---- edges: leave-one-out by division, normalise over q, max-diff, damped write (rows e0 .. e0 + d);
        // the old values of the next four slots are in flight while these four are worked on
        const T mgT = T(mg);
        for (unsigned k = 0; k < d; k += 4) {
            T oldn[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                oldn[u] = (k + 4 + u < d) ? __ldg(Sold + size_t(e0 + k + 4 + u) * QT + lane) : T(0);
                bv[u] = (k + u < d) ? sb[size_t(k + u) * kWideRow + lane] : T(1);
            }
            // four independent chains (no control flow between them, so the shuffles and reciprocals interleave)
            T cav[4];
            bool tiny = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                cav[u] = mgT * fast_rcp(bv[u]);
                tiny = tiny || !(double(bv[u]) >= kEps);
            }
            if (__any_sync(0xffffffffu, tiny)) {
                // a vanishing b_e[q]: leave-one-out product taken directly (see sweep_kernel.cuh); rare
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (k + u < d && __any_sync(0xffffffffu, !(double(bv[u]) >= kEps))) {
                        if (lane == 0) atomicAdd(&a.ctl->tiny_count, 1ull);
                        double p = 1.0;
                        for (unsigned kk = 0; kk < d; ++kk)
                            if (kk != k + u) p *= double(sb[size_t(kk) * kWideRow + lane]);
                        cav[u] = T(p * eta_q * F);
                    }
                }
            }
            T ssum[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ssum[u] = cav[u];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) ssum[u] += __shfl_xor_sync(0xffffffffu, ssum[u], o);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const T inv = fast_rcp(ssum[u]);
                const T nv = cav[u] * inv;
                if (k + u < d) {
                    if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;  // non-finite message: make it visible
                    mydiff = fmax(mydiff, fabs(double(oldv[u]) - double(nv)));
                    Snew[size_t(e0 + k + u) * QT + lane] = damp * nv + keep * oldv[u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) oldv[u] = oldn[u];
        }
        __syncwarp();  // the slab is free for the next node

        node = node_n;
        e0 = e0_n;
        d = d_n;
        g = g_n;
        node_n = node_nn;
    }

    // ---- one row per CTA: lane q holds column q; warps in a fixed order
    mydiff = warp_max(mydiff);
    s_rows[warp][lane] = wsum;
    if (lane == 0) s_rows[warp][QT] = mydiff;
    __syncthreads();
    if (tid <= QT) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v = (tid < QT) ? v + s_rows[w][tid] : fmax(v, s_rows[w][tid]);
        a.partial[size_t(a.rows_before + blockIdx.x) * (QT + 1) + tid] = v;
    }
    SweepArgsBase base;
    base.prm = a.prm;
    base.field[0] = a.field[0];
    base.field[1] = a.field[1];
    base.ctl = a.ctl;
    base.partial = a.partial;
    close_sweep_last_cta<QT>(base, gridDim.x + a.rows_before, sweeps_done, nullptr, gridDim.x);
}

}  // namespace sbmbp
