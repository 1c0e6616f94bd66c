// Degree-class (ELL) sweep kernel: the small-Q path for graphs whose message buffers are of the order of the 126 MB L2.
//
// Nodes are binned by (destination bucket, degree) -- degrees 0..31 -- and every class is cut into chunks of 32 nodes.
// ONE THREAD PER NODE, one warp per chunk; the index words of slot l of lane r sit at  chunk base + 32 l + r, so
//   * every lane of a warp runs the same trip count d -- no divergence, and for d <= DU the loops are unrolled with
//     the contracted messages b_l held in registers (no shared memory, no barrier, no edge -> node table);
//   * index words are read coalesced; the old out-messages (max-diff / damping) and the new ones of one (chunk, l)
//     are consecutive per destination bucket (engine.cu, build_bell_layout);
//   * the one random access is the gather of the in-message (rev[] = its position), confined to the region of the
//     bucket being processed: measured on B200, random 16-byte gathers run at 123-160 G/s inside 8-32 MiB windows
//     but only ~54 G/s inside 64 MiB (profiles/microbench_gather_r01.md);
//   * the streaming operands never wait on HBM: each warp issues prefetch.global.L2 for the index lines of the
//     chunk it will process two rounds later, and the whole grid streams the source buffer into the L2 a few MiB
//     ahead of the processing front, so the first gather into a line is an L2 hit too.
// Reference: sum_all_messages_to_i / norm_m_at_i (belief_propagation.cpp:991-1071), synchronous, product domain
// (all degrees here are < 50).  Same arithmetic as sweep_fast.cuh: product in slot order, leave-one-out as a
// product for Q <= 4; a node one of whose b_l[q] underflows 1e-50 takes the exact leave-one-out product.
// Nodes of degree >= 32 are left to bp_sweep_warp_kernel / bp_sweep_hub_kernel (launched just before); each kernel
// leaves one row of field partials per CTA and the last CTA of this kernel closes the sweep in a fixed order.
//
// Why not everywhere: with many destination buckets the 32 out-messages of a (chunk, l) scatter over as many
// regions and the writes stop coalescing; from 8 buckets on the engine uses the CTA-tile kernels (sweep_pipe.cuh).
#pragma once
#include "bp_device.cuh"
#include "sweep_fast.cuh"

namespace sbmbp {

#ifndef SBMBP_ELL_MINB
#define SBMBP_ELL_MINB 3
#endif

template <typename T>
struct EllSweepArgs {
    const EllClass *cls;
    unsigned ncls;
    unsigned nchunks;          // 32-node chunks over all classes
    const unsigned *ell_rev;   // per index word: buffer position of the in-message of that slot
    const unsigned *ell_pos;   // per index word: buffer position of its out-message
    const unsigned *ell_node;  // node ids, by class then ascending
    unsigned lines;            // 128-byte lines of one message buffer
    unsigned lpc;              // L2 stream-ahead: lines per chunk (0 = off) ...
    unsigned ahead;            // ... and how many chunks ahead of the processing front
    T *S[2];
    double *marg;
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    double *partial;           // [gridDim.x + rows_before][QT + 1]
    unsigned rows_before;      // rows left after this kernel's own by the warp / hub kernels of the same sweep
    unsigned dc;
    double damping;
};

template <typename T, int QT>
struct EllCtx {
    const T *Sold;
    T *Snew;
    const unsigned *ell_rev;
    const unsigned *ell_pos;
    const T *K;         // QT x QT kernel matrix (shared memory)
    const double *eta;  // shared
    T damp, keep;
};

template <int QT>
struct EllOut {  // what a node update adds to the lane's running row
    double w[QT];
    double maxdiff;
};

// out-message of one slot from its leave-one-out vector: normalise, max |old - new|, damped write
template <typename T, int QT>
__device__ __forceinline__ void ell_emit(const EllCtx<T, QT> &c, const T (&cav_in)[QT], const MsgVec<T, QT> &oldv, unsigned p,
                                         double &mydiff) {
    T s = T(0);
#pragma unroll
    for (int q = 0; q < QT; ++q) s += cav_in[q];
    const T inv = T(1) / s;
    if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;  // non-finite message: make it visible
    MsgVec<T, QT> out;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        const T nv = cav_in[q] * inv;
        mydiff = fmax(mydiff, fabs(double(oldv.v[q]) - double(nv)));
        out.v[q] = c.damp * nv + c.keep * oldv.v[q];
    }
    st_vec<T, QT>(out, c.Snew + size_t(p) * QT);
}

// node total (product of the b_l) -> normalised marginal, written out; tot becomes the marginal.
// F[q]: field factor of the class, exp(-d h_q / N) (dc) or exp(-beta h_q / N).
template <typename T, int QT>
__device__ __forceinline__ void ell_node_total(const EllCtx<T, QT> &c, const double *F, double wgt, double (&tot)[QT],
                                               double (&wsum)[QT], double *marg_out) {
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        tot[q] = tot[q] * c.eta[q] * F[q];
        sum += tot[q];
    }
    MsgVec<double, QT> mg;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        mg.v[q] = tot[q] / sum;
        tot[q] = mg.v[q];
        wsum[q] += wgt * mg.v[q];
    }
    st_vec<double, QT>(mg, marg_out);
}

// Run-time degree: two passes, the second gathers again instead of keeping d vectors per thread.  Used for the
// classes above the unrolled ones and as the fallback of a node one of whose b_l[q] underflows 1e-50 (there the
// leave-one-out product is taken directly, see sweep_kernel.cuh).  Out of line: rare.
// ib: index word of slot 0 of this lane; slot l: ib + 32 l.
template <typename T, int QT>
__device__ __noinline__ EllOut<QT> ell_update_loop(const EllCtx<T, QT> c, const double *F, double wgt, unsigned d, unsigned ib,
                                                   double *marg_out) {
    EllOut<QT> o;
    double tot[QT], wsum[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        tot[q] = 1.0;
        wsum[q] = 0.0;
    }
    double mydiff = 0.0;
    for (unsigned l = 0; l < d; ++l) {
        MsgVec<T, QT> m;
        ld_vec<T, QT>(m, c.Sold + size_t(__ldg(c.ell_rev + ib + 32u * l)) * QT);
        T b[QT];
        contract<T, QT>(m, c.K, b);
#pragma unroll
        for (int q = 0; q < QT; ++q) tot[q] *= double(b[q]);
    }
    ell_node_total<T, QT>(c, F, wgt, tot, wsum, marg_out);
    for (unsigned l = 0; l < d; ++l) {
        const unsigned p = __ldg(c.ell_pos + ib + 32u * l);
        MsgVec<T, QT> m, oldv;
        ld_vec<T, QT>(m, c.Sold + size_t(__ldg(c.ell_rev + ib + 32u * l)) * QT);
        ld_vec<T, QT>(oldv, c.Sold + size_t(p) * QT);
        T b[QT], cav[QT];
        contract<T, QT>(m, c.K, b);
        bool tiny = false;
#pragma unroll
        for (int q = 0; q < QT; ++q) tiny = tiny || !(double(b[q]) >= kEps);
        if (!tiny) {
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                T v = T(tot[q]);
#pragma unroll
                for (int r = 0; r < QT; ++r)
                    if (r != q) v *= b[r];
                cav[q] = v;
            }
        } else {
            double pr[QT];
#pragma unroll
            for (int q = 0; q < QT; ++q) pr[q] = 1.0;
            for (unsigned l2 = 0; l2 < d; ++l2) {
                if (l2 == l) continue;
                MsgVec<T, QT> m2;
                ld_vec<T, QT>(m2, c.Sold + size_t(__ldg(c.ell_rev + ib + 32u * l2)) * QT);
                T b2[QT];
                contract<T, QT>(m2, c.K, b2);
#pragma unroll
                for (int q = 0; q < QT; ++q) pr[q] *= double(b2[q]);
            }
#pragma unroll
            for (int q = 0; q < QT; ++q) cav[q] = T(pr[q] * c.eta[q] * F[q]);
        }
        ell_emit<T, QT>(c, cav, oldv, p, mydiff);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) o.w[q] = wsum[q];
    o.maxdiff = mydiff;
    return o;
}

// degree D known at compile time: b_l in registers, everything unrolled
template <typename T, int QT, int D>
__device__ __forceinline__ void ell_update_fixed(const EllCtx<T, QT> &c, const double *F, double wgt, unsigned ib,
                                                 double *marg_out, double (&wsum)[QT], double &mydiff) {
    constexpr int DD = D > 0 ? D : 1;
    // old out-messages in two batches (slots 0-3, then 4-7): the first rides with the gathers, the second is issued
    // after the node total so that at most 4 + D message vectors are live; the b_l stay in registers throughout
    constexpr int N0 = D < 4 ? D : 4, N1 = D - N0;
    T b[DD][QT];
    unsigned pw[DD];
    MsgVec<T, QT> old0[N0 > 0 ? N0 : 1], old1[N1 > 0 ? N1 : 1];
    double tot[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) tot[q] = 1.0;
    bool tiny = false;
    if constexpr (D > 0) {
        unsigned g[D];
#pragma unroll
        for (int l = 0; l < D; ++l) {
            g[l] = __ldg(c.ell_rev + ib + 32 * l);
            pw[l] = __ldg(c.ell_pos + ib + 32 * l);
        }
        MsgVec<T, QT> m[D];
#pragma unroll
        for (int l = 0; l < D; ++l) ld_vec<T, QT>(m[l], c.Sold + size_t(g[l]) * QT);
#pragma unroll
        for (int l = 0; l < N0; ++l) ld_vec<T, QT>(old0[l], c.Sold + size_t(pw[l]) * QT);
#pragma unroll
        for (int l = 0; l < D; ++l) {
            contract<T, QT>(m[l], c.K, b[l]);
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                tot[q] *= double(b[l][q]);
                tiny = tiny || !(double(b[l][q]) >= kEps);
            }
        }
    }
    if (tiny) {  // rare: the node goes through the general routine
        const EllOut<QT> o = ell_update_loop<T, QT>(c, F, wgt, unsigned(D), ib, marg_out);
#pragma unroll
        for (int q = 0; q < QT; ++q) wsum[q] += o.w[q];
        mydiff = fmax(mydiff, o.maxdiff);
        return;
    }
#pragma unroll
    for (int l = 0; l < N1; ++l) ld_vec<T, QT>(old1[l], c.Sold + size_t(pw[N0 + l]) * QT);
    ell_node_total<T, QT>(c, F, wgt, tot, wsum, marg_out);
#pragma unroll
    for (int l = 0; l < N0; ++l) {
        T cav[QT];
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            T v = T(tot[q]);
#pragma unroll
            for (int r = 0; r < QT; ++r)
                if (r != q) v *= b[l][r];
            cav[q] = v;
        }
        ell_emit<T, QT>(c, cav, old0[l], pw[l], mydiff);
    }
#pragma unroll
    for (int l = 0; l < N1; ++l) {
        T cav[QT];
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            T v = T(tot[q]);
#pragma unroll
            for (int r = 0; r < QT; ++r)
                if (r != q) v *= b[N0 + l][r];
            cav[q] = v;
        }
        ell_emit<T, QT>(c, cav, old1[l], pw[N0 + l], mydiff);
    }
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T, int QT>
__global__ void __launch_bounds__(kThreads, SBMBP_ELL_MINB) bp_sweep_ell_kernel(const EllSweepArgs<T> a) {
    static_assert(QT <= 4, "the degree-class kernel is the small-Q path");
    constexpr int NW = kThreads / 32;
    constexpr int DU = (QT * int(sizeof(T)) <= 16) ? 8 : 4;  // degrees unrolled with b_l in registers
    __shared__ EllClass s_cls[kEllMaxClasses];
    __shared__ __align__(16) T s_K[QT * QT];
    __shared__ double s_eta[QT];
    __shared__ double s_F[kEllDegrees][QT];  // field factor per degree: exp(-d h_q / N) (dc) or exp(-beta h_q / N)
    __shared__ double s_rows[NW][QT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctl *ctl = a.ctl;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;  // uniform over the grid
    const int par = int(sweeps_done & 1u);
    const Field *fld = par ? a.field[1] : a.field[0];
    for (unsigned i = tid; i < a.ncls; i += kThreads) s_cls[i] = a.cls[i];
    for (int i = tid; i < QT * QT; i += kThreads) s_K[i] = T(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) s_eta[tid] = a.prm->eta[tid];
    if (unsigned(tid) < kEllDegrees * QT) {
        const unsigned d = tid / QT, q = tid % QT;
        s_F[d][q] = (a.dc != 0) ? exp(-1.0 * double(d) * fld->h[q] / a.prm->N) : fld->exph[q];
    }
    __syncthreads();

    EllCtx<T, QT> c;
    c.Sold = par ? a.S[1] : a.S[0];
    c.Snew = par ? a.S[0] : a.S[1];
    c.ell_rev = a.ell_rev;
    c.ell_pos = a.ell_pos;
    c.K = s_K;
    c.eta = s_eta;
    c.damp = T(a.damping);
    c.keep = T(1.0 - a.damping);
    const bool dc = a.dc != 0;
    const char *sold_bytes = reinterpret_cast<const char *>(c.Sold);

    double wsum[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;

    const unsigned nwt = gridDim.x * NW;
    const unsigned ncls = a.ncls, nchunks = a.nchunks;
    unsigned ci = 0, ci2 = 0;  // class of the current chunk / of the chunk two rounds ahead (both only move forward)
    for (unsigned chunk = blockIdx.x * NW + warp; chunk < nchunks; chunk += nwt) {
        // ---- keep the L2 ahead of the loads: index lines of the chunk this warp takes two rounds from now ...
        const unsigned chunk2 = chunk + 2u * nwt;
        if (chunk2 < nchunks) {
            while (ci2 + 1 < ncls && s_cls[ci2 + 1].chunk_first <= chunk2) ++ci2;
            const unsigned d2 = s_cls[ci2].d, cc2 = chunk2 - s_cls[ci2].chunk_first;
            const unsigned ib2 = s_cls[ci2].base + cc2 * 32u * d2;
            for (unsigned j = lane; j < 2u * d2 + 1u; j += 32u) {
                const void *ptr = (j < d2)        ? static_cast<const void *>(a.ell_rev + ib2 + 32u * j)
                                  : (j < 2u * d2) ? static_cast<const void *>(a.ell_pos + ib2 + 32u * (j - d2))
                                                  : static_cast<const void *>(a.ell_node + s_cls[ci2].node_first + cc2 * 32u);
                prefetch_l2(ptr);
            }
        }
        // ... and this chunk's share of the source buffer, `ahead` chunks in front of the processing front
        if (a.lpc) {
            for (unsigned j = lane; j < a.lpc; j += 32u) {
                const unsigned long long line = (unsigned long long)(chunk + a.ahead) * a.lpc + j;
                if (line < a.lines) prefetch_l2(sold_bytes + line * 128ull);
                if (chunk < a.ahead) {
                    const unsigned long long line0 = (unsigned long long)chunk * a.lpc + j;
                    if (line0 < a.lines) prefetch_l2(sold_bytes + line0 * 128ull);
                }
            }
        }

        while (ci + 1 < ncls && s_cls[ci + 1].chunk_first <= chunk) ++ci;  // warp-uniform
        const EllClass cl = s_cls[ci];
        const unsigned cc = chunk - cl.chunk_first;
        const unsigned r = cc * 32u + unsigned(lane);
        if (r < cl.n) {
            const unsigned ib = cl.base + cc * 32u * cl.d + unsigned(lane);
            const unsigned node = __ldg(a.ell_node + cl.node_first + r);
            double *mo = a.marg + size_t(node) * QT;
            const double *F = s_F[cl.d];
            const double wgt = dc ? double(cl.d) : 1.0;
            bool done = true;
            switch (cl.d) {
                case 0: ell_update_fixed<T, QT, 0>(c, F, wgt, ib, mo, wsum, mydiff); break;
                case 1: ell_update_fixed<T, QT, 1>(c, F, wgt, ib, mo, wsum, mydiff); break;
                case 2: ell_update_fixed<T, QT, 2>(c, F, wgt, ib, mo, wsum, mydiff); break;
                case 3: ell_update_fixed<T, QT, 3>(c, F, wgt, ib, mo, wsum, mydiff); break;
                case 4: ell_update_fixed<T, QT, 4>(c, F, wgt, ib, mo, wsum, mydiff); break;
                default: done = false; break;
            }
            if constexpr (DU >= 8) {
                if (!done) {
                    done = true;
                    switch (cl.d) {
                        case 5: ell_update_fixed<T, QT, 5>(c, F, wgt, ib, mo, wsum, mydiff); break;
                        case 6: ell_update_fixed<T, QT, 6>(c, F, wgt, ib, mo, wsum, mydiff); break;
                        case 7: ell_update_fixed<T, QT, 7>(c, F, wgt, ib, mo, wsum, mydiff); break;
                        case 8: ell_update_fixed<T, QT, 8>(c, F, wgt, ib, mo, wsum, mydiff); break;
                        default: done = false; break;
                    }
                }
            }
            if (!done) {
                const EllOut<QT> o = ell_update_loop<T, QT>(c, F, wgt, cl.d, ib, mo);
#pragma unroll
                for (int q = 0; q < QT; ++q) wsum[q] += o.w[q];
                mydiff = fmax(mydiff, o.maxdiff);
            }
        }
    }

    // ---- one row per CTA: warps in a fixed order
    mydiff = warp_max(mydiff);
#pragma unroll
    for (int q = 0; q < QT; ++q) wsum[q] = warp_sum(wsum[q]);
    if (lane == 0) {
        s_rows[warp][QT] = mydiff;
#pragma unroll
        for (int q = 0; q < QT; ++q) s_rows[warp][q] = wsum[q];
    }
    __syncthreads();
    if (tid <= QT) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v = (tid < QT) ? v + s_rows[w][tid] : fmax(v, s_rows[w][tid]);
        a.partial[size_t(blockIdx.x) * (QT + 1) + tid] = v;
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
