// The BP sweep for sm_100a: one synchronous pass over all M directed edges.
//
// Replaces the body of converge() (belief_propagation.cpp:392-405) and the per-node routines it calls:
// sum_all_messages_to_i (:991-1049), norm_m_at_i (:1051-1071), bp_iter_update_psi_large_degree (:813-890),
// update_h / update_exph_with_h (:334-368) -- evaluated for every node at once from the previous sweep's
// messages (Jacobi), instead of N random in-place draws.
//
// One CTA owns one node-aligned tile (<= TE edges, <= TN nodes):
//   phase 0  row offsets of the tile's nodes -> smem; edge -> local node map
//   phase 1  per in-edge e (one thread each, coalesced over e): gather the message into i from S_old[rev[e]]
//            (the only random access of the sweep, one 16/32-byte vector load), contract it with the
//            Q x Q kernel, b_e[q] = sum_t K[t][q] psi[t], keep b_e in smem
//   phase 2  per node: combine the b_e -- as a product for degree < 50 and as a sum of logs for degree >= 50,
//            the reference's own split -- multiply by eta_q and the field term, normalise -> marginal (written
//            coalesced), and accumulate w_i * psi_i for the next sweep's h.  Degree < 32: one thread per node;
//            degree >= 32: one warp per node with shuffle reductions
//   phase 3  per out-edge e (one thread each, coalesced): cavity message = node total with b_e divided out
//            (leave-one-out), normalise, max |old - new| against S_old[e], damped write to S_new[e]
// A node with more than TE edges (a hub) gets a CTA of its own and a two-pass log-domain update.
// The last CTA to finish reduces the per-tile field partials in a fixed order (bitwise reproducible),
// publishes h / exp(-beta h/N) for the next sweep, the sweep's max-diff, and the convergence flag.
#pragma once
#include <type_traits>

#include "bp_device.cuh"
#include "dist_exchange.cuh"

namespace sbmbp {

// the type-independent part of the sweep arguments
struct SweepArgsBase {
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    double *partial;
};

template <typename T>
struct SweepArgs {
    const Tile *tiles;
    const unsigned long long *row_ptr;
    const unsigned *rev;     // rev[e]: buffer position of the message INTO row(e) along e
    const unsigned *pos;     // buffer positions of the messages OUT of a tile's nodes, sorted ascending within each
                             // tile; entry e0 + t belongs to the tile-local slot info[e0 + t] & 0xffff of the
                             // tile-local node (info >> 16) & 0x7fff; bit 31 of info: that node has degree >= 50.
                             // (Hub tiles: slot order, info unused.)
    const unsigned *info;
    const unsigned *degsrc;  // degree of col[e]; only read when dc == 2
    T *S[2];
    double *marg;
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    // multi-GPU (DIST kernels only; dist_exchange.cuh): the outbox of this rank's REMOTE out-messages (their old values
    // for phase 3, and what gets shipped), the message buffers of every rank (own + CUDA-IPC mapped peers), and the
    // exchange tables.  A pos word with bit 31 set is an outbox index, otherwise a position in this rank's own buffer.
    T *mirror;
    T *peer[2][8];
    DistArgs dx;
    double *partial;  // [ntiles][QT + 1]: per-tile sum of w_i psi_i^t (t < QT), then the tile's max |old - new|
    unsigned ntiles;
    unsigned Q;
    unsigned dc;
    double *row_out;  // multi-GPU: where the last CTA leaves this rank's reduced row (else nullptr)
    int fused_close;  // persistent kernels: the last CTA closes the sweep (no finalize launch)
    int gmode;     // load flavour of the message gather (see ld_gather16)
    int select_k;  // dc == 0 and beta != 1: the two degree classes use different kernels
    double damping;
    const int *clamp;  // general kernel only: conf_planted_ per node (-1 = free) when bp_conditional applies, else nullptr
    // general kernel only, coloured asynchronous schedule: one pass updates the nodes of colour cur_color and carries
    // every other node forward unchanged (nullptr: synchronous sweep)
    const unsigned char *color;
    unsigned cur_color;
};

// h_q = sum_t c_tq wsum_t ; exph_q = exp(-beta h_q / N)
__device__ inline void publish_field(const DevParams *prm, unsigned Q, const double *wsum, Field *out) {
    for (unsigned q = 0; q < Q; ++q) {
        double h = 0.0;
        for (unsigned t = 0; t < Q; ++t) h += prm->C[t * kMaxQ + q] * wsum[t];
        out->h[q] = h;
        out->exph[q] = exp(-prm->beta * h / prm->N);
        out->wsum[q] = wsum[q];
    }
}

// Closes a sweep (one CTA, launched right after the sweep kernel): fixed-order sum of the per-tile partial rows
// -> h / exp(-beta h / N) for the next sweep; max of the per-tile max-diffs; sweep counter, convergence flag.
// Keeping this out of the sweep kernel means a sweep CTA ends with one plain store per column -- no fence, no
// atomic, no "am I last" round trip while its registers and shared memory sit idle.
constexpr int kFinalThreads = 256;
template <int QT>
__global__ void __launch_bounds__(kFinalThreads) bp_finalize_kernel(const double *__restrict__ partial,
                                                                    unsigned ntiles, unsigned Q,
                                                                    const DevParams *prm, Field *f0, Field *f1,
                                                                    Ctl *ctl) {
    constexpr int NC = QT + 1;  // row = QT field partials, then the tile's max-diff
    __shared__ double sred[kFinalThreads / 32][NC];
    __shared__ double tot[NC];
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;  // the sweep before this was a no-op too
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[NC];
SBMBP_UNROLL_Q
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (unsigned b = tid; b < ntiles; b += kFinalThreads) {
        const double *row = partial + size_t(b) * NC;
SBMBP_UNROLL_Q
        for (int c = 0; c < NC; ++c) acc[c] = (c < QT) ? acc[c] + row[c] : fmax(acc[c], row[c]);
    }
SBMBP_UNROLL_Q
    for (int c = 0; c < NC; ++c) {
        const double v = (c < QT) ? warp_sum(acc[c]) : warp_max(acc[c]);
        if (lane == 0) sred[warp][c] = v;
    }
    __syncthreads();
    if (tid < NC) {  // fixed order over the warps
        double r = sred[0][tid];
        for (int w = 1; w < kFinalThreads / 32; ++w) r = (tid < QT) ? r + sred[w][tid] : fmax(r, sred[w][tid]);
        tot[tid] = r;
    }
    __syncthreads();
    if (tid == 0) {
        publish_field(prm, Q, tot, (sweeps_done & 1u) ? f0 : f1);
        const double md = tot[QT];
        ctl->last_maxdiff = md;
        ctl->sweeps_done = sweeps_done + 1;
        if (!(md == md) || md > 1.0e299) ctl->nan_count += 1;
        if (md < ctl->crit) {  // double < float, as belief_propagation.cpp:406
            ctl->converged = 1;
            ctl->niter = int(sweeps_done - ctl->sweep_base);
        }
    }
}

// Totals of a sweep's per-CTA rows, by one whole CTA: all threads read rows (every column of a row at once), then a
// fixed-shape tree -- deterministic, ~1 us.  s_tot: shared, [QT + 1]; valid for every thread on return.
template <int QT, int NT = kThreads>
__device__ __forceinline__ void reduce_rows_cta(const double *partial, unsigned nrows, double *s_tot) {
    constexpr int NC = QT + 1;
    __shared__ double s_part[NT / 32][NC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[NC];
SBMBP_UNROLL_Q
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (unsigned k = tid; k < nrows; k += NT) {
        double v[NC];
SBMBP_UNROLL_Q
        for (int c = 0; c < NC; ++c) v[c] = __ldcg(partial + size_t(k) * NC + c);
SBMBP_UNROLL_Q
        for (int c = 0; c < NC; ++c) acc[c] = (c < QT) ? acc[c] + v[c] : fmax(acc[c], v[c]);
    }
SBMBP_UNROLL_Q
    for (int c = 0; c < NC; ++c) {
        const double v = (c < QT) ? warp_sum(acc[c]) : warp_max(acc[c]);
        if (lane == 0) s_part[warp][c] = v;
    }
    __syncthreads();
    if (tid < NC) {
        double r = s_part[0][tid];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) r = (tid < QT) ? r + s_part[w][tid] : fmax(r, s_part[w][tid]);
        s_tot[tid] = r;
    }
    __syncthreads();
}

// h_q = sum_t c_tq wsum_t for one q (ccol[t] = c_tq), t ascending with fused multiply-adds: one definition, so that the
// sweep close and the lazy prologue of the next sweep (sweep_ell.cuh) publish the same bits
template <int QT>
__device__ __forceinline__ double field_component(const double *ccol, const double *tot) {
    double h = 0.0;
SBMBP_UNROLL_Q
    for (int t = 0; t < QT; ++t) h = fma(ccol[t], tot[t], h);
    return h;
}

// The same closing step, run by the LAST CTA of a persistent sweep kernel (one fence + one atomic per CTA, a few
// hundred per sweep): saves the finalize launch and the gap around it.  Rows are combined in CTA order.
// mode 0: publish the field and advance the control block; mode 1 (multi-GPU): only leave the rank's row in row_out.
template <int QT, int NT = kThreads>
__device__ __forceinline__ void close_sweep_last_cta(const SweepArgsBase &b, unsigned nrows, unsigned sweeps_done,
                                                     double *row_out, unsigned ndone = 0u) {
    if (ndone == 0u) ndone = nrows;  // CTAs that report in; rows beyond them were left by an earlier launch
    constexpr int NC = QT + 1;
    __shared__ int s_last;
    __shared__ double s_tot[NC];
    const int tid = threadIdx.x;
    // What the closing CTA needs besides the rows does not depend on them: fetch it now, so that the round trips overlap
    // the fence and the ticket below instead of following the reduction (small Q only: the column lives in registers).
    constexpr bool kPre = QT <= 8;
    double cpre[kPre ? QT : 1], beta_pre = 0.0, n_pre = 1.0;
    float crit_pre = 0.f;
    unsigned base_pre = 0u;
    if constexpr (kPre) {
        if (!row_out && tid < QT) {
SBMBP_UNROLL_Q
            for (int t = 0; t < QT; ++t) cpre[t] = b.prm->C[t * kMaxQ + tid];
            beta_pre = b.prm->beta;
            n_pre = b.prm->N;
        }
        if (!row_out && tid == 0) {
            crit_pre = b.ctl->crit;
            base_pre = b.ctl->sweep_base;
        }
    }
    // only the threads that just wrote the CTA's row (tid <= QT in every caller) need their stores ordered before the
    // ticket: a fence by all 256 threads makes every warp drain its message stores first (measured: ~3 us per CTA)
    if (tid < NC) __threadfence();
    __syncthreads();
    if (tid == 0) {
        s_last = (atomicAdd(&b.ctl->done, 1u) == ndone - 1);
        __threadfence();  // the tickets seen -> the rows read below (the barrier carries it to the other threads)
    }
    __syncthreads();
    if (!s_last) return;
    reduce_rows_cta<QT, NT>(b.partial, nrows, s_tot);
    if (row_out && tid < NC) row_out[tid] = s_tot[tid];
    if (!row_out && tid < QT) {  // one thread per component: the parameter loads and the exp() run side by side
        Field *out = (sweeps_done & 1u) ? b.field[0] : b.field[1];
        double h = 0.0;
        if constexpr (kPre) {
            h = field_component<QT>(cpre, s_tot);
            out->h[tid] = h;
            out->exph[tid] = exp(-beta_pre * h / n_pre);
        } else {
SBMBP_UNROLL_Q
            for (int t = 0; t < QT; ++t) h += b.prm->C[t * kMaxQ + tid] * s_tot[t];
            out->h[tid] = h;
            out->exph[tid] = exp(-b.prm->beta * h / b.prm->N);
        }
        out->wsum[tid] = s_tot[tid];
    }
    if (tid == 0) {
        b.ctl->done = 0;
        if (!row_out) {
            const double md = s_tot[QT];
            b.ctl->last_maxdiff = md;
            b.ctl->sweeps_done = sweeps_done + 1;
            if (!(md == md) || md > 1.0e299) b.ctl->nan_count += 1;
            const float crit = kPre ? crit_pre : b.ctl->crit;
            if (md < crit) {  // double < float, as belief_propagation.cpp:406
                b.ctl->converged = 1;
                b.ctl->niter = int(sweeps_done - (kPre ? base_pre : b.ctl->sweep_base));
            }
        }
    }
}

// Multi-GPU flavour of the closing step: the last CTA leaves the rank's row in EVERY rank's sync block and raises the
// rank's flag (dist_publish_row); the field and the convergence decision are taken from all ranks' rows by the next
// kernel on the stream (the next sweep's prologue, or bp_dist_close_kernel at the end of a batch).
template <typename T>
struct PeerPtrs {
    T *p[kMaxRanks];
};

// nrows: rows to reduce; ndone: CTAs that report in (0: one row per CTA)
template <int QT, int NT = kThreads>
__device__ __forceinline__ void close_sweep_dist(const SweepArgsBase &b, const DistArgs &d, unsigned nrows, unsigned seq,
                                                 unsigned ndone = 0u) {
    if (ndone == 0u) ndone = nrows;
    constexpr int NC = QT + 1;
    __shared__ int s_last;
    __shared__ double s_tot[NC];
    const int tid = threadIdx.x;
    if (tid < NC) __threadfence();
    __syncthreads();
    if (tid == 0) {
        s_last = (atomicAdd(&b.ctl->done, 1u) == ndone - 1);
        __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    reduce_rows_cta<QT, NT>(b.partial, nrows, s_tot);
    if (tid == 0) {
        b.ctl->done = 0;
        b.ctl->sweeps_done = seq + 1;
    }
    dist_publish_row<QT>(d, s_tot, seq);
}

// What a DIST sweep's prologue (from_rows) and the batch-closing kernel share: sweep seq - 1 of all ranks -> h, exp(-beta h
// / N), max-diff, convergence.  Whole CTA.  Returns 1 if that sweep converged (uniform over the grid and over the ranks).
// s_h / s_eh: shared, [QT].  `writer`: this CTA also records the outcome in the control block and the Field.
template <int QT>
__device__ __forceinline__ int dist_open_sweep(const SweepArgsBase &b, const DistArgs &d, unsigned seq, double *s_tot, double *s_h,
                                               double *s_eh, bool writer) {
    const int tid = threadIdx.x;
    double ccol[QT], beta = 0.0, n_nodes = 1.0;
    if (tid < QT) {
SBMBP_UNROLL_Q
        for (int t = 0; t < QT; ++t) ccol[t] = b.prm->C[t * kMaxQ + tid];
        beta = b.prm->beta;
        n_nodes = b.prm->N;
    }
    const float crit = b.ctl->crit;
    const unsigned sweep_base = b.ctl->sweep_base;
    dist_wait_and_reduce<QT>(d, seq, s_tot);
    const double md = s_tot[QT];
    if (tid < QT) {
        const double h = field_component<QT>(ccol, s_tot);
        s_h[tid] = h;
        s_eh[tid] = exp(-beta * h / n_nodes);
    }
    __syncthreads();
    const int conv = (md < crit) ? 1 : 0;  // double < float, as belief_propagation.cpp:406
    if (writer) {
        Field *out = (seq & 1u) ? b.field[1] : b.field[0];  // what a sweep starting from `seq` completed sweeps reads
        if (tid < QT) {
            out->h[tid] = s_h[tid];
            out->exph[tid] = s_eh[tid];
            out->wsum[tid] = s_tot[tid];
        }
        if (tid == 0) {
            b.ctl->last_maxdiff = md;
            b.ctl->sweeps_done = seq;
            if (!(md == md) || md > 1.0e299) b.ctl->nan_count += 1;
            if (conv) {
                b.ctl->converged = 1;
                b.ctl->niter = int(seq - 1u - sweep_base);
            }
        }
    }
    return conv;
}

// Multi-GPU sweep through the GENERAL kernel (padded Q, deg_corr_flag 2, beta != 1 -- bp_sweep_kernel has no persistent
// grid to ship from): this kernel follows it on the stream, carries the whole outbox to the owners (coalesced stores, as
// dist_ship_range, element by element because Q is a run-time value here) and its last CTA reduces the per-tile rows,
// publishes the rank's row and raises the flag.  No overlap of transfer and computation on this path.
template <typename T, int QT>
__global__ void __launch_bounds__(kThreads) dist_ship_publish_kernel(SweepArgsBase b, DistArgs d, const T *__restrict__ outbox,
                                                                      PeerPtrs<T> peers, unsigned n_remote, unsigned Q, unsigned ntiles) {
    const unsigned seq = b.ctl->sweeps_done;
    if (b.ctl->converged || seq >= b.ctl->max_sweeps) return;  // the sweep kernel before this one was a no-op too
    const unsigned long long total = (unsigned long long)n_remote * Q;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)kThreads + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * kThreads) {
        const unsigned k = unsigned(idx / Q), q = unsigned(idx - (unsigned long long)k * Q);
        const unsigned rp = __ldg(d.out_rpos + k);
        peers.p[rp >> 29][size_t(rp & ((1u << 29) - 1u)) * Q + q] = __ldcg(outbox + idx);
    }
    dist_ship_drain();
    close_sweep_dist<QT>(b, d, ntiles, seq, gridDim.x);
}

// closes the last sweep of a batch of DIST sweeps (one CTA): afterwards the field, the control block and the
// convergence decision are what a host-synchronised sweep would have left
template <int QT>
__global__ void __launch_bounds__(kFinalThreads) bp_dist_close_kernel(SweepArgsBase b, DistArgs d) {
    __shared__ double s_tot[QT + 1], s_h[QT], s_eh[QT];
    if (b.ctl->converged) return;  // a sweep of the batch already closed its predecessor and found it converged
    const unsigned seq = b.ctl->sweeps_done;
    if (seq == b.ctl->sweep_base) return;  // nothing was swept since the control block was armed
    dist_open_sweep<QT>(b, d, seq, s_tot, s_h, s_eh, true);
}

// Where an out-message lives (general kernel; the pipeline / fast kernels inline the same decode): multi-GPU engines keep
// remote out-messages in the outbox (a.mirror, pos word with bit 31 set), everything else in the rank's own buffer.
template <typename T>
__device__ __forceinline__ const T *out_msg_src(const SweepArgs<T> &a, const T *Sold, size_t w, unsigned Q) {
    return (a.mirror && (w & kRemoteBit)) ? a.mirror + (w & ~size_t(kRemoteBit)) * Q : Sold + w * Q;
}
template <typename T>
__device__ __forceinline__ T *out_msg_dst(const SweepArgs<T> &a, T *Snew, size_t w, unsigned Q) {
    return (a.mirror && (w & kRemoteBit)) ? a.mirror + (w & ~size_t(kRemoteBit)) * Q : Snew + w * Q;
}

// b[q] = sum_t K(t,q) m[t] for one in-edge.  FP32 storage with a long contraction (Q >= 8): the sum runs in double and is
// rounded once -- a float dot product of 32 terms carries ~3e-7, which a degree-400 log-domain node adds up to more
// than the 1e-5 bar of FP32 mode (measured 1.6e-5; 2e-6 with the double accumulator).
template <typename T, int QT>
__device__ __forceinline__ void contract(const MsgVec<T, QT> &m, const KernT<T, QT> *__restrict__ K, T (&b)[QT]) {
    using Acc = KernT<T, QT>;
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) {
        Acc acc = Acc(0);
SBMBP_UNROLL_Q
        for (int t = 0; t < QT; ++t) acc += Acc(K[t * QT + q]) * Acc(m.v[t]);
        b[q] = T(acc);
    }
}

// dc == 2: K(t,q) = tau / (1 + tau), tau = d_i d_l p_tq  (belief_propagation.cpp:1008-1010)
template <typename T, int QT>
__device__ __forceinline__ void contract_dc2(const MsgVec<T, QT> &m, const double *__restrict__ P, double didl,
                                             unsigned Q, T (&b)[QT]) {
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) {
        double acc = 0.0;
SBMBP_UNROLL_Q
        for (int t = 0; t < QT; ++t) {
            if (unsigned(t) < Q && unsigned(q) < Q) {
                double tau = didl * P[t * QT + q];
                acc += tau / (1.0 + tau) * double(m.v[t]);
            }
        }
        b[q] = T(acc);
    }
}

template <typename T, int QT>
__global__ void __launch_bounds__(kThreads) bp_sweep_kernel(const SweepArgs<T> a) {
    using Cfg = TileCfg<T, QT>;
    using Lay = TileSmem<T, QT>;
    constexpr int TE = Cfg::TE, TN = Cfg::TN;
    extern __shared__ __align__(16) unsigned char smem[];
    double *snum = reinterpret_cast<double *>(smem + Lay::off_num);
    double *sred = reinterpret_cast<double *>(smem + Lay::off_red);
    double *seta = reinterpret_cast<double *>(smem + Lay::off_par);
    double *slogeta = seta + QT;
    double *sh = seta + 2 * QT;
    double *sexph = seta + 3 * QT;
    KernT<T, QT> *sKs = reinterpret_cast<KernT<T, QT> *>(smem + Lay::off_ks);
    KernT<T, QT> *sKl = reinterpret_cast<KernT<T, QT> *>(smem + Lay::off_kl);
    double *sP = reinterpret_cast<double *>(smem + Lay::off_p);
    T *sb = reinterpret_cast<T *>(smem + Lay::off_b);
    unsigned *soff = reinterpret_cast<unsigned *>(smem + Lay::off_off);
    unsigned short *snode = reinterpret_cast<unsigned short *>(smem + Lay::off_node);

    Ctl *ctl = a.ctl;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;  // uniform over the grid
    const int par = int(sweeps_done & 1u);
    // (selected with ?: rather than indexed: a dynamically indexed kernel-parameter array is copied to local memory)
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const Field *fld = par ? a.field[1] : a.field[0];
    const unsigned Q = a.Q, dc = a.dc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double Nd = a.prm->N;

    // ---- parameters -> smem (padded to QT with zeros so padded components drop out)
    for (int i = tid; i < QT * QT; i += kThreads) {
        const int t = i / QT, q = i % QT;
        const bool in = unsigned(t) < Q && unsigned(q) < Q;
        sKs[i] = in ? KernT<T, QT>(a.prm->Ks[t * kMaxQ + q]) : KernT<T, QT>(0);
        sKl[i] = in ? KernT<T, QT>(a.prm->Kl[t * kMaxQ + q]) : KernT<T, QT>(0);
        sP[i] = in ? a.prm->P[t * kMaxQ + q] : 0.0;
    }
    if (tid < QT) {
        const bool in = unsigned(tid) < Q;
        seta[tid] = in ? a.prm->eta[tid] : 0.0;
        slogeta[tid] = in ? a.prm->logeta[tid] : 0.0;
        sh[tid] = in ? fld->h[tid] : 0.0;
        sexph[tid] = in ? fld->exph[tid] : 0.0;
    }

    const Tile tile = a.tiles[blockIdx.x];
    const unsigned long long e0 = tile.e0;
    const unsigned n0 = tile.n0, nn = tile.nn;
    const unsigned long long ne64 = tile.ne;

    double wsum[QT];  // this thread's share of sum_i w_i psi_i^t
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;
    unsigned mynan = 0;

    if (ne64 <= (unsigned long long)TE) {
        // =================================================================== regular tile
        const unsigned ne = unsigned(ne64);
        // Every thread owns EPT edges of the tile twice over: slots k_u = u*kThreads + tid (phase 1: gather and
        // contract) and buffer entries t_u = u*kThreads + tid (phase 3: old value, new value).  All their index
        // loads are issued here, before anything waits, and the old values follow as soon as the positions land:
        // by the time phase 1 computes, 2*EPT independent message loads per thread are in flight.
        constexpr int EPT = TE / kThreads;
        unsigned rv[EPT], ps[EPT], pm[EPT];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const unsigned k = u * kThreads + tid;
            const bool live = k < ne;
            rv[u] = live ? __ldg(a.rev + e0 + k) : 0u;
            ps[u] = live ? __ldg(a.pos + e0 + k) : 0u;
            pm[u] = live ? (__ldg(a.info + e0 + k) & 0xffffu) : k;
        }
        for (unsigned n = tid; n <= nn; n += kThreads) soff[n] = unsigned(a.row_ptr[n0 + n] - e0);
        __syncthreads();
        for (unsigned n = tid; n < nn; n += kThreads)
            for (unsigned k = soff[n]; k < soff[n + 1]; ++k) snode[k] = (unsigned short)n;
        __syncthreads();

        // ---- phase 1: gather + contract (slot order); the old values of phase 3 are fetched alongside
        MsgVec<T, QT> oldv[EPT];
        {
            MsgVec<T, QT> m[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (u * kThreads + tid < ne) m[u].gather(Sold + size_t(rv[u]) * Q, Q, a.gmode);
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (u * kThreads + tid < ne) oldv[u].load(out_msg_src<T>(a, Sold, ps[u], Q), Q);
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < ne) {
                    T b[QT];
                    if (dc == 2) {
                        const unsigned n = snode[k];
                        const double di = double(soff[n + 1] - soff[n]);
                        const double dl = double(__ldg(a.degsrc + e0 + k));
                        contract_dc2<T, QT>(m[u], sP, di * dl, Q, b);
                    } else {
                        const KernT<T, QT> *K = sKs;
                        if (a.select_k) {
                            const unsigned n = snode[k];
                            if (soff[n + 1] - soff[n] >= kLargeDegree) K = sKl;
                        }
                        contract<T, QT>(m[u], K, b);
                    }
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) sb[q * TE + k] = b[q];
                }
            }
        }
        __syncthreads();

        // ---- phase 2a: one thread per node of degree < 32 (product domain)
        for (unsigned n = tid; n < nn; n += kThreads) {
            const unsigned k0 = soff[n], d = soff[n + 1] - k0;
            if (d >= 32) continue;
            if ((a.clamp && a.clamp[n0 + n] != -1) || (a.color && a.color[n0 + n] != a.cur_color)) {
                // bp_conditional (belief_propagation.cpp:1100-1126): a planted node is not updated -- its marginal
                // stays (and keeps feeding h), its messages are copied forward in phase 3.  Same for the nodes that are
                // not of the current colour in the coloured schedule.
                const double w = (dc == 0) ? 1.0 : double(d);
                for (unsigned q = 0; q < Q; ++q) wsum[q] += w * a.marg[size_t(n0 + n) * Q + q];
                continue;
            }
            double tot[QT];
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) tot[q] = 1.0;
            for (unsigned k = k0; k < k0 + d; ++k) {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) tot[q] *= double(sb[q * TE + k]);
            }
            double sum = 0.0;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                if (unsigned(q) < Q) {
                    const double F = (dc == 0) ? sexph[q] : exp(-1.0 * double(d) * sh[q] / Nd);
                    tot[q] = tot[q] * seta[q] * F;
                    sum += tot[q];
                } else {
                    tot[q] = 0.0;
                }
            }
            const double w = (dc == 0) ? 1.0 : double(d);
            MsgVec<double, QT> mg;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                mg.v[q] = tot[q] / sum;
                snum[q * TN + n] = mg.v[q];
                wsum[q] += w * mg.v[q];
            }
            mg.store(a.marg + size_t(n0 + n) * Q, Q);
        }
        // ---- phase 2b: one warp per node of degree >= 32 (product below 50, log domain from 50 on)
        for (unsigned n = warp; n < (tile.nbig ? nn : 0u); n += kThreads / 32) {
            const unsigned k0 = soff[n], d = soff[n + 1] - k0;
            if (d < 32) continue;
            const bool logdom = d >= kLargeDegree;
            if ((!logdom && a.clamp && a.clamp[n0 + n] != -1) || (a.color && a.color[n0 + n] != a.cur_color)) {
                if (lane == 0) {  // planted, product domain: frozen (the reference's log-domain routine ignores conf_planted_)
                    for (unsigned q = 0; q < Q; ++q) wsum[q] += ((dc == 0) ? 1.0 : double(d)) * a.marg[size_t(n0 + n) * Q + q];
                }
                continue;
            }
            double acc[QT];
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) acc[q] = logdom ? 0.0 : 1.0;
            for (unsigned k = k0 + lane; k < k0 + d; k += 32) {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const double bv = double(sb[q * TE + k]);
                    if (logdom) acc[q] += (unsigned(q) < Q) ? log(bv) : 0.0;
                    else acc[q] *= bv;
                }
            }
            double mx = -1.0e300, sum = 0.0;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                if (logdom) {
                    acc[q] = warp_sum(acc[q]);
                    if (unsigned(q) < Q) {  // :850-853, no beta on this path
                        acc[q] = acc[q] + slogeta[q] - ((dc == 0) ? sh[q] / Nd : 1.0 * double(d) * sh[q] / Nd);
                        mx = fmax(mx, acc[q]);
                    }
                } else {
                    acc[q] = warp_prod(acc[q]);
                    if (unsigned(q) < Q) {
                        const double F = (dc == 0) ? sexph[q] : exp(-1.0 * double(d) * sh[q] / Nd);
                        acc[q] = acc[q] * seta[q] * F;
                        sum += acc[q];
                    } else {
                        acc[q] = 0.0;
                    }
                }
            }
            MsgVec<double, QT> mg;
            if (logdom) {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = (unsigned(q) < Q) ? exp(acc[q] - mx) : 0.0;
                    sum += mg.v[q];
                }
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = mg.v[q] / sum;
                    if (lane == 0) snum[q * TN + n] = acc[q] - mx;  // log of the unnormalised node total
                }
            } else {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = acc[q] / sum;
                    if (lane == 0) snum[q * TN + n] = mg.v[q];
                }
            }
            if (lane == 0) {
                const double w = (dc == 0) ? 1.0 : double(d);
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) wsum[q] += w * mg.v[q];
                mg.store(a.marg + size_t(n0 + n) * Q, Q);
            }
        }
        __syncthreads();

        // ---- phase 3: leave-one-out, normalise, max-diff, damped write -- in buffer order: in the bucketed layout
        // the positions of a tile's messages are contiguous per destination bucket, so consecutive lanes touch
        // consecutive addresses and a warp access spans a few 128-byte lines instead of 32.
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            if (u * kThreads + tid >= ne) continue;
            const unsigned k = pm[u];
            const size_t own = ps[u];
            const unsigned n = snode[k];
            const unsigned k0 = soff[n], d = soff[n + 1] - k0;
            const MsgVec<T, QT> &old = oldv[u];
            if ((a.clamp && d < kLargeDegree && a.clamp[n0 + n] != -1) || (a.color && a.color[n0 + n] != a.cur_color)) {
                // planted node / not this pass's colour: constant messages, no diff
                old.store(out_msg_dst<T>(a, Snew, own, Q), Q);
                continue;
            }
            T cav[QT];
            T s = T(0);
            if (d < kLargeDegree) {
                bool tiny = false;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const T bq = sb[q * TE + k];
                    if (unsigned(q) < Q) {
                        tiny = tiny || !(double(bq) >= kEps);
                        cav[q] = T(snum[q * TN + n]) / bq;
                    } else {
                        cav[q] = T(0);
                    }
                }
                if (tiny) {
                    // a vanishing b_e[q]: take the leave-one-out product directly instead of dividing.
                    // (The reference switches to a formula that drops eta_q * field here, :1029-1042, and for
                    // b == 0 reads stale scratch; such states are outside the parity contract -- they are
                    // counted in ctl->tiny_count and surfaced by sbmbp_tiny_events, see DESIGN.md.)
                    atomicAdd(&a.ctl->tiny_count, 1ull);
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        if (unsigned(q) < Q) {
                            double p = 1.0;
                            for (unsigned kk = k0; kk < k0 + d; ++kk)
                                if (kk != k) p *= double(sb[q * TE + kk]);
                            const double F = (dc == 0) ? sexph[q] : exp(-1.0 * double(d) * sh[q] / Nd);
                            cav[q] = T(p * seta[q] * F);
                        }
                    }
                }
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) s += cav[q];
            } else {
                double v[QT], mx = -1.0e300;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    if (unsigned(q) < Q) {
                        v[q] = snum[q * TN + n] - log(double(sb[q * TE + k]));  // :859
                        mx = fmax(mx, v[q]);
                    }
                }
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    cav[q] = (unsigned(q) < Q) ? T(exp(v[q] - mx)) : T(0);
                    s += cav[q];
                }
            }
            MsgVec<T, QT> out;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                const T nv = cav[q] / s;
                if (unsigned(q) < Q) {
                    const double df = fabs(double(old.v[q]) - double(nv));
                    mydiff = fmax(mydiff, df);
                    if (!(df == df)) ++mynan;
                }
                out.v[q] = T(a.damping) * nv + T(1.0 - a.damping) * old.v[q];
            }
            out.store(out_msg_dst<T>(a, Snew, own, Q), Q);
        }
    } else {
        // =================================================================== hub node (degree > TE)
        const unsigned long long d64 = ne64;
        const double dd = double(d64);
        if (a.color && a.color[n0] != a.cur_color) {  // not this pass's colour: carry the hub forward unchanged
            for (unsigned long long k = tid; k < d64; k += kThreads) {
                MsgVec<T, QT> old;
                const size_t own = size_t(__ldg(a.pos + e0 + k));
                old.load(out_msg_src<T>(a, Sold, own, Q), Q);
                old.store(out_msg_dst<T>(a, Snew, own, Q), Q);
            }
            if (tid == 0)
                for (unsigned q = 0; q < Q; ++q) wsum[q] += ((dc == 0) ? 1.0 : dd) * a.marg[size_t(n0) * Q + q];
            __syncthreads();  // matches the barrier of the regular path before the epilogue
        } else {
        double acc[QT];
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) acc[q] = 0.0;
        __syncthreads();  // parameters in smem
        for (unsigned long long k = tid; k < d64; k += kThreads) {
            MsgVec<T, QT> m;
            m.load(Sold + size_t(__ldg(a.rev + e0 + k)) * Q, Q);
            T b[QT];
            if (dc == 2) contract_dc2<T, QT>(m, sP, dd * double(__ldg(a.degsrc + e0 + k)), Q, b);
            else contract<T, QT>(m, sKl, b);
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q)
                if (unsigned(q) < Q) acc[q] += log(double(b[q]));
        }
        double mx = -1.0e300;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) {
            acc[q] = block_sum(acc[q], sred);
            if (unsigned(q) < Q) {
                acc[q] = acc[q] + slogeta[q] - ((dc == 0) ? sh[q] / Nd : 1.0 * dd * sh[q] / Nd);
                mx = fmax(mx, acc[q]);
            }
        }
        double sum = 0.0;
        MsgVec<double, QT> mg;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) {
            mg.v[q] = (unsigned(q) < Q) ? exp(acc[q] - mx) : 0.0;
            sum += mg.v[q];
        }
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) mg.v[q] /= sum;
        if (tid == 0) {
            const double w = (dc == 0) ? 1.0 : dd;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) wsum[q] += w * mg.v[q];
            mg.store(a.marg + size_t(n0) * Q, Q);
        }
        for (unsigned long long k = tid; k < d64; k += kThreads) {
            MsgVec<T, QT> m, old;
            m.load(Sold + size_t(__ldg(a.rev + e0 + k)) * Q, Q);
            const size_t own = size_t(__ldg(a.pos + e0 + k));
            old.load(out_msg_src<T>(a, Sold, own, Q), Q);
            T b[QT];
            if (dc == 2) contract_dc2<T, QT>(m, sP, dd * double(__ldg(a.degsrc + e0 + k)), Q, b);
            else contract<T, QT>(m, sKl, b);
            double v[QT], vmx = -1.0e300;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                if (unsigned(q) < Q) {
                    v[q] = (acc[q] - mx) - log(double(b[q]));
                    vmx = fmax(vmx, v[q]);
                }
            }
            T cav[QT], s = T(0);
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                cav[q] = (unsigned(q) < Q) ? T(exp(v[q] - vmx)) : T(0);
                s += cav[q];
            }
            MsgVec<T, QT> out;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                const T nv = cav[q] / s;
                if (unsigned(q) < Q) {
                    const double df = fabs(double(old.v[q]) - double(nv));
                    mydiff = fmax(mydiff, df);
                    if (!(df == df)) ++mynan;
                }
                out.v[q] = T(a.damping) * nv + T(1.0 - a.damping) * old.v[q];
            }
            out.store(out_msg_dst<T>(a, Snew, own, Q), Q);
        }
        }  // hub of this pass's colour
    }

    // ---- CTA epilogue: reduce the field partials and the max-diff over the CTA, store one row, done
    // (bp_finalize_kernel closes the sweep).  Non-finite values count as a huge difference so they cannot hide.
    if (mynan) mydiff = 1.0e300;
    mydiff = warp_max(mydiff);
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = warp_sum(wsum[q]);
    __syncthreads();  // sred may still be in use by the hub path's reductions
    if (lane == 0) {
        sred[warp * (QT + 1) + QT] = mydiff;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) sred[warp * (QT + 1) + q] = wsum[q];
    }
    __syncthreads();
    if (tid <= QT) {  // fixed order over the warps: bitwise reproducible
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w)
            v = (tid < QT) ? v + sred[w * (QT + 1) + tid] : fmax(v, sred[w * (QT + 1) + tid]);
        a.partial[size_t(blockIdx.x) * (QT + 1) + tid] = v;
    }
}

}  // namespace sbmbp
