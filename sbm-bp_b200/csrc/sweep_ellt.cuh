// Degree-class sweep kernel with TMA data movement ("ELL-T"): the small-Q path for graphs whose message buffers fit the
// L2 (BASELINE configs[1]: 48 MB per buffer).  Same unit of work as sweep_ell.cuh -- one warp per chunk of 32 nodes of
// equal degree, one thread per node -- but a layout in which everything except the gather is one contiguous block per
// chunk, moved by the TMA unit (cp.async.bulk, mbarrier-completed) instead of by per-thread loads and stores:
//
//   * ONE destination bucket: the out-message of slot l of lane r of a chunk sits at  chunk base + 32 l + r  of the
//     message buffer (partial chunks are padded to 32 lanes), which is also the offset of that slot's index word.  So
//       - the old out-messages of a chunk are ONE block of d x 32 messages          -> one bulk load  (global -> shared)
//       - its new out-messages, written in place over the old ones in shared memory  -> one bulk store (shared -> global)
//       - its gather index words (rev) are one block of d x 128 bytes                -> one bulk load
//       - its marginals go to a chunk-ordered array (marg_ell), 32 x Q doubles       -> one bulk store
//     and no `pos` array exists.  None of this traffic touches the LSU / L1 miss path any more.
//   * The one random access, the gather of the in-messages, is issued ONE UNIT AHEAD with per-lane cp.async (LDGSTS,
//     16 / 8 bytes, L1 bypass) into shared memory, so a warp always has the gathers of its next chunk in flight while
//     it computes the current one and no message vector is ever parked in registers: the kernel runs at ~60
//     registers and its residency is set by shared memory (3 CTAs of 4 warps per SM).
//   * Three-deep software pipeline per warp, warp-private (no CTA barrier in the loop): index words two units ahead,
//     gathers + old values one unit ahead, compute + stores now.  Completion: one mbarrier per stage for the bulk loads
//     (expect_tx), cp.async groups for the gathers (every lane reads back only what it copied), bulk groups for the
//     stores (wait_group.read before a staging buffer is reused).
//
// Arithmetic is that of sweep_ell.cuh, operation for operation (product in slot order, leave-one-out as a product for
// Q <= 4, exact leave-one-out product where a b_l[q] underflows 1e-50), so messages and marginals are bit-identical to
// it.  Reference: sum_all_messages_to_i / norm_m_at_i (belief_propagation.cpp:991-1071), synchronous, product domain
// (degrees < 32 here; higher degrees go to bp_sweep_warp_kernel / bp_sweep_hub_kernel, launched just before).
#pragma once
#include "bp_device.cuh"
#include "sweep_ell.cuh"

namespace sbmbp {

template <typename T, int QT>
struct ElltCfg {
    static constexpr int MB = QT * int(sizeof(T));  // bytes per message: 8 or 16 on this path
    static constexpr int DS = (MB <= 8) ? 8 : 6;    // largest degree staged through shared memory; above: direct loads
    static constexpr int NW = 4;                    // warps per CTA
    static constexpr int NT = NW * 32;
    static constexpr int kRevStages = 2, kGatStages = 2, kOutStages = 3, kMargStages = 2;
    static constexpr size_t rev_bytes = size_t(DS) * 128;            // one stage of index words
    static constexpr size_t msg_bytes = size_t(DS) * 32 * MB;        // one stage of messages (gathered, or old -> new)
    static constexpr size_t marg_bytes = size_t(32) * QT * 8;        // one stage of marginals
    static constexpr size_t off_rev = 0;
    static constexpr size_t off_gat = off_rev + kRevStages * rev_bytes;
    static constexpr size_t off_out = off_gat + kGatStages * msg_bytes;
    static constexpr size_t off_marg = off_out + kOutStages * msg_bytes;
    static constexpr size_t off_bar = off_marg + kMargStages * marg_bytes;  // u64[kRevStages + kOutStages]
    static constexpr size_t warp_bytes = (off_bar + 8 * (kRevStages + kOutStages) + 127) & ~size_t(127);
    static constexpr size_t bytes = warp_bytes * NW;
};

template <typename T>
struct ElltSweepArgs {
    const uint4 *sched;       // [warps of the grid][sched_len]: x = chunk base (buffer position == index-word offset),
                              // y = first entry of the chunk in marg_ell / ell_node (32 per chunk), z = degree | lanes << 8
    unsigned sched_len;
    const unsigned *ell_rev;  // per index word: buffer position of the in-message of that slot
    T *S[2];
    double *marg_ell;         // marginals in chunk order (32 x Q per chunk)
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    double *partial;          // [gridDim.x + rows_before][QT + 1]
    unsigned rows_before;     // rows left by the warp / hub kernels of the same sweep
    unsigned dc;
    double damping;
};

// ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX; SASS: SYNCS, UBLKCP)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion (bytes) on an mbarrier of this CTA
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // all but the N latest groups have finished READING shared memory
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (the TMA unit) that reads them next
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// per-lane gather of one message into shared memory, bypassing registers (and, for 16 bytes, the L1)
template <int MB>
__device__ __forceinline__ void cp_async_msg(void *smem_dst, const void *gsrc) {
    static_assert(MB == 8 || MB == 16, "message size on the ELL-T path");
    if constexpr (MB == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <typename T, int QT>
__device__ __forceinline__ void lds_msg(MsgVec<T, QT> &m, const T *p) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes == 16) *reinterpret_cast<uint4 *>(m.v) = *reinterpret_cast<const uint4 *>(p);
    else *reinterpret_cast<uint2 *>(m.v) = *reinterpret_cast<const uint2 *>(p);
}
template <typename T, int QT>
__device__ __forceinline__ void sts_msg(T *p, const MsgVec<T, QT> &m) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes == 16) *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(m.v);
    else *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(m.v);
}

// A chunk of degree > DS (rare: 3 % of the nodes of a c = 3 graph): two passes over direct loads, direct stores.
// Out of line.  base: buffer position / index-word offset of slot 0 of this lane; slot l: base + 32 l.
template <typename T, int QT>
__device__ __noinline__ EllOut<QT> ellt_update_direct(const T *Sold, T *Snew, const unsigned *ell_rev, const T *K, const double *eta,
                                                      const double *F, double wgt, T damp, T keep, unsigned d, unsigned base,
                                                      double *marg_out, unsigned long long *tiny_count) {
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
        ld_vec<T, QT>(m, Sold + size_t(__ldg(ell_rev + base + 32u * l)) * QT);
        T b[QT];
        contract<T, QT>(m, K, b);
#pragma unroll
        for (int q = 0; q < QT; ++q) tot[q] *= double(b[q]);
    }
    {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            tot[q] = tot[q] * eta[q] * F[q];
            sum += tot[q];
        }
        MsgVec<double, QT> mg;
        const double rsum = fast_rcp(sum);
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            mg.v[q] = tot[q] * rsum;
            tot[q] = mg.v[q];
            wsum[q] += wgt * mg.v[q];
        }
        st_vec<double, QT>(mg, marg_out);
    }
    for (unsigned l = 0; l < d; ++l) {
        const unsigned p = base + 32u * l;
        MsgVec<T, QT> m, oldv;
        ld_vec<T, QT>(m, Sold + size_t(__ldg(ell_rev + p)) * QT);
        ld_vec<T, QT>(oldv, Sold + size_t(p) * QT);
        T b[QT], cav[QT];
        contract<T, QT>(m, K, b);
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
            atomicAdd(tiny_count, 1ull);
            double pr[QT];
#pragma unroll
            for (int q = 0; q < QT; ++q) pr[q] = 1.0;
            for (unsigned l2 = 0; l2 < d; ++l2) {
                if (l2 == l) continue;
                MsgVec<T, QT> m2;
                ld_vec<T, QT>(m2, Sold + size_t(__ldg(ell_rev + base + 32u * l2)) * QT);
                T b2[QT];
                contract<T, QT>(m2, K, b2);
#pragma unroll
                for (int q = 0; q < QT; ++q) pr[q] *= double(b2[q]);
            }
#pragma unroll
            for (int q = 0; q < QT; ++q) cav[q] = T(pr[q] * eta[q] * F[q]);
        }
        T s = T(0);
#pragma unroll
        for (int q = 0; q < QT; ++q) s += cav[q];
        const T inv = fast_rcp(s);
        if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;
        MsgVec<T, QT> out;
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            const T nv = cav[q] * inv;
            mydiff = fmax(mydiff, fabs(double(oldv.v[q]) - double(nv)));
            out.v[q] = damp * nv + keep * oldv.v[q];
        }
        st_vec<T, QT>(out, Snew + size_t(p) * QT);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) o.w[q] = wsum[q];
    o.maxdiff = mydiff;
    return o;
}

template <typename T, int QT>
__global__ void __launch_bounds__(ElltCfg<T, QT>::NT, 3) bp_sweep_ellt_kernel(const ElltSweepArgs<T> a) {
    using Cfg = ElltCfg<T, QT>;
    constexpr int NT = Cfg::NT, NW = Cfg::NW, DS = Cfg::DS, MB = Cfg::MB;
    static_assert(QT <= 4 && (MB == 8 || MB == 16), "the ELL-T kernel is the 8 / 16-byte-message path");
    static_assert(NT >= int(kEllDegrees) * QT, "one thread per (degree, component) of the field table");
    __shared__ __align__(16) T s_K[QT * QT];
    __shared__ double s_eta[QT];
    __shared__ double s_F[kEllDegrees][QT];  // field factor per degree: exp(-d h_q / N) (dc) or exp(-beta h_q / N)
    __shared__ double s_rows[NW][QT + 1];
    extern __shared__ __align__(128) unsigned char ellt_smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned gw = blockIdx.x * NW + warp;
    unsigned char *wsm = ellt_smem + size_t(warp) * Cfg::warp_bytes;
    unsigned *rbuf = reinterpret_cast<unsigned *>(wsm + Cfg::off_rev);   // [kRevStages][DS][32]
    T *gbuf = reinterpret_cast<T *>(wsm + Cfg::off_gat);                 // [kGatStages][DS][32][QT]
    T *obuf = reinterpret_cast<T *>(wsm + Cfg::off_out);                 // [kOutStages][DS][32][QT]
    double *mbuf = reinterpret_cast<double *>(wsm + Cfg::off_marg);      // [kMargStages][32][QT]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(wsm + Cfg::off_bar);
    unsigned long long *rev_bar = bars, *old_bar = bars + Cfg::kRevStages;
    constexpr unsigned kRevWords = DS * 32, kMsgElems = DS * 32 * QT;

    // Programmatic dependent launch: everything up to griddepcontrol.wait is independent of the previous sweep
    asm volatile("griddepcontrol.launch_dependents;");
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < Cfg::kRevStages + Cfg::kOutStages; ++i) mbar_init(bars + i, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const unsigned len = a.sched_len;
    const uint4 *my = a.sched + size_t(gw) * len;
    const uint4 none = make_uint4(0u, 0u, 0u, 0u);
    uint4 d0 = len > 0 ? __ldg(my) : none;       // unit u
    uint4 d1 = len > 1 ? __ldg(my + 1) : none;   // unit u + 1
    uint4 d2 = len > 2 ? __ldg(my + 2) : none;   // unit u + 2

    double n_nodes = 1.0;
    if (unsigned(tid) < kEllDegrees * QT) n_nodes = a.prm->N;
    for (int i = tid; i < QT * QT; i += NT) s_K[i] = T(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) s_eta[tid] = a.prm->eta[tid];

    auto staged = [&](const uint4 &ds) {
        const unsigned d = ds.z & 0xffu;
        return (ds.z >> 8) != 0u && d >= 1u && d <= unsigned(DS);
    };
    // stage A: index words of a unit -> rbuf[st] (one bulk load)
    auto stage_rev = [&](const uint4 &ds, unsigned st) {
        if (staged(ds) && lane == 0) {
            const unsigned bytes = (ds.z & 0xffu) * 128u;
            mbar_expect_tx(rev_bar + st, bytes);
            bulk_load(rbuf + st * kRevWords, a.ell_rev + ds.x, bytes, rev_bar + st);
        }
    };
    stage_rev(d0, 0u);
    stage_rev(d1, 1u);

    // ---- from here on the previous sweep's results are needed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Ctl *ctl = a.ctl;
    {
        double fh[2] = {0.0, 0.0}, fe[2] = {0.0, 0.0};
        if (unsigned(tid) < kEllDegrees * QT) {
            const unsigned q = tid % QT;
            fh[0] = a.field[0]->h[q];
            fh[1] = a.field[1]->h[q];
            fe[0] = a.field[0]->exph[q];
            fe[1] = a.field[1]->exph[q];
        }
        const unsigned sd = ctl->sweeps_done;
        if (ctl->converged || sd >= ctl->max_sweeps) {  // uniform over the grid; the bulk loads in flight must land first
            __syncwarp();
            if (staged(d0)) mbar_wait(rev_bar + 0, 0u);
            if (staged(d1)) mbar_wait(rev_bar + 1, 0u);
            return;
        }
        if (unsigned(tid) < kEllDegrees * QT) {
            const unsigned d = tid / QT, q = tid % QT;
            const int p = int(sd & 1u);
            s_F[d][q] = (a.dc != 0) ? exp(-1.0 * double(d) * (p ? fh[1] : fh[0]) / n_nodes) : (p ? fe[1] : fe[0]);
        }
    }
    const unsigned sweeps_done = ctl->sweeps_done;
    const int par = int(sweeps_done & 1u);
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const T damp = T(a.damping), keep = T(1.0 - a.damping);
    const bool dc = a.dc != 0;
    __syncthreads();  // parameters in shared memory

    unsigned rph = 0u, oph = 0u;  // phase bits of the mbarriers (bit s = parity the next wait on stage s expects)
    // stage B: gathers of a unit -> gbuf[gs] (per-lane cp.async), its old out-messages -> obuf[os] (one bulk load)
    auto stage_msgs = [&](const uint4 &ds, unsigned rs, unsigned gs, unsigned os) {
        if (staged(ds)) {
            const unsigned d = ds.z & 0xffu;
            mbar_wait(rev_bar + rs, (rph >> rs) & 1u);
            rph ^= 1u << rs;
            const unsigned *rw = rbuf + rs * kRevWords + lane;
            T *g = gbuf + size_t(gs) * kMsgElems + lane * QT;
#pragma unroll
            for (int l = 0; l < DS; ++l)
                if (unsigned(l) < d) cp_async_msg<MB>(g + l * 32 * QT, Sold + size_t(rw[32 * l]) * QT);
            if (lane == 0) {
                const unsigned bytes = d * 32u * unsigned(MB);
                mbar_expect_tx(old_bar + os, bytes);
                bulk_load(obuf + size_t(os) * kMsgElems, Sold + size_t(ds.x) * QT, bytes, old_bar + os);
            }
        }
        cp_async_commit();  // one group per unit, empty or not: wait_group counts units
    };
    stage_msgs(d0, 0u, 0u, 0u);

    double wsum[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;

    for (unsigned u = 0; u < len; ++u) {
        const unsigned rs = u & 1u, gs = u & 1u, os = u % 3u, ms = u & 1u;
        // the stores of unit u - 2 have left shared memory: mbuf[ms] and obuf[(u + 1) % 3] are free again
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        // ---- keep the pipeline fed: gathers + old values of unit u + 1, index words of unit u + 2
        stage_msgs(d1, rs ^ 1u, gs ^ 1u, (u + 1u) % 3u);
        const uint4 d3 = (u + 3u < len) ? __ldg(my + u + 3u) : none;
        __syncwarp();  // every lane has read its words of rbuf[rs] (stage B of unit u, one iteration ago)
        stage_rev(d2, rs);
        cp_async_wait_group1();  // this lane's gathers of unit u have landed

        const unsigned d = d0.z & 0xffu, cnt = d0.z >> 8;
        if (cnt != 0u) {
            const double *F = s_F[d];
            const double wgt = dc ? double(d) : 1.0;
            if (d <= unsigned(DS)) {
                if (d != 0u) {
                    mbar_wait(old_bar + os, (oph >> os) & 1u);
                    oph ^= 1u << os;
                }
                if (unsigned(lane) < cnt) {
                    const T *g = gbuf + size_t(gs) * kMsgElems + lane * QT;
                    T *o = obuf + size_t(os) * kMsgElems + lane * QT;
                    double tot[QT];
#pragma unroll
                    for (int q = 0; q < QT; ++q) tot[q] = 1.0;
                    for (unsigned l = 0; l < d; ++l) {
                        MsgVec<T, QT> m;
                        lds_msg<T, QT>(m, g + l * 32 * QT);
                        T b[QT];
                        contract<T, QT>(m, s_K, b);
#pragma unroll
                        for (int q = 0; q < QT; ++q) tot[q] *= double(b[q]);
                    }
                    {
                        double sum = 0.0;
#pragma unroll
                        for (int q = 0; q < QT; ++q) {
                            tot[q] = tot[q] * s_eta[q] * F[q];
                            sum += tot[q];
                        }
                        const double rsum = fast_rcp(sum);
                        MsgVec<double, QT> mg;
#pragma unroll
                        for (int q = 0; q < QT; ++q) {
                            mg.v[q] = tot[q] * rsum;
                            tot[q] = mg.v[q];
                            wsum[q] += wgt * mg.v[q];
                        }
                        double *mo = mbuf + size_t(ms) * 32 * QT + lane * QT;
#pragma unroll
                        for (int q = 0; q < QT; q += 2) *reinterpret_cast<double2 *>(mo + q) = make_double2(mg.v[q], mg.v[q + 1]);
                    }
                    for (unsigned l = 0; l < d; ++l) {
                        MsgVec<T, QT> m, oldv;
                        lds_msg<T, QT>(m, g + l * 32 * QT);
                        lds_msg<T, QT>(oldv, o + l * 32 * QT);
                        T b[QT], cav[QT];
                        contract<T, QT>(m, s_K, b);
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
                        } else {  // a vanishing b_l[q]: the exact leave-one-out product (see sweep_kernel.cuh); rare
                            atomicAdd(&ctl->tiny_count, 1ull);
                            double pr[QT];
#pragma unroll
                            for (int q = 0; q < QT; ++q) pr[q] = 1.0;
                            for (unsigned l2 = 0; l2 < d; ++l2) {
                                if (l2 == l) continue;
                                MsgVec<T, QT> m2;
                                lds_msg<T, QT>(m2, g + l2 * 32 * QT);
                                T b2[QT];
                                contract<T, QT>(m2, s_K, b2);
#pragma unroll
                                for (int q = 0; q < QT; ++q) pr[q] *= double(b2[q]);
                            }
#pragma unroll
                            for (int q = 0; q < QT; ++q) cav[q] = T(pr[q] * s_eta[q] * F[q]);
                        }
                        T s = T(0);
#pragma unroll
                        for (int q = 0; q < QT; ++q) s += cav[q];
                        const T inv = fast_rcp(s);
                        if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;  // non-finite message: make it visible
                        MsgVec<T, QT> out;
#pragma unroll
                        for (int q = 0; q < QT; ++q) {
                            const T nv = cav[q] * inv;
                            mydiff = fmax(mydiff, fabs(double(oldv.v[q]) - double(nv)));
                            out.v[q] = damp * nv + keep * oldv.v[q];
                        }
                        sts_msg<T, QT>(o + l * 32 * QT, out);
                    }
                }
                // new messages and marginals of the chunk -> global memory, one bulk store each
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (d != 0u) bulk_store(Snew + size_t(d0.x) * QT, obuf + size_t(os) * kMsgElems, d * 32u * unsigned(MB));
                    bulk_store(a.marg_ell + size_t(d0.y) * QT, mbuf + size_t(ms) * 32 * QT, 32u * QT * 8u);
                }
            } else if (unsigned(lane) < cnt) {
                const EllOut<QT> o = ellt_update_direct<T, QT>(Sold, Snew, a.ell_rev, s_K, s_eta, F, wgt, damp, keep, d, d0.x + unsigned(lane),
                                                               a.marg_ell + size_t(d0.y + unsigned(lane)) * QT, &ctl->tiny_count);
#pragma unroll
                for (int q = 0; q < QT; ++q) wsum[q] += o.w[q];
                mydiff = fmax(mydiff, o.maxdiff);
            }
        }
        if (lane == 0) bulk_commit();  // one bulk group per unit, empty or not
        d0 = d1;
        d1 = d2;
        d2 = d3;
    }
    cp_async_wait_all();
    if (lane == 0) bulk_wait_all();  // this warp's stores are complete before its row is published

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
    close_sweep_last_cta<QT, NT>(base, gridDim.x + a.rows_before, sweeps_done, nullptr, gridDim.x);
}

}  // namespace sbmbp
