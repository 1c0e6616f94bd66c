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
// (all degrees here are < 50).  Same arithmetic as sweep_tile.cuh: product in slot order, leave-one-out as a
// product for Q <= 4; a node one of whose b_l[q] underflows 1e-50 takes the exact leave-one-out product.
// Nodes of degree >= 32 are left to bp_sweep_warp_kernel / bp_sweep_hub_kernel (launched just before); each kernel
// leaves one row of field partials per CTA and the last CTA of this kernel closes the sweep in a fixed order.
//
// Why not everywhere: with many destination buckets the 32 out-messages of a (chunk, l) scatter over as many
// regions and the writes stop coalescing; from 8 buckets on the engine uses the CTA-tile kernels (sweep_tile.cuh).
#pragma once
#include "bp_device.cuh"
#include "sweep_tile.cuh"

namespace sbmbp {

#ifndef SBMBP_ELL_NT_WIDE
#define SBMBP_ELL_NT_WIDE 256  // Q = 2 in FP64: threads per CTA ...
#endif
#ifndef SBMBP_ELL_MINB_WIDE
#define SBMBP_ELL_MINB_WIDE 2  // ... and resident CTAs per SM the kernel is compiled for
#endif
#ifndef SBMBP_ELL_MINB
#define SBMBP_ELL_MINB 3
#endif
#ifndef SBMBP_ELL_MINB_COMPACT
#define SBMBP_ELL_MINB_COMPACT 2  // resident CTAs per SM of the compact-storage instantiation (Q = 2, FP64 arithmetic, 8-byte messages)
#endif
// Timing experiments (switching parts of the update off, per-warp timelines) exist only in builds made with
// -DSBMBP_TUNING (make EXTRA_NVFLAGS=-DSBMBP_TUNING): a production kernel carries no switch that can corrupt a result.
#ifdef SBMBP_TUNING
#define SBMBP_ELL_DBG(c, bit) (((c).dbg & (bit)) != 0u)
#else
#define SBMBP_ELL_DBG(c, bit) false
#endif
#ifndef SBMBP_ELL_BATCHED
#define SBMBP_ELL_BATCHED 0  // 1: degrees 5 .. DU keep b_l in a shared-memory slab instead of registers (measured: no gain)
#endif

#ifndef SBMBP_ELL_NOALLOC
#define SBMBP_ELL_NOALLOC 0
#endif
// The random gather: read-only path, no L1 allocation.  A gathered line is used once per SM, and letting it allocate
// evicts what the L1 is needed for here (register spill slots: measured, 90 % of local loads missed the L1).
template <typename T, int QT>
__device__ __forceinline__ void ld_gather_vec(MsgVec<T, QT> &m, const T *__restrict__ p) {
#if SBMBP_ELL_NOALLOC
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
        uint4 *d = reinterpret_cast<uint4 *>(m.v);
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(d[i].x), "=r"(d[i].y), "=r"(d[i].z), "=r"(d[i].w)
                         : "l"(reinterpret_cast<const char *>(p) + 16 * i));
    } else {
        static_assert(bytes == 8, "Q x sizeof(T) must be 8 or a multiple of 16");
        uint2 *d = reinterpret_cast<uint2 *>(m.v);
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(d->x), "=r"(d->y) : "l"(p));
    }
#else
    ld_vec<T, QT>(m, p);
#endif
}

template <typename T, int QT>
__device__ __forceinline__ void sts_own(T *p, const T (&v)[QT]) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i) reinterpret_cast<uint4 *>(p)[i] = reinterpret_cast<const uint4 *>(v)[i];
    } else {
        *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(v);
    }
}
template <typename T, int QT>
__device__ __forceinline__ void lds_own(T (&v)[QT], const T *p) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i) reinterpret_cast<uint4 *>(v)[i] = reinterpret_cast<const uint4 *>(p)[i];
    } else {
        *reinterpret_cast<uint2 *>(v) = *reinterpret_cast<const uint2 *>(p);
    }
}

template <typename T, int QT>
struct EllUnroll {
    static constexpr int DU = (QT * int(sizeof(T)) <= 16) ? 8 : 4;  // largest degree on the unrolled paths
    // resident CTAs per SM the kernel is compiled for: three with 8-byte messages (Q = 2 FP32 fits 80 registers), two
    // otherwise -- measured: the 80-register FP64 build spills 12 registers and runs 70 % slower, because the gathers
    // leave the L1 no room for spill slots (profiles/ell_investigation_r01.md)
    static constexpr bool kD2 = QT == 2 && sizeof(T) == 8;  // the configuration the block-size experiments are about
    static constexpr int MINB = (QT * int(sizeof(T)) <= 8) ? SBMBP_ELL_MINB : (kD2 ? SBMBP_ELL_MINB_WIDE : 2);
    // threads per CTA.  Registers are handed out per warp, so what the block size decides is the granularity of
    // residency: at ~100 registers an SM holds 20 warps as five 4-warp CTAs but only 16 as two 8-warp CTAs.
    static constexpr int NT = kD2 ? SBMBP_ELL_NT_WIDE : 256;
};

// dynamic shared memory of the kernel: index words staged one chunk ahead, and the b-slab of degrees 5 .. DU
template <typename T, int QT>
struct EllSmem {
    static constexpr int NW = EllUnroll<T, QT>::NT / 32;
    static constexpr int DU = EllUnroll<T, QT>::DU;
    static constexpr int SW = 2 * DU + 1;                                       // rev words | pos words | node
    static constexpr size_t off_idx = 0;                                        // u32[NW][2][SW][32]
    static constexpr size_t off_slab = off_idx + sizeof(unsigned) * NW * 2 * SW * 32;  // T[NW][DU * 32 * QT]
    static constexpr size_t bytes = off_slab + ((DU > 4 && SBMBP_ELL_BATCHED) ? sizeof(T) * NW * DU * 32 * QT : 0);
};

template <typename T>
struct EllSweepArgs {
    const uint4 *sched;        // [warps of the grid][sched_len]: x = index-array offset of the chunk, y = offset into
                               // ell_node, z = degree | lanes << 8 (lanes = 0: padding); built on the host so that every
                               // warp gets about the same work (engine.cu, build_ell_schedule)
    unsigned sched_len;
    const unsigned *ell_rev;   // per index word: buffer position of the in-message of that slot
    const unsigned *ell_pos;   // per index word: buffer position of its out-message
    const unsigned *ell_node;  // node ids, by class then ascending
    unsigned long long pf_bytes[3];  // bulk L2 prefetch at kernel start (0 = off): bytes of the source buffer, ell_rev, ell_pos
#ifdef SBMBP_TUNING
    unsigned long long *trace; // SBMBP_ELL_TRACE=1: 16 globaltimer stamps per warp, or nullptr
    unsigned dbg;              // SBMBP_ELL_DEBUG: 1 no old loads, 2 no message stores, 4 no marginal stores, 8 gathers replaced
                               // by a coalesced load -- results are wrong with any bit set
#endif
    T *S[2];
    double *marg;
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    double *partial;           // [gridDim.x + rows_before][QT + 1]
    unsigned rows_before;      // rows left after this kernel's own by the warp / hub kernels of the same sweep
    unsigned dc;
    double damping;
    // Lazy sweep close (a batch of sweeps launched back to back, this kernel alone carrying the sweep): sweep k of the
    // batch only leaves its rows (in the row buffer of parity k & 1) and sweep k + 1 turns them into the field and the
    // convergence decision in its own prologue -- every CTA redundantly, same order, same bits -- so the chain "fence,
    // ticket, last CTA reads the rows, exp, control block" (~7 us with the GPU idle) runs once per batch, not once per
    // sweep.  lazy_base: completed sweeps when the batch started (the host knows); lazy_last: this launch closes.
    int lazy;
    unsigned lazy_k, lazy_base;
    int lazy_last;
    int implicit_pos;  // the one-bucket padded layout (build_ell_padded_layout, engine.cu): no pos words, marginals chunk-ordered (marg = marg_ell)
};

// Compact storage of a normalised Q = 2 message (FP64): ONE double -- the smaller component, with the sign bit saying
// which one it is -- instead of two; the other component is 1 - it.  The small component is kept to full relative
// precision and the large one is within an ulp of what normalisation produced, so a round trip changes a message by
// <= 2.2e-16 absolute.  Halves the bytes of the gather, the old-value read and the store.
__device__ __forceinline__ void compact_decode(double x, double (&v)[2]) {
    const double s = fabs(x);
    const bool second = __double_as_longlong(x) < 0;  // sign bit (also on -0.0): the small component is psi_1
    v[0] = second ? 1.0 - s : s;
    v[1] = second ? s : 1.0 - s;
}
__device__ __forceinline__ double compact_encode(double v0, double v1) { return (v0 <= v1) ? v0 : -v1; }

// CP: compact storage (T = double, QT = 2 only)
template <typename T, int QT, bool CP>
struct EllCtx {
    static_assert(!CP || (QT == 2 && sizeof(T) == 8), "compact storage is the Q = 2 FP64 mode");
    const T *Sold;
    T *Snew;
    const unsigned *ell_rev;
    const unsigned *ell_pos;
    const T *K;         // QT x QT kernel matrix (shared memory)
    const double *eta;  // shared
    T damp, keep;
#ifdef SBMBP_TUNING
    unsigned dbg;
#endif
    unsigned long long *tiny_count;  // Ctl::tiny_count
    bool implicit_pos;  // one-bucket padded layout (build_ell_padded_layout): the out-message of index word w sits at position w

    __device__ __forceinline__ void load(MsgVec<T, QT> &m, unsigned pos) const {
        if constexpr (CP) compact_decode(__ldg(reinterpret_cast<const double *>(Sold) + pos), m.v);
        else ld_vec<T, QT>(m, Sold + size_t(pos) * QT);
    }
    __device__ __forceinline__ void gather(MsgVec<T, QT> &m, unsigned pos) const {
        if constexpr (CP) compact_decode(__ldg(reinterpret_cast<const double *>(Sold) + pos), m.v);
        else ld_gather_vec<T, QT>(m, Sold + size_t(pos) * QT);
    }
    __device__ __forceinline__ void store(const MsgVec<T, QT> &m, unsigned pos) const {
        if constexpr (CP) reinterpret_cast<double *>(Snew)[pos] = compact_encode(m.v[0], m.v[1]);
        else st_vec<T, QT>(m, Snew + size_t(pos) * QT);
    }
};

template <int QT>
struct EllOut {  // what a node update adds to the lane's running row
    double w[QT];
    double maxdiff;
};

// out-message of one slot from its leave-one-out vector: normalise, max |old - new|, damped write
template <typename T, int QT, bool CP>
__device__ __forceinline__ void ell_emit(const EllCtx<T, QT, CP> &c, const T (&cav_in)[QT], const MsgVec<T, QT> &oldv, unsigned p,
                                         double &mydiff) {
    T s = T(0);
#pragma unroll
    for (int q = 0; q < QT; ++q) s += cav_in[q];
    const T inv = fast_rcp(s);
    if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;  // non-finite message: make it visible
    MsgVec<T, QT> out;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        const T nv = cav_in[q] * inv;
        mydiff = fmax(mydiff, fabs(double(oldv.v[q]) - double(nv)));
        out.v[q] = c.damp * nv + c.keep * oldv.v[q];
    }
    if (!SBMBP_ELL_DBG(c, 2u)) c.store(out, p);
}

// node total (product of the b_l) -> normalised marginal, written out; tot becomes the marginal.
// F[q]: field factor of the class, exp(-d h_q / N) (dc) or exp(-beta h_q / N).
template <typename T, int QT, bool CP>
__device__ __forceinline__ void ell_node_total(const EllCtx<T, QT, CP> &c, const double *F, double wgt, double (&tot)[QT],
                                               double (&wsum)[QT], double *marg_out) {
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
        tot[q] = tot[q] * c.eta[q] * F[q];
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
    if (!SBMBP_ELL_DBG(c, 4u)) st_vec<double, QT>(mg, marg_out);
}

// Run-time degree: two passes, the second gathers again instead of keeping d vectors per thread.  Used for the
// classes above the unrolled ones and as the fallback of a node one of whose b_l[q] underflows 1e-50 (there the
// leave-one-out product is taken directly, see sweep_kernel.cuh).  Out of line: rare.
// ib: index word of slot 0 of this lane; slot l: ib + 32 l.
template <typename T, int QT, bool CP>
__device__ __noinline__ EllOut<QT> ell_update_loop(const EllCtx<T, QT, CP> c, const double *F, double wgt, unsigned d, unsigned ib,
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
        c.load(m, __ldg(c.ell_rev + ib + 32u * l));
        T b[QT];
        contract<T, QT>(m, c.K, b);
#pragma unroll
        for (int q = 0; q < QT; ++q) tot[q] *= double(b[q]);
    }
    ell_node_total<T, QT, CP>(c, F, wgt, tot, wsum, marg_out);
    for (unsigned l = 0; l < d; ++l) {
        const unsigned p = c.implicit_pos ? ib + 32u * l : __ldg(c.ell_pos + ib + 32u * l);
        MsgVec<T, QT> m, oldv;
        c.load(m, __ldg(c.ell_rev + ib + 32u * l));
        c.load(oldv, p);
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
            atomicAdd(c.tiny_count, 1ull);
            double pr[QT];
#pragma unroll
            for (int q = 0; q < QT; ++q) pr[q] = 1.0;
            for (unsigned l2 = 0; l2 < d; ++l2) {
                if (l2 == l) continue;
                MsgVec<T, QT> m2;
                c.load(m2, __ldg(c.ell_rev + ib + 32u * l2));
                T b2[QT];
                contract<T, QT>(m2, c.K, b2);
#pragma unroll
                for (int q = 0; q < QT; ++q) pr[q] *= double(b2[q]);
            }
#pragma unroll
            for (int q = 0; q < QT; ++q) cav[q] = T(pr[q] * c.eta[q] * F[q]);
        }
        ell_emit<T, QT, CP>(c, cav, oldv, p, mydiff);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) o.w[q] = wsum[q];
    o.maxdiff = mydiff;
    return o;
}

// degree D known at compile time: b_l in registers, everything unrolled.  The index words of this lane were staged
// in shared memory one chunk ahead (cp.async): sw[32 l] = rev word of slot l, sw[32 (DU + l)] = pos word.
template <typename T, int QT, bool CP, int D, int DU>
__device__ __forceinline__ void ell_update_fixed(const EllCtx<T, QT, CP> &c, const double *F, double wgt, const unsigned *sw,
                                                 unsigned ib, double *marg_out, double (&wsum)[QT], double &mydiff) {
    constexpr int DD = D > 0 ? D : 1;
    // old out-messages in two batches (slots 0-3, then 4-7): the first rides with the gathers, the second is issued
    // after the contraction so that at most 4 + D message vectors are live; the b_l stay in registers throughout
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
            g[l] = sw[32 * l];
            pw[l] = c.implicit_pos ? ib + 32u * unsigned(l) : sw[32 * (DU + l)];
            if (SBMBP_ELL_DBG(c, 8u)) g[l] = pw[l];
        }
        MsgVec<T, QT> m[D];
#pragma unroll
        for (int l = 0; l < D; ++l) c.gather(m[l], g[l]);
#pragma unroll
        for (int l = 0; l < N0; ++l) {
            if (!SBMBP_ELL_DBG(c, 1u)) c.load(old0[l], pw[l]);
            else
#pragma unroll
                for (int q = 0; q < QT; ++q) old0[l].v[q] = T(0.5);
        }
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
        const EllOut<QT> o = ell_update_loop<T, QT, CP>(c, F, wgt, unsigned(D), ib, marg_out);
#pragma unroll
        for (int q = 0; q < QT; ++q) wsum[q] += o.w[q];
        mydiff = fmax(mydiff, o.maxdiff);
        return;
    }
#pragma unroll
    for (int l = 0; l < N1; ++l) {
        if (!SBMBP_ELL_DBG(c, 1u)) c.load(old1[l], pw[N0 + l]);
        else
#pragma unroll
            for (int q = 0; q < QT; ++q) old1[l].v[q] = T(0.5);
    }
    ell_node_total<T, QT, CP>(c, F, wgt, tot, wsum, marg_out);
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
        ell_emit<T, QT, CP>(c, cav, old0[l], pw[l], mydiff);
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
        ell_emit<T, QT, CP>(c, cav, old1[l], pw[N0 + l], mydiff);
    }
}

// degrees 5 .. DU: same arithmetic, but the b_l wait in the lane's column of the warp's shared-memory slab instead of
// in registers (slot l of lane r at sbw[32 l QT]), and gathers / old values come four slots at a time -- the register
// budget of the degree-4 variant, so the kernel keeps three CTAs per SM without spilling
template <typename T, int QT, bool CP, int D, int DU>
__device__ __forceinline__ void ell_update_batched(const EllCtx<T, QT, CP> &c, const double *F, double wgt, const unsigned *sw,
                                                   unsigned ib, T *sbw, double *marg_out, double (&wsum)[QT], double &mydiff) {
    constexpr int NBT = (D + 3) / 4;
    double tot[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) tot[q] = 1.0;
    bool tiny = false;
#pragma unroll
    for (int bt = 0; bt < NBT; ++bt) {
        constexpr int kDummy = 0;
        (void)kDummy;
        const int n = (D - 4 * bt) < 4 ? (D - 4 * bt) : 4;
        MsgVec<T, QT> m[4];
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (l < n) c.gather(m[l], sw[32 * (4 * bt + l)]);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (l < n) {
                T b[QT];
                contract<T, QT>(m[l], c.K, b);
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    tot[q] *= double(b[q]);
                    tiny = tiny || !(double(b[q]) >= kEps);
                }
                sts_own<T, QT>(sbw + size_t(4 * bt + l) * 32 * QT, b);
            }
        }
    }
    if (tiny) {  // rare: the node goes through the general routine
        const EllOut<QT> o = ell_update_loop<T, QT, CP>(c, F, wgt, unsigned(D), ib, marg_out);
#pragma unroll
        for (int q = 0; q < QT; ++q) wsum[q] += o.w[q];
        mydiff = fmax(mydiff, o.maxdiff);
        return;
    }
    unsigned pw[4];
    MsgVec<T, QT> oldv[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        pw[l] = c.implicit_pos ? ib + 32u * unsigned(l) : sw[32 * (DU + l)];
        if (!SBMBP_ELL_DBG(c, 1u)) c.load(oldv[l], pw[l]);
        else
#pragma unroll
            for (int q = 0; q < QT; ++q) oldv[l].v[q] = T(0.5);
    }
    ell_node_total<T, QT, CP>(c, F, wgt, tot, wsum, marg_out);
#pragma unroll
    for (int bt = 0; bt < NBT; ++bt) {
        const int n = (D - 4 * bt) < 4 ? (D - 4 * bt) : 4;
        unsigned pn[4];
        MsgVec<T, QT> oldn[4];
        if (bt + 1 < NBT) {  // next batch's old values fly while this one is emitted
            const int nn = (D - 4 * (bt + 1)) < 4 ? (D - 4 * (bt + 1)) : 4;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                if (l < nn) {
                    pn[l] = c.implicit_pos ? ib + 32u * unsigned(4 * (bt + 1) + l) : sw[32 * (DU + 4 * (bt + 1) + l)];
                    if (!SBMBP_ELL_DBG(c, 1u)) c.load(oldn[l], pn[l]);
                    else
#pragma unroll
                        for (int q = 0; q < QT; ++q) oldn[l].v[q] = T(0.5);
                }
            }
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (l < n) {
                T b[QT], cav[QT];
                lds_own<T, QT>(b, sbw + size_t(4 * bt + l) * 32 * QT);
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    T v = T(tot[q]);
#pragma unroll
                    for (int r = 0; r < QT; ++r)
                        if (r != q) v *= b[r];
                    cav[q] = v;
                }
                ell_emit<T, QT, CP>(c, cav, oldv[l], pw[l], mydiff);
            }
        }
        if (bt + 1 < NBT) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                pw[l] = pn[l];
                oldv[l] = oldn[l];
            }
        }
    }
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// TMA bulk prefetch into the L2: does not occupy the LSU / L1 miss path the gathers depend on
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <typename T, int QT, bool CP = false>
#ifdef SBMBP_ELL_MAXNREG  // register-budget experiments (applies to every instantiation of the unit it is compiled in)
__global__ void __maxnreg__(SBMBP_ELL_MAXNREG) bp_sweep_ell_kernel(const EllSweepArgs<T> a) {
#else
__global__ void __launch_bounds__(EllUnroll<T, QT>::NT, CP ? SBMBP_ELL_MINB_COMPACT : EllUnroll<T, QT>::MINB) bp_sweep_ell_kernel(const EllSweepArgs<T> a) {
#endif
    static_assert(QT <= 4, "the degree-class kernel is the small-Q path");
    constexpr int NT = EllUnroll<T, QT>::NT;
    static_assert(NT >= int(kEllDegrees) * QT, "one thread per (degree, component) of the field table");
    constexpr int NW = NT / 32;
    constexpr int DU = EllUnroll<T, QT>::DU;  // degrees unrolled with b_l in registers
    __shared__ __align__(16) T s_K[QT * QT];
    __shared__ double s_eta[QT];
    __shared__ double s_F[kEllDegrees][QT];  // field factor per degree: exp(-d h_q / N) (dc) or exp(-beta h_q / N)
    __shared__ double s_rows[NW][QT + 1];
    // index words of a chunk, staged one chunk ahead per warp: [stage][rev words DU | pos words DU | node][lane];
    // b_l of degrees 5 .. DU, one column per lane
    constexpr int SW = 2 * DU + 1;
    using Lay = EllSmem<T, QT>;
    extern __shared__ __align__(16) unsigned char ell_smem[];
    unsigned(*s_idx)[2][SW][32] = reinterpret_cast<unsigned(*)[2][SW][32]>(ell_smem + Lay::off_idx);
    T *s_slab = reinterpret_cast<T *>(ell_smem + Lay::off_slab);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned gw = blockIdx.x * NW + warp;
    // Programmatic dependent launch (sweeps of one batch follow each other on the stream): the next sweep's CTAs may
    // become resident as this sweep's CTAs retire, run their prologue -- everything that does not depend on this
    // sweep's results -- and wait at griddepcontrol.wait for this grid to complete.  Both instructions are no-ops when
    // the launch does not carry the attribute.
    asm volatile("griddepcontrol.launch_dependents;");
#ifdef SBMBP_TUNING
    unsigned long long *trace = a.trace ? a.trace + size_t(gw) * 16 : nullptr;
#else
    unsigned long long *const trace = nullptr;
#endif
    if (trace && lane == 0) trace[0] = global_ns();
    // the warp's work list does not depend on the control block: first descriptors go out at once
    const unsigned len = a.sched_len;
    const uint4 *my = a.sched + size_t(gw) * len;
    const uint4 none = make_uint4(0u, 0u, 0u, 0u);
    uint4 d0 = len > 0 ? __ldg(my) : none;
    uint4 d1 = len > 1 ? __ldg(my + 1) : none;
    // model parameters: constant while sweeps are in flight
    double n_nodes = 1.0;
    if (unsigned(tid) < kEllDegrees * QT) n_nodes = a.prm->N;
    for (int i = tid; i < QT * QT; i += NT) s_K[i] = T(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) s_eta[tid] = a.prm->eta[tid];

    EllCtx<T, QT, CP> c;
    c.ell_rev = a.ell_rev;
    c.ell_pos = a.ell_pos;
    c.K = s_K;
    c.eta = s_eta;
    c.damp = T(a.damping);
    c.keep = T(1.0 - a.damping);
#ifdef SBMBP_TUNING
    c.dbg = a.dbg;
#endif
    c.tiny_count = &a.ctl->tiny_count;
    c.implicit_pos = a.implicit_pos != 0;
    const bool dc = a.dc != 0;

    // stage the index words of a chunk (degrees up to DU; higher degrees read them from global memory as they go)
    auto stage_idx = [&](const uint4 &ds, int st) {
        const unsigned d = ds.z & 0xffu, cnt = ds.z >> 8;
        if (unsigned(lane) < cnt) {
            if (!a.implicit_pos) cp_async4(&s_idx[warp][st][2 * DU][lane], a.ell_node + ds.y + lane);
            if (d <= unsigned(DU)) {
                const unsigned *rv = a.ell_rev + ds.x + lane, *pv = a.ell_pos + ds.x + lane;
#pragma unroll
                for (int l = 0; l < DU; ++l) {
                    if (unsigned(l) < d) {
                        cp_async4(&s_idx[warp][st][l][lane], rv + 32 * l);
                        if (!a.implicit_pos) cp_async4(&s_idx[warp][st][DU + l][lane], pv + 32 * l);
                    }
                }
            }
        }
        cp_async_commit();
    };
    stage_idx(d0, 0);

    // ---- from here on the previous sweep's results are needed: its messages, the field it published, the control block
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Ctl *ctl = a.ctl;
    const bool from_rows = a.lazy && a.lazy_k > 0;  // the previous sweep of the batch left rows, not a field
    double *my_rows = a.partial + size_t(a.lazy ? (a.lazy_k & 1u) : 0u) * (gridDim.x * (QT + 1));
    unsigned sweeps_done;
    if (!from_rows) {
        // the field of BOTH parities goes out before the control block is known: one round trip, not two
        double fh[2] = {0.0, 0.0}, fe[2] = {0.0, 0.0};
        if (unsigned(tid) < kEllDegrees * QT) {
            const unsigned q = tid % QT;
            fh[0] = a.field[0]->h[q];
            fh[1] = a.field[1]->h[q];
            fe[0] = a.field[0]->exph[q];
            fe[1] = a.field[1]->exph[q];
        }
        sweeps_done = ctl->sweeps_done;
        if (ctl->converged || sweeps_done >= ctl->max_sweeps) {  // uniform over the grid
            cp_async_wait_all();
            return;
        }
        if (unsigned(tid) < kEllDegrees * QT) {
            const unsigned d = tid / QT, q = tid % QT;
            const int p = int(sweeps_done & 1u);
            s_F[d][q] = (a.dc != 0) ? exp(-1.0 * double(d) * (p ? fh[1] : fh[0]) / n_nodes) : (p ? fe[1] : fe[0]);
        }
    } else {
        // close the previous sweep of the batch here: its rows -> totals -> h, exp(-beta h / N), max-diff -- the arithmetic
        // of close_sweep_last_cta, by every CTA (a few KB out of the L2).  Nothing below reads ctl->sweeps_done or a Field.
        __shared__ double s_tot[QT + 1];
        __shared__ double s_fld[2][QT];
        sweeps_done = a.lazy_base + a.lazy_k;
        const int conv = ctl->converged;
        const unsigned max_sweeps = ctl->max_sweeps;
        const float crit = ctl->crit;
        const unsigned sweep_base = ctl->sweep_base;
        double ccol[QT], beta = 0.0;
        if (tid < QT) {
#pragma unroll
            for (int t = 0; t < QT; ++t) ccol[t] = a.prm->C[t * kMaxQ + tid];
            beta = a.prm->beta;
        }
        const double *prev_rows = a.partial + size_t((a.lazy_k & 1u) ^ 1u) * (gridDim.x * (QT + 1));
        if (conv || sweeps_done >= max_sweeps) {  // uniform over the grid
            cp_async_wait_all();
            return;
        }
        reduce_rows_cta<QT, NT>(prev_rows, gridDim.x, s_tot);
        const double md = s_tot[QT];
        double h = 0.0, eh = 0.0;
        if (tid < QT) {
            h = field_component<QT>(ccol, s_tot);
            eh = exp(-beta * h / a.prm->N);
            s_fld[0][tid] = h;
            s_fld[1][tid] = eh;
        }
        const bool nan_md = !(md == md) || md > 1.0e299;
        if (blockIdx.x == 0 && tid == 0 && nan_md) ctl->nan_count += 1;
        if (md < crit) {  // double < float, as belief_propagation.cpp:406: the previous sweep converged -- uniform
            if (blockIdx.x == 0) {  // one CTA leaves what the close would have left
                Field *out = (sweeps_done & 1u) ? a.field[1] : a.field[0];
                if (tid < QT) {
                    out->h[tid] = h;
                    out->exph[tid] = eh;
                    out->wsum[tid] = s_tot[tid];
                }
                if (tid == 0) {
                    ctl->last_maxdiff = md;
                    ctl->sweeps_done = sweeps_done;
                    ctl->niter = int(sweeps_done - 1u - sweep_base);
                    ctl->converged = 1;
                }
            }
            cp_async_wait_all();
            return;
        }
        __syncthreads();
        if (unsigned(tid) < kEllDegrees * QT) {
            const unsigned d = tid / QT, q = tid % QT;
            s_F[d][q] = (a.dc != 0) ? exp(-1.0 * double(d) * s_fld[0][q] / n_nodes) : s_fld[1][q];
        }
    }
    const int par = int(sweeps_done & 1u);
    c.Sold = par ? a.S[1] : a.S[0];
    c.Snew = par ? a.S[0] : a.S[1];
    // Stream the sweep's sequential operands into the L2 up front, through the TMA unit: the source buffer (gathered at
    // random later, so its lines are wanted BEFORE their first gather) and the two index arrays.  The demand loads of the
    // SMs then see L2 latency, which is what the number of misses an SM can keep in flight is divided by.
    {
        constexpr unsigned kPiece = 2048u;
        const unsigned nwt_pf = gridDim.x * NW;
        const void *src[3] = {c.Sold, a.ell_rev, a.ell_pos};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned long long bytes = a.pf_bytes[k];
            const unsigned long long npieces = (bytes + kPiece - 1) / kPiece;
            for (unsigned long long pc = gw + (unsigned long long)lane * nwt_pf; pc < npieces; pc += 32ull * nwt_pf) {
                const unsigned long long off = pc * kPiece;
                const unsigned sz = unsigned(bytes - off < kPiece ? ((bytes - off) & ~15ull) : kPiece);
                if (sz) bulk_prefetch_l2(reinterpret_cast<const char *>(src[k]) + off, sz);
            }
        }
    }
    __syncthreads();  // parameters in shared memory
    if (trace && lane == 0) trace[1] = global_ns();

    double wsum[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;
    unsigned tslot = 2;

    int st = 0;
    for (unsigned i = 0; i < len; ++i, st ^= 1) {
        // ---- keep the loads fed: descriptor two chunks ahead, index words of the next chunk -> the other stage
        const uint4 d2 = (i + 2u < len) ? __ldg(my + i + 2u) : none;
        stage_idx(d1, st ^ 1);
        cp_async_wait_group1();  // this chunk's index words (staged one chunk ago; each lane reads back its own) are in

        const unsigned d = d0.z & 0xffu, cnt = d0.z >> 8;
        if (unsigned(lane) < cnt) {
            const unsigned ib = d0.x + unsigned(lane);
            const unsigned *sw = &s_idx[warp][st][0][lane];
            double *mo = a.marg + size_t(a.implicit_pos ? d0.y + unsigned(lane) : sw[32 * 2 * DU]) * QT;
            const double *F = s_F[d];
            const double wgt = dc ? double(d) : 1.0;
            bool done = true;
            switch (d) {
                case 0: ell_update_fixed<T, QT, CP, 0, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                case 1: ell_update_fixed<T, QT, CP, 1, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                case 2: ell_update_fixed<T, QT, CP, 2, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                case 3: ell_update_fixed<T, QT, CP, 3, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                case 4: ell_update_fixed<T, QT, CP, 4, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                default: done = false; break;
            }
            if constexpr (DU >= 8) {
                if (!done) {
                    T *sbw = s_slab + size_t(warp) * DU * 32 * QT + lane * QT;
                    (void)sbw;
                    done = true;
                    switch (d) {
#if SBMBP_ELL_BATCHED
                        case 5: ell_update_batched<T, QT, CP, 5, DU>(c, F, wgt, sw, ib, sbw, mo, wsum, mydiff); break;
                        case 6: ell_update_batched<T, QT, CP, 6, DU>(c, F, wgt, sw, ib, sbw, mo, wsum, mydiff); break;
                        case 7: ell_update_batched<T, QT, CP, 7, DU>(c, F, wgt, sw, ib, sbw, mo, wsum, mydiff); break;
                        case 8: ell_update_batched<T, QT, CP, 8, DU>(c, F, wgt, sw, ib, sbw, mo, wsum, mydiff); break;
#else
                        case 5: ell_update_fixed<T, QT, CP, 5, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                        case 6: ell_update_fixed<T, QT, CP, 6, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                        case 7: ell_update_fixed<T, QT, CP, 7, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
                        case 8: ell_update_fixed<T, QT, CP, 8, DU>(c, F, wgt, sw, ib, mo, wsum, mydiff); break;
#endif
                        default: done = false; break;
                    }
                }
            }
            if (!done) {
                const EllOut<QT> o = ell_update_loop<T, QT, CP>(c, F, wgt, d, ib, mo);
#pragma unroll
                for (int q = 0; q < QT; ++q) wsum[q] += o.w[q];
                mydiff = fmax(mydiff, o.maxdiff);
            }
        }
        if (trace && lane == 0 && tslot < 14) trace[tslot++] = global_ns();
        d0 = d1;
        d1 = d2;
    }
    cp_async_wait_all();
    if (trace && lane == 0) trace[14] = global_ns();

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
        my_rows[size_t(blockIdx.x) * (QT + 1) + tid] = v;
    }
    if (a.lazy && !a.lazy_last) {  // the next sweep of the batch closes this one (kernel boundary: no fence, no ticket)
        if (trace && lane == 0) trace[15] = global_ns();
        return;
    }
    SweepArgsBase base;
    base.prm = a.prm;
    base.field[0] = a.field[0];
    base.field[1] = a.field[1];
    base.ctl = a.ctl;
    base.partial = my_rows;
    close_sweep_last_cta<QT, NT>(base, gridDim.x + a.rows_before, sweeps_done, nullptr, gridDim.x);
    if (trace && lane == 0) trace[15] = global_ns();
}

}  // namespace sbmbp
