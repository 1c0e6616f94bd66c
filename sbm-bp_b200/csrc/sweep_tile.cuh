// CTA-tile sweep for the common case: Q equals the compiled width QT, deg_corr_flag 0 or 1, and one Q x Q kernel for
// every degree class (beta == 1, or dc == 1 where beta never enters).  Same algorithm and tiles as bp_sweep_kernel
// (sweep_kernel.cuh) -- that one stays as the general path (padded Q, dc == 2, beta != 1) -- with the run-time
// generality stripped from the per-edge code:
//   * per buffer entry ONE packed word says which tile-local slot and node it belongs to and whether that node
//     updates in the log domain, so phase 3 needs no edge->node table and no degree lookups
//   * the leave-one-out division becomes a product for Q <= 4:  psi~_q / b_q  ~  psi~_q * prod_{q' != q} b_q'
//     (the common factor prod_q b_q cancels in the normalisation), leaving one reciprocal per edge instead of
//     Q + Q long FP64 division chains
//   * the rare per-entry paths of phase 3 (log-domain node, vanishing b_e) are out of line
// Reference: sum_all_messages_to_i / norm_m_at_i / bp_iter_update_psi_large_degree
// (belief_propagation.cpp:991-1071, :813-890), evaluated synchronously.
//
// ONE body (tile_sweep), two staging policies (template parameter PIPE), two kernel symbols:
//   bp_sweep_pipe_kernel (PIPE = true)   two-stage asynchronous pipeline.  Per CTA (persistent), in iteration i:
//     - the messages of tile i -- gathered in-messages AND old out-messages -- are already in shared memory: they were
//       fetched with cp.async (LDGSTS, no staging registers) during iteration i-1;
//     - the index arrays of tile i+1 are in shared memory as well, so the first thing iteration i does is to put tile
//       i+1's message fetches in flight;
//     - then it computes tile i entirely out of shared memory (contract in place, node combine, leave-one-out).
//   bp_sweep_fast_kernel (PIPE = false)  register-staged: a tile's messages are gathered into registers at the top of
//     its own iteration (less shared memory: the instantiations whose message ring does not fit, and the big-node leg
//     of the wide path).
// In both, every CONTIGUOUS stream of a tile -- rev / pos / info and the row_ptr slice -- is moved by the TMA unit:
// one cp.async.bulk per stream, issued by thread 0 one iteration ahead, completed on an mbarrier (SASS: UBLKCP,
// SYNCS).  Only the random message gather stays per-thread.  The per-thread field partials and max-diff accumulate in
// registers over all of a CTA's tiles and are reduced once, in a fixed order (bitwise reproducible), at the end.
// Tile descriptors come through a small ring in shared memory, fetched four tiles ahead (see tile_at below).
// Measured history of the body, with the variants that were slower: profiles/tile_kernel_rework_r02.md.  The multi-GPU
// PIPELINE path does not use this body yet (sweep_pipe_dist.cuh says why); the multi-GPU register-staged path does.
#pragma once
#include "bp_device.cuh"
#include "sweep_kernel.cuh"

#ifndef SBMBP_TILE_BULK
#define SBMBP_TILE_BULK 1  // 1: cp.async.bulk + mbarrier staging of the contiguous streams; 0: per-thread cp.async (comparison build)
#endif
#ifndef SBMBP_PIPE_MINB
#define SBMBP_PIPE_MINB 2  // resident CTAs per SM the pipeline kernel is compiled for
#endif
#ifndef SBMBP_TILE_LPN
#define SBMBP_TILE_LPN 1   // lanes per node in phase 2a (1, 2 or 4; more than one: the product over a node's slots is split and joined by shuffles)
#endif

namespace sbmbp {

constexpr unsigned kInfLarge = 0x80000000u;  // info word: bit 31 = node updates in the log domain (degree >= 50)
constexpr unsigned kPosBits = 29;  // multi-GPU plan (host side, mirror pull): owner rank << 29 | position on the owner
constexpr unsigned kPosMask = (1u << kPosBits) - 1u;

template <typename T, int QT>
__device__ __forceinline__ void ld_vec(MsgVec<T, QT> &m, const T *__restrict__ p) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
        const uint4 *s = reinterpret_cast<const uint4 *>(p);
        uint4 *d = reinterpret_cast<uint4 *>(m.v);
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i) d[i] = __ldg(s + i);
    } else {
        static_assert(bytes == 8, "Q x sizeof(T) must be 8 or a multiple of 16");
        *reinterpret_cast<uint2 *>(m.v) = __ldg(reinterpret_cast<const uint2 *>(p));
    }
}

template <typename T, int QT>
__device__ __forceinline__ void st_vec(const MsgVec<T, QT> &m, T *__restrict__ p) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
        uint4 *d = reinterpret_cast<uint4 *>(p);
        const uint4 *s = reinterpret_cast<const uint4 *>(m.v);
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i) d[i] = s[i];
    } else {
        *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(m.v);
    }
}

// cp.async helpers (LDGSTS): global -> shared without staging registers
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <typename T, int QT>
__device__ __forceinline__ void cp_async_vec(T *smem_dst, const T *gsrc) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i)
            cp_async16(reinterpret_cast<char *>(smem_dst) + 16 * i, reinterpret_cast<const char *>(gsrc) + 16 * i);
    } else {
        static_assert(bytes == 8, "Q x sizeof(T) must be 8 or a multiple of 16");
        cp_async8(smem_dst, gsrc);
    }
}

// ---- mbarrier / TMA bulk-copy primitives (SASS: SYNCS, UBLKCP)
__device__ __forceinline__ unsigned tile_smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tile_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tile_smem_u32(bar)), "r"(bytes) : "memory");
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
        "}\n" ::"r"(tile_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion (bytes) on an mbarrier of this CTA; 16-byte aligned on both sides, size a multiple of 16
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tile_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tile_smem_u32(bar))
                 : "memory");
}

static_assert(sizeof(Tile) == 24, "tile descriptors travel as three 8-byte words");
constexpr unsigned kTileRing = 8;  // descriptor ring (shared memory): the descriptor of tile i+4 is fetched in iteration i

// Dynamic shared memory of the tile sweep.  Staged index arrays and row offsets carry the slack the 16-byte alignment
// of the bulk copies needs: a copy starts at the aligned address below its first element and ends at the aligned
// address above its last (engine.cu pads the device arrays accordingly).
template <typename T, int QT, bool PIPE>
struct TileLay {
    using Cfg = TileCfg<T, QT>;
    static constexpr int NI = PIPE ? 2 : 1;                                                   // ring slots of the index arrays
    static constexpr int IW = Cfg::TE + 8;                                                    // words per staged index array
    static constexpr int RW = Cfg::TN + 8;                                                    // staged row offsets per slot
    static constexpr size_t msg_bytes = sizeof(T) * QT * Cfg::TE;                             // one tile of messages
    static constexpr size_t off_num = 0;                                                      // double[QT*TN]
    static constexpr size_t off_red = off_num + sizeof(double) * QT * Cfg::TN;                // double[8*(QT+2)]
    static constexpr size_t off_par = off_red + sizeof(double) * (kThreads / 32) * (QT + 2);  // double[5*QT]
    static constexpr size_t off_k = off_par + sizeof(double) * 5 * QT;                        // KernT[QT*QT]
    static constexpr size_t off_bar = (off_k + sizeof(KernT<T, QT>) * QT * QT + 15) & ~size_t(15);  // u64[4]: mbarriers
    static constexpr size_t off_tile = off_bar + 4 * sizeof(unsigned long long);              // u64[kTileRing][3]: tile descriptors
    static constexpr size_t off_row = off_tile + kTileRing * 3 * sizeof(unsigned long long);  // u64[2][RW]
    static constexpr size_t off_idx = off_row + 2 * sizeof(unsigned long long) * RW;          // u32[NI][3][IW]
    static constexpr size_t off_msg = (off_idx + NI * 3 * sizeof(unsigned) * IW + 127) & ~size_t(127);  // whole shared-memory rows per warp access
    // PIPE: T[2][2][QT*TE], ring slot x (in-messages, contracted in place into b_e | old out-messages), entry-major;
    // else: T[QT*TE], b_e component-major
    static constexpr size_t bytes = off_msg + (PIPE ? 4 : 1) * msg_bytes;
    static_assert(off_row % 16 == 0 && off_idx % 16 == 0 && off_msg % 16 == 0, "bulk-copy destinations are 16-byte aligned");
};

// What the per-edge code needs of the current tile in shared memory.
template <typename T, int QT, bool PIPE>
struct TileView {
    const T *sb;             // b_e of the tile
    const double *snum;      // per node: normalised total (product domain) or log total - max (log domain); [q * TN + n]
    const double *par;       // eta | logeta | h | exph
    const unsigned *row32;   // low words of the staged row offsets, already shifted to the tile's first node
    unsigned e0lo;
    unsigned dc;
    double Nd;
    __device__ __forceinline__ T b(unsigned k, int q) const {
        return PIPE ? sb[size_t(k) * QT + q] : sb[size_t(q) * TileCfg<T, QT>::TE + k];
    }
    __device__ __forceinline__ unsigned off(unsigned n) const { return row32[2 * n] - e0lo; }  // tile-local first slot of node n
};

// Rare per-entry paths of phase 3, out of line: the node updates in the log domain (degree >= 50,
// belief_propagation.cpp:859), or some b_e[q] vanishes and the leave-one-out product is taken directly (the reference's
// own fallback there, :1029-1042, is not a function of the inputs -- DESIGN.md section 6 -- so the event is counted).
template <typename T, int QT, bool PIPE>
__device__ __noinline__ MsgVec<T, QT> tile_cavity_rare(const TileView<T, QT, PIPE> v, unsigned k, unsigned n, bool large,
                                                       unsigned long long *tiny_count) {
    constexpr int TN = TileCfg<T, QT>::TN;
    MsgVec<T, QT> cav;
    if (large) {
        double w[QT], mx = -1.0e300;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) {
            w[q] = v.snum[q * TN + n] - log(double(v.b(k, q)));
            mx = fmax(mx, w[q]);
        }
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) cav.v[q] = T(exp(w[q] - mx));
    } else {
        atomicAdd(tiny_count, 1ull);
        const unsigned k0 = v.off(n), d = v.off(n + 1) - k0;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) {
            double p = 1.0;
            for (unsigned kk = k0; kk < k0 + d; ++kk)
                if (kk != k) p *= double(v.b(kk, q));
            const double F = v.dc ? exp(-1.0 * double(d) * v.par[2 * QT + q] / v.Nd) : v.par[3 * QT + q];
            cav.v[q] = T(p * v.par[q] * F);
        }
    }
    return cav;
}

// DIST = false: single GPU, old values read from S_old[pos], new values written to S_new[pos].
// DIST = true : the message buffers belong to the DESTINATION's rank.  A pos word with bit 31 set is an index into this
//   rank's outbox (old value read there, new value written there); every CTA works on whole super-tiles and ships each
//   one's part of the outbox to the owners when it is complete (dist_exchange.cuh).
template <typename T, int QT, bool DIST, bool PIPE>
__device__ __forceinline__ void tile_sweep(const SweepArgs<T> &a) {
    using Cfg = TileCfg<T, QT>;
    using Lay = TileLay<T, QT, PIPE>;
    using View = TileView<T, QT, PIPE>;
    constexpr int TE = Cfg::TE, TN = Cfg::TN, IW = Lay::IW, RW = Lay::RW;
    constexpr int EPT = TE / kThreads;
    constexpr unsigned Q = QT;
    extern __shared__ __align__(16) unsigned char smem[];
    double *snum = reinterpret_cast<double *>(smem + Lay::off_num);
    double *sred = reinterpret_cast<double *>(smem + Lay::off_red);
    double *seta = reinterpret_cast<double *>(smem + Lay::off_par);
    double *slogeta = seta + QT;
    double *sh = seta + 2 * QT;
    double *sexph = seta + 3 * QT;
    KernT<T, QT> *sK = reinterpret_cast<KernT<T, QT> *>(smem + Lay::off_k);
    unsigned long long *bar_row = reinterpret_cast<unsigned long long *>(smem + Lay::off_bar);  // [2]
    unsigned long long *bar_idx = bar_row + 2;                                                  // [NI]
    unsigned long long *stile = reinterpret_cast<unsigned long long *>(smem + Lay::off_tile);
    unsigned long long *srow = reinterpret_cast<unsigned long long *>(smem + Lay::off_row);
    unsigned *sidx = reinterpret_cast<unsigned *>(smem + Lay::off_idx);
    T *smsg = reinterpret_cast<T *>(smem + Lay::off_msg);

    Ctl *ctl = a.ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // multi-GPU, inside a batch: the previous sweep was left open -- wait for every rank's flag, close it from the rows in
    // the sync block (dist_exchange.cuh); the sweep count is the host's (no round trip through the control block)
    bool lazy = false;
    if constexpr (DIST) lazy = a.dx.from_rows != 0;
    const unsigned sweeps_done = lazy ? a.dx.seq : ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;  // uniform over the grid
    if constexpr (DIST) {
        if (lazy) {
            __shared__ double s_open[QT + 1];
            SweepArgsBase ob;
            ob.prm = a.prm;
            ob.field[0] = a.field[0];
            ob.field[1] = a.field[1];
            ob.ctl = a.ctl;
            ob.partial = a.partial;
            if (dist_open_sweep<QT>(ob, a.dx, sweeps_done, s_open, sh, sexph, blockIdx.x == 0)) return;  // converged: uniform
        }
    }
    const int par = int(sweeps_done & 1u);
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const Field *fld = par ? a.field[1] : a.field[0];
    const bool dc = a.dc != 0;
    const double Nd = a.prm->N;
    const T damp = T(a.damping), keep = T(1.0 - a.damping);

    for (int i = tid; i < QT * QT; i += kThreads) sK[i] = KernT<T, QT>(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) {
        seta[tid] = a.prm->eta[tid];
        slogeta[tid] = a.prm->logeta[tid];
        if (!lazy) {
            sh[tid] = fld->h[tid];
            sexph[tid] = fld->exph[tid];
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(bar_row + i, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- staging.  A regular tile (<= TE edges) has its index arrays and row offsets staged; a hub tile reads its own.
    auto regular = [&](const Tile &t) { return t.ne <= unsigned(TE); };
    auto has_idx = [&](const Tile &t) { return t.ne <= unsigned(TE) && t.ne != 0u; };
    constexpr bool kBulk = SBMBP_TILE_BULK != 0;
    // thread 0: the three index arrays of a tile -> slot s, one bulk copy each
    auto issue_idx = [&](const Tile &t, int s) {
        if (!has_idx(t)) return;
        const unsigned shift = unsigned(t.e0) & 3u;
        const unsigned bytes = ((shift + t.ne + 3u) & ~3u) * unsigned(sizeof(unsigned));
        const unsigned long long base = t.e0 - shift;
        unsigned *dst = sidx + size_t(s) * 3 * IW;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot's previous readers are behind a barrier
        mbar_expect_tx(bar_idx + s, 3u * bytes);
        bulk_load(dst, a.rev + base, bytes, bar_idx + s);
        bulk_load(dst + IW, a.pos + base, bytes, bar_idx + s);
        bulk_load(dst + 2 * IW, a.info + base, bytes, bar_idx + s);
    };
    // thread 0: row_ptr[n0 .. n0 + nn] of a tile -> slot s
    auto issue_row = [&](const Tile &t, int s) {
        if (!regular(t)) return;
        const unsigned shift = t.n0 & 1u;
        const unsigned bytes = ((shift + t.nn + 2u) & ~1u) * unsigned(sizeof(unsigned long long));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar_row + s, bytes);
        bulk_load(srow + size_t(s) * RW, a.row_ptr + (t.n0 - shift), bytes, bar_row + s);
    };
    // Staging of the index arrays of tile ti (slot si, if do_idx) and the row offsets of tile tr (slot sr, if do_row).
    // Comparison build: every thread copies the entries it will read itself (indices) / its share (rows, published by
    // a barrier).  (Issuing from lane 0 of four otherwise idle warps during phase 2, one stream each, was measured:
    // slower -- the copies then have half an iteration to land.)
    auto stage = [&](const Tile &ti, int si, bool do_idx, const Tile &tr, int sr, bool do_row) {
        if constexpr (kBulk) {
            if (tid == 0) {
                if (do_idx) issue_idx(ti, si);
                if (do_row) issue_row(tr, sr);
            }
        } else {
            if (do_idx && has_idx(ti)) {
                unsigned *dst = sidx + size_t(si) * 3 * IW + (unsigned(ti.e0) & 3u);
#pragma unroll
                for (int u = 0; u < EPT; ++u) {
                    const unsigned k = u * kThreads + tid;
                    if (k < ti.ne) {
                        cp_async4(dst + k, a.rev + ti.e0 + k);
                        cp_async4(dst + IW + k, a.pos + ti.e0 + k);
                        cp_async4(dst + 2 * IW + k, a.info + ti.e0 + k);
                    }
                }
            }
            if (do_row && regular(tr)) {
                for (unsigned n = tid; n <= tr.nn; n += kThreads)
                    cp_async8(srow + size_t(sr) * RW + (tr.n0 & 1u) + n, a.row_ptr + tr.n0 + n);
            }
            cp_async_commit();
        }
    };
    unsigned phase = 0u;  // bits 0-1: parity the next wait on bar_row[s] expects; bits 2-3: bar_idx[s]
    auto wait_row = [&](int s) {
        if constexpr (kBulk) {
            mbar_wait(bar_row + s, (phase >> s) & 1u);
            phase ^= 1u << s;
        }
    };
    auto wait_idx = [&](int s) {
        if constexpr (kBulk) {
            mbar_wait(bar_idx + s, (phase >> (2 + s)) & 1u);
            phase ^= 4u << s;
        } else {
            cp_async_wait_all();  // this thread's own copies
        }
    };
    // all threads: the staged indices of a tile (slot s) -> registers; own / info of dead entries are 0
    auto read_idx = [&](const Tile &t, int s, unsigned (&gat)[EPT], unsigned (&own)[EPT], unsigned (&inf)[EPT]) {
        const unsigned *src = sidx + size_t(s) * 3 * IW + (unsigned(t.e0) & 3u);
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const unsigned k = u * kThreads + tid;
            const bool live = k < t.ne;
            gat[u] = live ? src[k] : 0u;
            own[u] = live ? src[IW + k] : 0u;
            inf[u] = live ? src[2 * IW + k] : 0u;
        }
    };
    auto old_src = [&](unsigned own) -> const T * {
        bool remote = DIST && (own & kRemoteBit);
#ifdef SBMBP_TUNING
        if (DIST && (a.dx.dbg & 8u) && remote) return Sold;  // timing only: the outbox is not touched
#endif
        return remote ? a.mirror + size_t(own & ~kRemoteBit) * Q : Sold + size_t(own) * Q;
    };
    // PIPE, all threads: a tile's messages -> ring slot s: in-messages by gather index (slot order), old out-messages by
    // own position (buffer order; the outbox for remote ones).  Every cp.async destination is read back by the thread
    // that issued it.
    auto fetch_msgs = [&](const Tile &t, int s, unsigned (&own)[EPT], unsigned (&inf)[EPT]) {
        if (regular(t)) {
            unsigned gat[EPT];
            read_idx(t, s, gat, own, inf);
            T *min = smsg + size_t(s) * 2 * QT * TE;
            T *mold = min + QT * TE;
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < t.ne) {
                    cp_async_vec<T, QT>(min + size_t(k) * QT, Sold + size_t(gat[u]) * Q);
                    cp_async_vec<T, QT>(mold + size_t(k) * QT, old_src(own[u]));
                }
            }
        }
        cp_async_commit();
    };

    // multi-GPU shipping state (dist_exchange.cuh)
    __shared__ T *s_peer[kMaxRanks];
    if constexpr (DIST) {
        if (tid < kMaxRanks) s_peer[tid] = (par ? a.peer[0] : a.peer[1])[tid];
    }
    // The j-th tile of this CTA.  Single GPU: blockIdx + j * gridDim (strided: the CTAs of a wave work on neighbouring
    // tiles, i.e. in the same destination bucket).  Multi-GPU: whole super-tiles of tps (a power of two) consecutive
    // tiles, strided by super-tile, so that a CTA ships what it computed; a wave still spans only gridDim * tps tiles.
    const unsigned G = gridDim.x;
    const unsigned tps_shift = DIST ? unsigned(31 - __clz(int(a.dx.tps))) : 0u;
    const unsigned tps_mask = (1u << tps_shift) - 1u;
    auto nth = [&](unsigned j) -> unsigned {
        const unsigned long long t = (((unsigned long long)(j >> tps_shift) * G + blockIdx.x) << tps_shift) + (j & tps_mask);
        return t < a.ntiles ? unsigned(t) : 0xffffffffu;
    };
    unsigned jt = 0;
    unsigned tile_id = nth(0);
    if (tile_id == 0xffffffffu) return;
    // Tile descriptors come through a ring in shared memory: thread 0 fetches the descriptor of this CTA's tile i+4 with
    // cp.async during iteration i, its wait at the top of iteration i+1 and that iteration's barriers publish it, and
    // everybody reads it at the end of iteration i+1 or later.  (Loaded straight into registers the descriptor -- uniform
    // over the CTA -- is moved into uniform registers right behind the load, which stalls every warp for the full
    // memory latency once per tile: 16 % of all stall samples in the first capture of this kernel.)
    auto tile_at = [&](unsigned j) -> Tile {
        const unsigned long long *p = stile + (j & (kTileRing - 1u)) * 3;
        const unsigned long long w1 = p[1], w2 = p[2];
        Tile t;
        t.e0 = p[0];
        t.n0 = unsigned(w1);
        t.nn = unsigned(w1 >> 32);
        t.ne = unsigned(w2);
        t.nbig = unsigned(w2 >> 32);
        return t;
    };
    auto fetch_tile = [&](unsigned j) {  // thread 0; joins the caller's next cp.async group
        const unsigned id = nth(j);
        if (id != 0xffffffffu) {
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.tiles + id);
            unsigned long long *dst = stile + (j & (kTileRing - 1u)) * 3;
            cp_async8(dst, src);
            cp_async8(dst + 1, src + 1);
            cp_async8(dst + 2, src + 2);
        }
    };
    if (tid < 4) {
        const unsigned id = nth(unsigned(tid));
        if (id != 0xffffffffu) {
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.tiles + id);
#pragma unroll
            for (int w = 0; w < 3; ++w) stile[tid * 3 + w] = __ldg(src + w);
        }
    }
    unsigned own[EPT], inf[EPT], own_n[EPT], inf_n[EPT];
    __syncthreads();  // parameters, mbarriers and the first descriptors are in shared memory
    Tile t0 = tile_at(0);                                      // tile i
    Tile t1 = (nth(1) != 0xffffffffu) ? tile_at(1) : t0;       // tile i+1
    Tile t2 = (PIPE && nth(2) != 0xffffffffu) ? tile_at(2) : t0;  // tile i+2 (PIPE only)
    stage(t0, 0, true, t0, 0, true);
    if (PIPE && nth(1) != 0xffffffffu) stage(t1, 1, true, t1, 1, false);
    if constexpr (PIPE) {
        if (has_idx(t0)) wait_idx(0);
        fetch_msgs(t0, 0, own, inf);
        __syncthreads();  // slot 0 of the index ring has been read: iteration 0 refills it
    }
    double wsum[QT];  // this thread's share of sum_i w_i psi_i^t and of the max-diff, over all the CTA's tiles
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;
    int ring = 0;  // slot of tile i (rows; PIPE: indices and messages too); tile i+1 uses ring ^ 1

    for (; tile_id != 0xffffffffu; tile_id = nth(++jt), ring ^= 1) {
        const Tile tile = t0;
        const unsigned long long e0 = tile.e0;
        const unsigned n0 = tile.n0, nn = tile.nn, ne = tile.ne;
        const bool have1 = nth(jt + 1) != 0xffffffffu;
        const bool have2 = PIPE && nth(jt + 2) != 0xffffffffu;
        T *sb = PIPE ? smsg + size_t(ring) * 2 * QT * TE : smsg;  // b_e of tile i
        const T *sold = sb + QT * TE;                              // PIPE: old out-messages of tile i
        MsgVec<T, QT> m[PIPE ? 1 : EPT], oldv[PIPE ? 1 : EPT];     // register staging (PIPE: unused)

        // everything this thread issued one iteration ago has landed: (PIPE) its messages of tile i, (thread 0) a descriptor
        cp_async_wait_all();
        if (tid == 0) fetch_tile(jt + 4);
        if constexpr (PIPE) {
            // put the next tile's messages in flight (ring ^ 1: its previous user, tile i-1, finished before the barrier
            // that ended the last iteration), then the indices of the tile after next into the slot tile i's came from
            if (have1) {
                if (has_idx(t1)) wait_idx(ring ^ 1);
                fetch_msgs(t1, ring ^ 1, own_n, inf_n);
            }
            stage(t2, ring, have2, t1, ring ^ 1, have1);
        } else {
            cp_async_commit();  // the descriptor fetch
            if (regular(tile)) {
                unsigned gat[EPT];
                if (has_idx(tile)) wait_idx(0);
                read_idx(tile, 0, gat, own, inf);
                // gather (slot order); the old values of phase 3 (buffer order) ride along
#pragma unroll
                for (int u = 0; u < EPT; ++u)
                    if (u * kThreads + tid < ne) ld_vec<T, QT>(m[u], Sold + size_t(gat[u]) * Q);
#pragma unroll
                for (int u = 0; u < EPT; ++u)
                    if (u * kThreads + tid < ne) ld_vec<T, QT>(oldv[u], old_src(own[u]));
                __syncthreads();  // the staged indices have been read; the previous tile is done with sb
            }
            stage(t1, 0, have1, t1, ring ^ 1, have1);
        }

        if (regular(tile)) {
            // =============================================================== regular tile
            // ---- phase 1: contract (PIPE: in place, each thread reads and rewrites only its own slots)
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < ne) {
                    T b[QT];
                    if constexpr (PIPE) {
                        MsgVec<T, QT> mk;
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) mk.v[q] = sb[size_t(k) * QT + q];
                        contract<T, QT>(mk, sK, b);
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) sb[size_t(k) * QT + q] = b[q];
                    } else {
                        contract<T, QT>(m[u], sK, b);
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) sb[q * TE + k] = b[q];
                    }
                }
            }
            wait_row(ring);  // row offsets of tile i (issued one iteration ago)
            View v;
            v.sb = sb;
            v.snum = snum;
            v.par = seta;
            v.row32 = reinterpret_cast<const unsigned *>(srow + size_t(ring) * RW + (n0 & 1u));
            v.e0lo = unsigned(e0);
            v.dc = a.dc;
            v.Nd = Nd;
            __syncthreads();

            // ---- phase 2a: one thread per node of degree < 32 (product domain).  (LPN > 1: the lanes of a node take every
            // LPN-th slot and join their partial products with shuffles -- shorter chains, more warps busy; measured slower,
            // 1.95 / 2.00 ms with 2 / 4 lanes against 1.86 ms on the configs[3] shard, so the default is 1.)  Uniform trip
            // count: the shuffles are warp-wide.
            constexpr int LPN = SBMBP_TILE_LPN;
            for (unsigned base = 0; base < nn; base += kThreads / LPN) {
                const unsigned n = base + unsigned(tid) / LPN, hl = unsigned(tid) % LPN;
                unsigned k0 = 0u, d = 32u;
                if (n < nn) {
                    k0 = v.off(n);
                    d = v.off(n + 1) - k0;
                }
                const bool mine = d < 32;  // a node of this tile, degree < 32
                double tot[QT];
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) tot[q] = 1.0;
                if (mine) {
                    for (unsigned k = k0 + hl; k < k0 + d; k += LPN) {
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) tot[q] *= double(v.b(k, q));
                    }
                }
                if constexpr (LPN > 1) {
#pragma unroll
                    for (int o = LPN / 2; o > 0; o >>= 1) {
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) tot[q] *= __shfl_xor_sync(0xffffffffu, tot[q], o);
                    }
                }
                if (!mine) continue;
                double sum = 0.0;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const double F = dc ? exp(-1.0 * double(d) * sh[q] / Nd) : sexph[q];
                    tot[q] = tot[q] * seta[q] * F;
                    sum += tot[q];
                }
                const double w = dc ? double(d) : 1.0;
                const double rsum = fast_rcp(sum);
                MsgVec<double, QT> mg;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) mg.v[q] = tot[q] * rsum;
                if (hl == 0) {
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        snum[q * TN + n] = mg.v[q];
                        wsum[q] += w * mg.v[q];
                    }
                    st_vec<double, QT>(mg, a.marg + size_t(n0 + n) * Q);
                }
            }
            // ---- phase 2b: one warp per node of degree >= 32 (product below 50, log domain from 50 on)
            for (unsigned n = warp; n < (tile.nbig ? nn : 0u); n += kThreads / 32) {
                const unsigned k0 = v.off(n), d = v.off(n + 1) - k0;
                if (d < 32) continue;
                const bool logdom = d >= kLargeDegree;
                double acc[QT];
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) acc[q] = logdom ? 0.0 : 1.0;
                for (unsigned k = k0 + lane; k < k0 + d; k += 32) {
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        const double bv = double(v.b(k, q));
                        if (logdom) acc[q] += log(bv);
                        else acc[q] *= bv;
                    }
                }
                double mx = -1.0e300, sum = 0.0;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    if (logdom) {
                        acc[q] = warp_sum(acc[q]) + slogeta[q] - (dc ? 1.0 * double(d) * sh[q] / Nd : sh[q] / Nd);
                        mx = fmax(mx, acc[q]);
                    } else {
                        const double F = dc ? exp(-1.0 * double(d) * sh[q] / Nd) : sexph[q];
                        acc[q] = warp_prod(acc[q]) * seta[q] * F;
                        sum += acc[q];
                    }
                }
                MsgVec<double, QT> mg;
                if (logdom) {
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        mg.v[q] = exp(acc[q] - mx);
                        sum += mg.v[q];
                    }
                }
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = (logdom ? mg.v[q] : acc[q]) / sum;
                    if (lane == 0) snum[q * TN + n] = logdom ? acc[q] - mx : mg.v[q];
                }
                if (lane == 0) {
                    const double w = dc ? double(d) : 1.0;
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) wsum[q] += w * mg.v[q];
                    st_vec<double, QT>(mg, a.marg + size_t(n0 + n) * Q);
                }
            }
            __syncthreads();

            // ---- phase 3 (buffer order): leave-one-out, normalise, max-diff, damped write
            // One pass per entry.  (Three passes over a thread's entries -- straight-line cavities for all, rare paths, then
            // normalise + store -- were measured 17 % slower; hoisting only the loads of all entries in front, no change.)
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned t = u * kThreads + tid;
                if (t >= ne) continue;
                const unsigned k = inf[u] & 0xffffu, n = (inf[u] >> 16) & 0x7fffu;
                T b[QT], cav[QT], old[QT];
                bool tiny = false;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    b[q] = v.b(k, q);
                    if constexpr (PIPE) old[q] = sold[size_t(t) * QT + q];
                    else old[q] = oldv[u].v[q];
                    tiny = tiny || !(double(b[q]) >= kEps);
                }
                if (!tiny && !(inf[u] & kInfLarge)) {
                    if constexpr (QT <= 4) {
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) {
                            T c = T(snum[q * TN + n]);
SBMBP_UNROLL_Q
                            for (int r = 0; r < QT; ++r)
                                if (r != q) c *= b[r];
                            cav[q] = c;
                        }
                    } else {
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) cav[q] = T(snum[q * TN + n]) / b[q];
                    }
                } else {
                    const MsgVec<T, QT> r = tile_cavity_rare<T, QT, PIPE>(v, k, n, (inf[u] & kInfLarge) != 0u, &a.ctl->tiny_count);
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) cav[q] = r.v[q];
                }
                T s = cav[0];
SBMBP_UNROLL_Q
                for (int q = 1; q < QT; ++q) s += cav[q];
                const T inv = fast_rcp(s);
                if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;  // non-finite message: make it visible
                MsgVec<T, QT> out;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const T nv = cav[q] * inv;
                    mydiff = fmax(mydiff, fabs(double(old[q]) - double(nv)));
                    out.v[q] = damp * nv + keep * old[q];
                }
                if (DIST && (own[u] & kRemoteBit)) st_vec<T, QT>(out, a.mirror + size_t(own[u] & ~kRemoteBit) * Q);  // outbox
                else st_vec<T, QT>(out, Snew + size_t(own[u]) * Q);
            }
        } else {
            // =============================================================== hub node (degree > TE): log domain
            const double dd = double(ne);
            double acc[QT];
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) acc[q] = 0.0;
            for (unsigned k = tid; k < ne; k += kThreads) {
                MsgVec<T, QT> mk;
                ld_vec<T, QT>(mk, Sold + size_t(__ldg(a.rev + e0 + k)) * Q);
                T b[QT];
                contract<T, QT>(mk, sK, b);
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) acc[q] += log(double(b[q]));
            }
            double mx = -1.0e300;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                acc[q] = block_sum(acc[q], sred) + slogeta[q] - (dc ? 1.0 * dd * sh[q] / Nd : sh[q] / Nd);
                mx = fmax(mx, acc[q]);
            }
            double sum = 0.0;
            MsgVec<double, QT> mg;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                mg.v[q] = exp(acc[q] - mx);
                sum += mg.v[q];
            }
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) mg.v[q] /= sum;
            if (tid == 0) {
                const double w = dc ? dd : 1.0;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) wsum[q] += w * mg.v[q];
                st_vec<double, QT>(mg, a.marg + size_t(n0) * Q);
            }
            for (unsigned k = tid; k < ne; k += kThreads) {
                MsgVec<T, QT> mk, oldk;
                const unsigned o = __ldg(a.pos + e0 + k);  // hub tiles keep slot order
                ld_vec<T, QT>(mk, Sold + size_t(__ldg(a.rev + e0 + k)) * Q);
                const bool remote = DIST && (o & kRemoteBit);
                ld_vec<T, QT>(oldk, remote ? a.mirror + size_t(o & ~kRemoteBit) * Q : Sold + size_t(o) * Q);
                T b[QT];
                contract<T, QT>(mk, sK, b);
                double w[QT], wmx = -1.0e300;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    w[q] = (acc[q] - mx) - log(double(b[q]));
                    wmx = fmax(wmx, w[q]);
                }
                T cav[QT], s = T(0);
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    cav[q] = T(exp(w[q] - wmx));
                    s += cav[q];
                }
                const T inv = fast_rcp(s);
                MsgVec<T, QT> out;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const T nv = cav[q] * inv;
                    mydiff = fmax(mydiff, fabs(double(oldk.v[q]) - double(nv)));
                    out.v[q] = damp * nv + keep * oldk.v[q];
                }
                if (remote) st_vec<T, QT>(out, a.mirror + size_t(o & ~kRemoteBit) * Q);  // outbox
                else st_vec<T, QT>(out, Snew + size_t(o) * Q);
            }
        }

        // ---- the tile is closed: nobody reads its b_e / old values / row offsets / snum after this barrier, so the ring
        // slots can be refilled; in multi-GPU mode every thread's outbox stores precede it
        __syncthreads();
        if constexpr (DIST) {
            if (((jt + 1) & tps_mask) == 0u || !have1) {  // last tile of one of this CTA's super-tiles: carry it to the owners
                const unsigned sp = tile_id >> tps_shift;
                bool tma = false;
                if constexpr (PIPE && (QT * sizeof(T)) % 16 == 0) tma = a.dx.ship_tma != 0;
                if constexpr (PIPE && (QT * sizeof(T)) % 16 == 0) {
                    if (tma)
                        dist_ship_supertile_tma<T, QT, kThreads>(a.dx, sp, a.mirror, s_peer, reinterpret_cast<unsigned char *>(sb), unsigned(2 * Lay::msg_bytes));
                }
                if (!tma) dist_ship_range<T, QT, kThreads>(a.dx, a.dx.out_start[sp], a.dx.out_start[sp + 1], a.mirror, s_peer);
            }
        }
        // rotate the pipeline registers
        t0 = t1;
        if constexpr (PIPE) {
            t1 = t2;
            if (nth(jt + 3) != 0xffffffffu) t2 = tile_at(jt + 3);
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                own[u] = own_n[u];
                inf[u] = inf_n[u];
            }
        } else {
            if (nth(jt + 2) != 0xffffffffu) t1 = tile_at(jt + 2);
        }
    }
    cp_async_wait_all();
    if constexpr (DIST) {
        dist_ship_drain();  // what this CTA shipped has landed at its owners
    }
    // ---- this CTA's row: field partials and max-diff over its tiles, reduced in a fixed order (bitwise reproducible)
    mydiff = warp_max(mydiff);
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = warp_sum(wsum[q]);
    if (lane == 0) {
        sred[warp * (QT + 1) + QT] = mydiff;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) sred[warp * (QT + 1) + q] = wsum[q];
    }
    __syncthreads();
    if (tid <= QT) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) r = (tid < QT) ? r + sred[w * (QT + 1) + tid] : fmax(r, sred[w * (QT + 1) + tid]);
        a.partial[size_t(blockIdx.x) * (QT + 1) + tid] = r;  // one row per CTA
    }
    if (a.fused_close) {
        SweepArgsBase base;
        base.prm = a.prm;
        base.field[0] = a.field[0];
        base.field[1] = a.field[1];
        base.ctl = a.ctl;
        base.partial = a.partial;
        if constexpr (DIST) close_sweep_dist<QT>(base, a.dx, gridDim.x, sweeps_done);
        else close_sweep_last_cta<QT>(base, gridDim.x, sweeps_done, a.row_out);
    }
}

template <typename T, int QT, bool DIST>
__global__ void __launch_bounds__(kThreads, SBMBP_PIPE_MINB) bp_sweep_pipe_kernel(const SweepArgs<T> a) {
    tile_sweep<T, QT, DIST, true>(a);
}

template <typename T, int QT, bool DIST>
__global__ void __launch_bounds__(kThreads, SBMBP_MINB) bp_sweep_fast_kernel(const SweepArgs<T> a) {
    tile_sweep<T, QT, DIST, false>(a);
}

}  // namespace sbmbp
