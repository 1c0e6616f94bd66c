// Per-Q instantiation unit of the tile kernels: compiled once per INST_QT (2, 4, 8, 16, 32), both precisions.
#include "energy_kernel.cuh"
#include <algorithm>
#include <cstring>

#include "engine.hpp"
#include "sweep_tile.cuh"
#include "sweep_pipe_dist.cuh"
#include "sweep_warp.cuh"
#include "sweep_ell.cuh"
#include "sweep_wide.cuh"
#include "state_kernels.cuh"
#include "sweep_kernel.cuh"

#ifndef INST_QT
#error "compile with -DINST_QT=<2|4|8|16|32>"
#endif

template <typename T>
static SweepArgs<T> make_args(sbmbp_engine *e, double damping);
static DistArgs make_dist_args(sbmbp_engine *e);

template <typename T, int QT>
int ell_kernel_config(int *ctas_per_sm, int *unroll_degree, int *warps_per_cta) {
    *ctas_per_sm = 0;
    *unroll_degree = 0;
    *warps_per_cta = kThreads / 32;
    if constexpr (QT <= 4) {
        CUDA_TRY(cudaFuncSetAttribute(bp_sweep_ell_kernel<T, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(EllSmem<T, QT>::bytes)));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, bp_sweep_ell_kernel<T, QT, false>, EllUnroll<T, QT>::NT, EllSmem<T, QT>::bytes));
        if (*ctas_per_sm < 1) *ctas_per_sm = 1;
        *unroll_degree = EllUnroll<T, QT>::DU;
        *warps_per_cta = EllUnroll<T, QT>::NT / 32;
    }
    return SBMBP_OK;
}

template <typename T, int QT>
static void dist_ship_publish_launch(sbmbp_engine *e, const SweepArgsBase &b, const SweepArgs<T> &g, unsigned blocks);

template <typename T, int QT>
int launch_dist_sweep(sbmbp_engine *e, double damping) {
    constexpr bool can_fast = (QT * sizeof(T)) % 16 == 0 || QT * sizeof(T) == 8;
    {
        // padded Q, deg_corr_flag 2, beta != 1 (two kernel matrices): the general kernel, then a kernel that ships the outbox
        // and publishes the rank's row.  The open sweep before it is closed by its own kernel so that the general kernel
        // finds the field and the control block current.
        SweepArgs<T> g = make_args<T>(e, damping);
        const bool general = !can_fast || e->Q != unsigned(QT) || e->dc == 2 || g.select_k;
        if (general) {
            static bool attr_by_device[kMaxDevices] = {};
            const size_t smem = TileSmem<T, QT>::bytes;
            if (!attr_by_device[e->device]) {
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_kernel<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                attr_by_device[e->device] = true;
            }
            if (!e->ntiles) {
                set_error("multi-GPU engine: every rank needs at least one node");
                return SBMBP_ERR_UNSUPPORTED;
            }
            SweepArgsBase b;
            b.prm = e->d_prm;
            b.field[0] = e->d_field[0];
            b.field[1] = e->d_field[1];
            b.ctl = e->d_ctl;
            b.partial = e->d_partial;
            if (e->dist_open) bp_dist_close_kernel<QT><<<1, kFinalThreads, 0, e->stream>>>(b, g.dx);
            g.fused_close = 0;
            bp_sweep_kernel<T, QT><<<e->ntiles, kThreads, smem, e->stream>>>(g);
            const unsigned blocks = std::max(1u, std::min<unsigned>(unsigned((e->n_remote * e->Q + kThreads - 1) / kThreads), 2u * unsigned(e->sm_count)));
            dist_ship_publish_launch<T, QT>(e, b, g, blocks);
            CUDA_TRY(cudaGetLastError());
            e->stat_launches += 2 + (e->dist_open ? 1 : 0);
            return SBMBP_OK;
        }
    }
    if constexpr (can_fast) {
        // pipeline policy: the previous-generation body (sweep_pipe_dist.cuh, see there); register-staged policy: the unified one
        constexpr bool can_pipe = PipeDistSmem<T, QT>::bytes <= 220 * 1024;
        const bool pipe = can_pipe && e->pipe_path;
        const size_t fast_smem = pipe ? PipeDistSmem<T, QT>::bytes : TileLay<T, QT, false>::bytes;
        static int ctas_by_device[kMaxDevices][2] = {};  // attributes and occupancy are per device
        int *ctas_per_sm = ctas_by_device[e->device];
        if (!ctas_per_sm[pipe]) {
            if (pipe) {
                if constexpr (can_pipe) {
                    CUDA_TRY(cudaFuncSetAttribute(bp_sweep_pipe_dist_kernel<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem)));
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[1], bp_sweep_pipe_dist_kernel<T, QT>, kThreads, fast_smem));
                }
            } else {
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_fast_kernel<T, QT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem)));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[0], bp_sweep_fast_kernel<T, QT, true>, kThreads, fast_smem));
            }
            if (ctas_per_sm[pipe] < 1) ctas_per_sm[pipe] = 1;
        }
        SweepArgs<T> a = make_args<T>(e, damping);
        a.fused_close = 1;
        a.row_out = nullptr;
        a.dx.from_rows = e->dist_open ? 1 : 0;  // the previous sweep of the batch is still open: close it in the prologue
        a.dx.seq = e->dist_seq;
        // every CTA owns whole super-tiles (and must report in at the close): at most one CTA per super-tile
        const unsigned grid = std::min<unsigned>(e->nsuper, unsigned(ctas_per_sm[pipe]) * unsigned(e->sm_count));
        if (e->ntiles) {
            if (pipe) {
                if constexpr (can_pipe) bp_sweep_pipe_dist_kernel<T, QT><<<grid, kThreads, fast_smem, e->stream>>>(a);
            } else {
                bp_sweep_fast_kernel<T, QT, true><<<grid, kThreads, fast_smem, e->stream>>>(a);
            }
        }
        if (!e->ntiles) {
            set_error("multi-GPU engine: every rank needs at least one node");
            return SBMBP_ERR_UNSUPPORTED;
        }
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
        return SBMBP_OK;
    }
    set_error("unreachable: every message width has a multi-GPU sweep");
    return SBMBP_ERR_UNSUPPORTED;
}

// The general-kernel sweep has already run on the stream; the destination buffer is the one of parity (sweeps_done & 1) ^ 1,
// and sweeps_done as the HOST counts it is exact here: the general path closes every open sweep first and is only taken
// outside device-side convergence batches of the persistent kernels (a converged batch re-synchronises the count at
// sbmbp_dist_close).
template <typename T, int QT>
static void dist_ship_publish_launch(sbmbp_engine *e, const SweepArgsBase &b, const SweepArgs<T> &g, unsigned blocks) {
    PeerPtrs<T> pp;
    const int par = int(e->dist_seq & 1u);
    for (int k = 0; k < kMaxRanks; ++k) pp.p[k] = par ? g.peer[0][k] : g.peer[1][k];
    dist_ship_publish_kernel<T, QT><<<blocks, kThreads, 0, e->stream>>>(b, g.dx, g.mirror, pp, unsigned(e->n_remote), e->Q, e->ntiles);
}

static DistArgs make_dist_args(sbmbp_engine *e) {
    DistArgs d;
    d.out_start = e->d_out_start;
    d.out_rpos = e->d_out_rpos;
    d.ship = e->d_ship;
    d.ship_start = e->d_ship_start;
    d.ship_tma = e->ship_tma;
    d.tps = e->tps;
    d.nsuper = e->nsuper;
    for (int k = 0; k < kMaxRanks; ++k) d.sync[k] = static_cast<SyncBlock *>(e->sync_peer[k]);
    d.rank = e->rank;
    d.world = e->world;
    d.from_rows = 0;
    d.seq = 0;
#ifdef SBMBP_TUNING
    d.dbg = 0;
    if (const char *env = std::getenv("SBMBP_DIST_DBG")) d.dbg = unsigned(std::atoi(env));
    d.trace = e->d_trace;
#endif
    return d;
}

template <int QT>
int launch_dist_close(sbmbp_engine *e) {
    SweepArgsBase b;
    b.prm = e->d_prm;
    b.field[0] = e->d_field[0];
    b.field[1] = e->d_field[1];
    b.ctl = e->d_ctl;
    b.partial = e->d_partial;
    bp_dist_close_kernel<QT><<<1, kFinalThreads, 0, e->stream>>>(b, make_dist_args(e));
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    return SBMBP_OK;
}

template <typename T>
static SweepArgs<T> make_args(sbmbp_engine *e, double damping) {
    SweepArgs<T> a;
    a.tiles = e->d_tiles;
    a.row_ptr = e->d_row_ptr;
    a.rev = e->d_rev;
    a.pos = e->d_pos;
    a.info = e->d_info;
    a.degsrc = e->d_degsrc;
    a.S[0] = static_cast<T *>(e->d_S[0]);
    a.S[1] = static_cast<T *>(e->d_S[1]);
    a.marg = e->d_marg;
    a.prm = e->d_prm;
    a.field[0] = e->d_field[0];
    a.field[1] = e->d_field[1];
    a.ctl = e->d_ctl;
    a.partial = e->d_partial;
    a.ntiles = e->ntiles;
    a.Q = e->Q;
    a.dc = e->dc;
    a.gmode = e->gather_mode;
    a.select_k = (e->dc == 0 && e->beta != 1.0) ? 1 : 0;
    a.clamp = (e->conditional && e->n_planted) ? e->d_clamp : nullptr;
    a.color = (e->schedule == SBMBP_SCHED_COLORED) ? e->d_color : nullptr;
    a.cur_color = e->cur_color;
    a.damping = damping;
    a.row_out = nullptr;
    a.fused_close = 0;
    a.mirror = static_cast<T *>(e->d_mirror);
    for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 8; ++k) a.peer[b][k] = static_cast<T *>(e->peer[b][k]);
    if (e->dist) a.dx = make_dist_args(e);
    else std::memset(&a.dx, 0, sizeof(a.dx));
    return a;
}

template <typename T, int QT>
int launch_sweeps(sbmbp_engine *e, unsigned count, double damping) {
    static bool attr_by_device[kMaxDevices] = {};  // cudaFuncSetAttribute is per device
    bool &attr_set = attr_by_device[e->device];
    const size_t smem = TileSmem<T, QT>::bytes;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(bp_sweep_kernel<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_set = true;
    }
    SweepArgs<T> a = make_args<T>(e, damping);
    constexpr bool can_fast = (QT * sizeof(T)) % 16 == 0 || QT * sizeof(T) == 8;
    // planted nodes under bp_conditional are frozen: only the general kernel knows how
    const bool fast = can_fast && e->fast_path && e->Q == unsigned(QT) && e->dc != 2 && !a.select_k && !a.clamp && !a.color;
    constexpr bool can_pipe = can_fast && TileLay<T, QT, true>::bytes <= 220 * 1024;
    const bool pipe = fast && can_pipe && e->pipe_path;
    const size_t fast_smem = pipe ? TileLay<T, QT, true>::bytes : TileLay<T, QT, false>::bytes;
    unsigned fast_grid = e->ntiles;
    if (fast) {
        static int ctas_by_device[kMaxDevices][2] = {};
        int *ctas_per_sm = ctas_by_device[e->device];
        if (!ctas_per_sm[pipe]) {
            if (pipe) {
                if constexpr (can_pipe) {
                    CUDA_TRY(cudaFuncSetAttribute(bp_sweep_pipe_kernel<T, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem)));
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[1], bp_sweep_pipe_kernel<T, QT, false>, kThreads, fast_smem));
                }
            } else {
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_fast_kernel<T, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem)));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[0], bp_sweep_fast_kernel<T, QT, false>, kThreads, fast_smem));
            }
            if (ctas_per_sm[pipe] < 1) ctas_per_sm[pipe] = 1;
        }
        // persistent: one resident wave of CTAs strides over the tiles
        fast_grid = std::min<unsigned>(e->ntiles, unsigned(ctas_per_sm[pipe]) * unsigned(e->sm_count));
    }
    a.fused_close = fast ? 1 : 0;
    if constexpr (QT <= 4) {
        if (fast && (e->ell_path || (e->warp_path && e->d_wtiles))) {
            // small-Q paths.  ELL: hubs (if any), warp tiles for degrees 32..WE (if any), then the degree-class kernel,
            // which closes the sweep.  Warp-main (SBMBP_WARP_MAIN=1): hubs, then the warp kernel over all other nodes.
            static int small_by_device[kMaxDevices][2] = {};
            int &warp_ctas_per_sm = small_by_device[e->device][0], &ell_ctas_per_sm = small_by_device[e->device][1];
            if (!warp_ctas_per_sm) {
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_warp_kernel<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WarpSmem<T, QT>::bytes)));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&warp_ctas_per_sm, bp_sweep_warp_kernel<T, QT>, kThreads, WarpSmem<T, QT>::bytes));
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_ell_kernel<T, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(EllSmem<T, QT>::bytes)));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ell_ctas_per_sm, bp_sweep_ell_kernel<T, QT, false>, EllUnroll<T, QT>::NT, EllSmem<T, QT>::bytes));
                if (warp_ctas_per_sm < 1) warp_ctas_per_sm = 1;
                if (ell_ctas_per_sm < 1) ell_ctas_per_sm = 1;
            }
            constexpr unsigned NW = kThreads / 32;
            const bool ell = e->ell_path;
            // compact message storage (one double per normalised Q = 2 message): when the whole sweep is this kernel
            constexpr bool can_compact = QT == 2 && sizeof(T) == 8;
            bool compact = false;
            if constexpr (can_compact) {
                if (ell && e->ell_padded && e->compact_ok && e->nwtiles == 0 && e->nhubs == 0) TRY(ensure_compact(e, &compact));
                static bool attr_by_device[kMaxDevices] = {};
                if (compact && !attr_by_device[e->device]) {
                    CUDA_TRY(cudaFuncSetAttribute(bp_sweep_ell_kernel<T, QT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(EllSmem<T, QT>::bytes)));
                    attr_by_device[e->device] = true;
                }
            }
            if (!compact) TRY(ensure_full(e));
            unsigned ell_rows = 0;
            EllSweepArgs<T> x;
            if (ell) {
                x.sched = e->d_ell_sched;
                x.sched_len = e->ell_sched_len;
                x.ell_rev = e->d_ell_rev;
                x.ell_pos = e->d_ell_pos;
                {
                    int pf = 0;  // measured on configs[1]: 64.7 us with the bulk prefetch, 63.3 us without -- off by default
                    if (const char *env = std::getenv("SBMBP_ELL_BULKPF")) pf = std::atoi(env);
                    const unsigned long long idx_bytes = (unsigned long long)e->ell_nidx * sizeof(unsigned);
                    x.pf_bytes[0] = pf ? (unsigned long long)e->M * QT * sizeof(T) : 0ull;
                    x.pf_bytes[1] = pf ? idx_bytes : 0ull;
                    x.pf_bytes[2] = pf ? idx_bytes : 0ull;
                }
#ifdef SBMBP_TUNING
                x.trace = e->d_trace;
                x.dbg = 0;
                if (const char *env = std::getenv("SBMBP_ELL_DEBUG")) x.dbg = unsigned(std::atoi(env));
#endif
                x.ell_node = e->d_ell_node;
                x.S[0] = static_cast<T *>(e->d_S[0]);
                x.S[1] = static_cast<T *>(e->d_S[1]);
                x.marg = e->d_marg;
                x.prm = e->d_prm;
                x.field[0] = e->d_field[0];
                x.field[1] = e->d_field[1];
                x.ctl = e->d_ctl;
                x.partial = e->d_partial;
                x.dc = e->dc;
                x.damping = damping;
                x.implicit_pos = e->ell_padded ? 1 : 0;
                if (e->ell_padded) x.marg = e->d_marg_ell;
                if (compact) {  // one double per message; the kernel indexes by message position
                    x.S[0] = static_cast<T *>(e->d_C[0]);
                    x.S[1] = static_cast<T *>(e->d_C[1]);
                }
                ell_rows = e->ell_grid;
            }
            WarpSweepArgs<T> w;
            w.tiles = e->d_wtiles;
            w.ntiles = e->nwtiles;
            w.hubs = e->d_hubs;
            w.nhubs = e->nhubs;
            w.row_ptr = e->d_row_ptr;
            w.rev = e->d_rev;
            w.pos = e->d_wpos;
            w.info = e->d_winfo;
            w.S[0] = static_cast<T *>(e->d_S[0]);
            w.S[1] = static_cast<T *>(e->d_S[1]);
            w.marg = e->d_marg;
            w.prm = e->d_prm;
            w.field[0] = e->d_field[0];
            w.field[1] = e->d_field[1];
            w.ctl = e->d_ctl;
            w.partial = e->d_partial;
            w.dc = e->dc;
            w.damping = damping;
            const unsigned want = (e->nwtiles + NW - 1) / NW;
            w.warp_rows = std::min<unsigned>(ell ? want : std::max(1u, want), unsigned(warp_ctas_per_sm) * unsigned(e->sm_count));
            w.hub_rows = std::min<unsigned>(e->nhubs, 4u * unsigned(e->sm_count));
            w.row_base = ell_rows;
            w.hub_row_base = ell_rows + w.warp_rows;
            w.close = ell ? 0 : 1;
            x.rows_before = w.warp_rows + w.hub_rows;
            if (!e->ell_padded) TRY(sync_marg(e));  // this sweep writes node-ordered marginals: bring the chunk-ordered ones home first
            unsigned launches = 0;
            // SBMBP_PDL (default 1): programmatic dependent launch, measured +2 % per step and +4 % on the converge loop.
            // SBMBP_LAZY_CLOSE (default 0): lazy sweep close, measured +0.4 .. 1 % on top of that (the tail of a sweep is
            // its stores draining, not the closing chain) -- kept as an option, exercised by the tests.
            int pdl = 1, lazy_env = 0;
            if (const char *env = std::getenv("SBMBP_PDL")) pdl = std::atoi(env);
            if (const char *env = std::getenv("SBMBP_LAZY_CLOSE")) lazy_env = std::atoi(env);
            // lazy close: only when the degree-class kernel carries the sweep alone and there is a batch to spread it over
            const bool lazy = ell && lazy_env && count > 1 && x.rows_before == 0 && !e->time_kernel;
            x.lazy = lazy ? 1 : 0;
            x.lazy_base = e->sweeps_done;
            x.lazy_k = 0;
            x.lazy_last = 1;
            for (unsigned s = 0; s < count; ++s) {
                if (e->time_kernel) CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
                if (w.hub_rows) bp_sweep_hub_kernel<T, QT><<<w.hub_rows, kThreads, 0, e->stream>>>(w);
                if (w.warp_rows) bp_sweep_warp_kernel<T, QT><<<w.warp_rows, kThreads, WarpSmem<T, QT>::bytes, e->stream>>>(w);
                if (ell) {
                    if (lazy) {
                        x.lazy_k = s;
                        x.lazy_last = (s + 1 == count) ? 1 : 0;
                    }
                    // programmatic dependent launch: this sweep's prologue overlaps the tail of whatever kernel precedes it
                    // on the stream (the previous sweep, the arm kernel); see the top of bp_sweep_ell_kernel
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(ell_rows);
                    cfg.blockDim = dim3(EllUnroll<T, QT>::NT);
                    cfg.dynamicSmemBytes = EllSmem<T, QT>::bytes;
                    cfg.stream = e->stream;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    attr[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = attr;
                    cfg.numAttrs = (pdl && x.rows_before == 0 && !e->time_kernel) ? 1 : 0;
                    if constexpr (can_compact) {
                        if (compact) CUDA_TRY(cudaLaunchKernelEx(&cfg, bp_sweep_ell_kernel<T, QT, true>, x));
                        else CUDA_TRY(cudaLaunchKernelEx(&cfg, bp_sweep_ell_kernel<T, QT, false>, x));
                    } else {
                        CUDA_TRY(cudaLaunchKernelEx(&cfg, bp_sweep_ell_kernel<T, QT, false>, x));
                    }
                }
                if (e->time_kernel) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
            }
            launches = (w.hub_rows ? 1u : 0u) + (w.warp_rows ? 1u : 0u) + (ell ? 1u : 0u);
            CUDA_TRY(cudaGetLastError());
            if (e->ell_padded && ell) e->marg_ell_dirty = true;
            e->stat_launches += uint64_t(launches) * count;
            return SBMBP_OK;
        }
    }
    if constexpr (QT == 32) {
        if (fast && e->wide_path) {
            // wide-Q path (sweep_wide.cuh): the few nodes of degree > 32 through the fast tile kernel (no close), then
            // one warp per node; the wide kernel's last CTA closes the sweep over both sets of rows
            static int wide_by_device[kMaxDevices][2] = {};
            int &wide_ctas_per_sm = wide_by_device[e->device][0], &big_ctas_per_sm = wide_by_device[e->device][1];
            const size_t wide_smem = WideSmem<T>::bytes, big_smem = TileLay<T, QT, false>::bytes;
            if (!wide_ctas_per_sm) {
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_wide_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(wide_smem)));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wide_ctas_per_sm, bp_sweep_wide_kernel<T>, kThreads, wide_smem));
                CUDA_TRY(cudaFuncSetAttribute(bp_sweep_fast_kernel<T, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(big_smem)));
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&big_ctas_per_sm, bp_sweep_fast_kernel<T, QT, false>, kThreads, big_smem));
                if (wide_ctas_per_sm < 1) wide_ctas_per_sm = 1;
                if (big_ctas_per_sm < 1) big_ctas_per_sm = 1;
            }
            TRY(ensure_full(e));
            SweepArgs<T> b = a;
            b.tiles = e->d_btiles;
            b.ntiles = e->nbtiles;
            b.pos = e->d_bpos;
            b.info = e->d_binfo;
            b.fused_close = 0;
            const unsigned big_grid = std::min<unsigned>(e->nbtiles, unsigned(big_ctas_per_sm) * unsigned(e->sm_count));
            WideSweepArgs<T> w;
            w.row_ptr = e->d_row_ptr;
            w.rev = e->d_rev;
            w.nodes = e->d_wide_nodes;
            w.nnodes = e->n_wide_nodes;
            w.S[0] = static_cast<T *>(e->d_S[0]);
            w.S[1] = static_cast<T *>(e->d_S[1]);
            w.marg = e->d_marg;
            w.prm = e->d_prm;
            w.field[0] = e->d_field[0];
            w.field[1] = e->d_field[1];
            w.ctl = e->d_ctl;
            w.partial = e->d_partial;
            w.rows_before = big_grid;
            w.dc = e->dc;
            w.damping = damping;
            constexpr unsigned NW = kThreads / 32;
            const unsigned wide_grid = std::max(1u, std::min<unsigned>((e->n_wide_nodes + NW - 1) / NW, unsigned(wide_ctas_per_sm) * unsigned(e->sm_count)));
            for (unsigned s = 0; s < count; ++s) {
                if (e->time_kernel) CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
                if (big_grid) bp_sweep_fast_kernel<T, QT, false><<<big_grid, kThreads, big_smem, e->stream>>>(b);
                bp_sweep_wide_kernel<T><<<wide_grid, kThreads, wide_smem, e->stream>>>(w);
                if (e->time_kernel) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
            }
            CUDA_TRY(cudaGetLastError());
            e->stat_launches += uint64_t(big_grid ? 2 : 1) * count;
            return SBMBP_OK;
        }
    }
    TRY(sync_marg(e));  // these kernels write node-ordered marginals
    TRY(ensure_full(e));
    for (unsigned s = 0; s < count; ++s) {
        if (e->time_kernel) CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
        if (pipe) {
            if constexpr (can_pipe) bp_sweep_pipe_kernel<T, QT, false><<<fast_grid, kThreads, fast_smem, e->stream>>>(a);
        } else if (fast) {
            bp_sweep_fast_kernel<T, QT, false><<<fast_grid, kThreads, fast_smem, e->stream>>>(a);
        } else {
            bp_sweep_kernel<T, QT><<<e->ntiles, kThreads, smem, e->stream>>>(a);
        }
        if (e->time_kernel) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
        if (!fast)
            bp_finalize_kernel<QT><<<1, kFinalThreads, 0, e->stream>>>(e->d_partial, e->ntiles, e->Q, e->d_prm,
                                                                  e->d_field[0], e->d_field[1], e->d_ctl);
    }
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += (fast ? 1 : 2) * count;
    return SBMBP_OK;
}

template <typename T, int QT>
int launch_energy(sbmbp_engine *e, int which, std::vector<double> &out) {
    static bool attr_by_device[kMaxDevices] = {};
    bool &attr_set = attr_by_device[e->device];
    const size_t smem = EnergySmem<T, QT>::bytes;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(bp_energy_kernel<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_set = true;
    }
    constexpr unsigned ncols = kEnergyHead + QT * QT;
    out.assign(ncols, 0.0);
    if (e->ntiles == 0) return SBMBP_OK;
    TRY(ensure_full(e));
    TRY(ensure_scratch(e, size_t(e->ntiles) * ncols + ncols));
    EnergyArgs<T> a;
    a.tiles = e->d_tiles;
    a.row_ptr = e->d_row_ptr;
    a.rev = e->d_rev;
    a.pos = e->d_pos;
    a.info = e->d_info;
    a.degsrc = e->d_degsrc;
    a.S = static_cast<const T *>(e->d_S[e->sweeps_done & 1u]);
    a.mirror = e->dist ? static_cast<const T *>(e->d_mirror) : nullptr;
    a.prm = e->d_prm;
    a.Kmat = (which == 0) ? e->d_prm->Ks : e->d_prm->C;
    a.field = e->d_field[e->sweeps_done & 1u];
    a.partial = e->d_scratch;
    a.Q = e->Q;
    a.dc = e->dc;
    a.fcoef = (which == 0) ? e->beta : 1.0;
    bp_energy_kernel<T, QT><<<e->ntiles, kThreads, smem, e->stream>>>(a);
    double *d_res = e->d_scratch + size_t(e->ntiles) * ncols;
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    TRY(reduce_columns(e, e->d_scratch, e->ntiles, ncols, d_res));
    CUDA_TRY(cudaMemcpyAsync(out.data(), d_res, ncols * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}


template int launch_sweeps<double, INST_QT>(sbmbp_engine *, unsigned, double);
template int launch_sweeps<float, INST_QT>(sbmbp_engine *, unsigned, double);
template int launch_energy<double, INST_QT>(sbmbp_engine *, int, std::vector<double> &);
template int launch_energy<float, INST_QT>(sbmbp_engine *, int, std::vector<double> &);
template int ell_kernel_config<double, INST_QT>(int *, int *, int *);
template int ell_kernel_config<float, INST_QT>(int *, int *, int *);
template int launch_dist_sweep<double, INST_QT>(sbmbp_engine *, double);
template int launch_dist_sweep<float, INST_QT>(sbmbp_engine *, double);
template int launch_dist_close<INST_QT>(sbmbp_engine *);
