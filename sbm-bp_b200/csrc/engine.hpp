// Engine state shared by engine.cu (C ABI, drivers, small kernels) and the per-Q instantiation units (inst.cu).
#pragma once
#include <cstdint>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sbmbp.h"
#include "bp_device.cuh"
#include "dist_exchange.cuh"
#include "graph.hpp"

using namespace sbmbp;

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _err = (expr);                                                                       \
        if (_err != cudaSuccess) {                                                                       \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_err) + " (" + __FILE__ + ":" +     \
                      std::to_string(__LINE__) + ")");                                                   \
            return SBMBP_ERR_CUDA;                                                                       \
        }                                                                                                \
    } while (0)

#define TRY(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != SBMBP_OK) return _rc; \
    } while (0)

// engines may live on several devices of one process: per-device caches of function attributes / occupancy are
// indexed by sbmbp_engine::device (sbmbp_create refuses higher indices)
constexpr int kMaxDevices = 64;

struct sbmbp_engine {
    const sbmbp_graph *g = nullptr;
    uint32_t N = 0, Q = 0, dc = 0;
    uint64_t M = 0;
    int prec = SBMBP_F64, qt = 2, device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    uint32_t exact_pairs_max_n = 1u << 17;  // non-edge terms: exact O(N^2) pair sum up to this N, moment series beyond
    unsigned nbuckets = 1;  // destination buckets of the message layout (see build_layout in engine.cu)
    bool fast_path = false;  // bp_sweep_fast_kernel applies (Q == qt, dc != 2, one kernel matrix)
    bool pipe_path = true;   // bp_sweep_pipe_kernel (cp.async two-stage pipeline) where its shared memory fits
    bool warp_path = false;  // bp_sweep_warp_kernel (warp tiles, QT <= 4): built at create time, used when the fast path applies
    bool time_kernel = false;  // bracket the sweep kernel alone with ev0/ev1 (sbmbp_time_sweep_kernel)
    int gather_mode = 0;  // ld_gather16 flavour (SBMBP_GATHER_MODE while tuning)

    // device
    unsigned long long *d_row_ptr = nullptr;
    unsigned *d_rev = nullptr, *d_pos = nullptr, *d_info = nullptr, *d_col = nullptr, *d_degsrc = nullptr, *d_true = nullptr;
    void *d_S[2] = {nullptr, nullptr};
    double *d_marg = nullptr;
    Tile *d_tiles = nullptr;
    unsigned ntiles = 0;
    // warp-tile view of the same graph (sweep_warp.cuh): descriptors, out positions sorted per warp tile, slot/node words
    WTile *d_wtiles = nullptr;
    Tile *d_hubs = nullptr;
    unsigned nwtiles = 0, nhubs = 0;
    unsigned *d_wpos = nullptr;
    unsigned short *d_winfo = nullptr;
    // wide-Q path (sweep_wide.cuh, Q = 32): nodes of degree <= 32 by the warp-per-node kernel, the rest as a tile list of
    // their own through bp_sweep_fast_kernel
    bool wide_path = false;
    unsigned *d_wide_nodes = nullptr;
    unsigned n_wide_nodes = 0;
    Tile *d_btiles = nullptr;
    unsigned nbtiles = 0;
    unsigned *d_bpos = nullptr, *d_binfo = nullptr;
    // degree-class (ELL) layout of the message buffers (sweep_ell.cuh): chosen at create time when they fit the L2
    bool ell_path = false;
    // ... and its padded one-bucket variant (build_ell_padded_layout): no pos words, marginals land in a chunk-ordered array and are scattered back to
    // node order lazily (sync_marg), the message buffers carry the lane padding of the layout (buf_slots >= M)
    bool ell_padded = false;
    double *d_marg_ell = nullptr;
    unsigned ell_entries = 0;     // entries of d_marg_ell / d_ell_node (32 per chunk)
    bool marg_ell_dirty = false;   // d_marg_ell is newer than d_marg for the nodes of the degree classes
    uint64_t buf_slots = 1;        // message slots per buffer
    // compact storage (Q = 2, FP64, the whole graph on the degree-class kernel): one double per message in d_C, used by
    // the sweeps; everything else works on d_S, converted on demand (ensure_full / ensure_compact)
    bool compact_ok = false;       // the engine may use compact storage (creation-time eligibility, SBMBP_COMPACT != 0)
    bool compact_refused = false;  // the current state holds un-normalised messages: full storage until a new state arrives
    bool compact = false;          // the current state lives in d_C[sweeps_done & 1]
    void *d_C[2] = {nullptr, nullptr};
    EllClass *d_ell_cls = nullptr;
    unsigned ell_ncls = 0, ell_nchunks = 0;
    unsigned *d_ell_rev = nullptr, *d_ell_pos = nullptr, *d_ell_node = nullptr;
    unsigned long long *d_trace = nullptr;  // SBMBP_ELL_TRACE=1: per-warp globaltimer stamps of the last ELL sweep (tuning only)
    unsigned trace_warps = 0;
    uint4 *d_ell_sched = nullptr;  // per-warp work lists of the ELL kernel (build_ell_schedule)
    unsigned ell_sched_len = 0, ell_grid = 0, ell_wpc = 8;  // work-list length, CTAs, warps per CTA
    uint64_t ell_nidx = 0;  // words in ell_rev / ell_pos
    DevParams *d_prm = nullptr;
    Field *d_field[2] = {nullptr, nullptr};
    Ctl *d_ctl = nullptr;
    double *d_partial = nullptr;  // sweep: [ntiles][qt]; other reductions reuse d_scratch
    double *d_scratch = nullptr;
    size_t scratch_doubles = 0;
    double *d_out = nullptr;  // small result vector
    Ctl *h_ctl = nullptr;     // pinned
    double *h_out = nullptr;  // pinned

    // multi-GPU (sbmbp_create_dist): messages INTO this rank's nodes live here; out-messages are written into the
    // owner's buffer through CUDA IPC; `mirror` keeps this rank's out-messages for the max-diff / damping
    bool dist = false;
    int rank = 0, world = 1;
    uint32_t N_global = 0;
    void *d_mirror = nullptr;  // outbox of the remote out-messages (dist_exchange.cuh): old values + what is shipped
    void *peer[2][8] = {};
    unsigned *d_rpos = nullptr;  // per tile entry: owner << 29 | position at the owner (mirror pull)
    unsigned *d_out_start = nullptr, *d_out_rpos = nullptr;  // outbox range per super-tile; destination of every outbox entry
    ShipDesc *d_ship = nullptr;                              // runs contiguous at one owner, per super-tile (TMA shipping)
    unsigned *d_ship_start = nullptr;
    int ship_tma = 0;                                        // SBMBP_SHIP_TMA
    unsigned tps = 8, nsuper = 0;
    uint64_t n_remote = 0;
    SyncBlock *d_sync = nullptr;        // this rank's flags + rows, written by every rank
    void *sync_peer[8] = {};            // every rank's sync block (own: d_sync)
    bool dist_open = false;             // the last DIST sweep has not been closed yet (its rows sit in the sync block)
    unsigned dist_seq = 0;              // sweeps launched / completed as the host counts them (== sweeps_done when closed)
    std::vector<void *> ipc_opened;
    double *d_row = nullptr;  // [qt + 1] this rank's reduced row (sent to the all-gather)

    // init_messages flags 1-3 (belief_propagation.cpp:132-215): conf_planted_, and whether planted nodes are frozen
    // (bp_conditional, -m infer, main.cpp:322) or updated like any other (bp_basic, -m learn)
    int *d_clamp = nullptr;
    std::vector<int> conf_planted;
    uint32_t n_planted = 0;
    bool conditional = true;

    // update schedule: synchronous (default) or coloured asynchronous (sbmbp_set_schedule): a greedy colouring of the
    // graph; one sweep = one pass per colour through the general kernel, every pass seeing the previous one's messages
    int schedule = 0;
    unsigned char *d_color = nullptr;
    unsigned ncolors = 0, cur_color = 0;

    // reference-exact replay (SBMBP_SCHED_REPLAY, sweep_replay.cuh): the generator converge() draws its schedule from
    // (left by init_messages in the state the reference's engine has after its draws, or seeded explicitly), and the
    // reference-order index arrays / scratch of the one-warp kernel, allocated on first use
    std::mt19937 rng;
    bool rng_valid = false;
    unsigned *d_rp_rev = nullptr, *d_rp_degn = nullptr, *d_rp_sched = nullptr;
    double *d_rp_h = nullptr, *d_rp_scratch = nullptr;  // d_rp_h: [kMaxQ] field, then [1] the sweep's maxdiffm

    // host mirrors
    std::vector<uint32_t> na;
    std::vector<double> cab, eta;
    double beta = 1.0;
    bool have_params = false, have_state = false, field_valid = false;
    unsigned sweeps_done = 0;  // mirrors ctl->sweeps_done
    uint64_t stat_edge_updates = 0, stat_sweeps = 0, stat_launches = 0;
    double stat_seconds = 0.0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // cached edge-pass results, valid for (state_version, kernel choice)
    uint64_t state_version = 1, energy_version[2] = {0, 0};
    std::vector<double> energy_out[2];
};


// defined in inst.cu, one translation unit per QT (compiled in parallel)
template <typename T, int QT>
int launch_sweeps(sbmbp_engine *e, unsigned count, double damping);
template <typename T, int QT>
int launch_energy(sbmbp_engine *e, int which, std::vector<double> &out);
// multi-GPU: one DIST sweep kernel + the reduction of its rows into e->d_row (no finalisation)
template <typename T, int QT>
int launch_dist_sweep(sbmbp_engine *e, double damping);
template <int QT>
int launch_dist_close(sbmbp_engine *e);
// resident CTAs per SM of bp_sweep_ell_kernel<T, QT>, its unroll limit and its warps per CTA (0 / 0 where the kernel does not exist: QT > 4)
template <typename T, int QT>
int ell_kernel_config(int *ctas_per_sm, int *unroll_degree, int *warps_per_cta);
// d_marg_ell -> d_marg where the degree-class kernel left newer (chunk-ordered) marginals (no-op otherwise); call before anything reads d_marg
int sync_marg(sbmbp_engine *e);
// representation of the current message state: full (d_S, what every kernel but the compact sweep reads) or compact (d_C)
int ensure_full(sbmbp_engine *e);
int ensure_compact(sbmbp_engine *e, bool *now_compact);
int ensure_scratch(sbmbp_engine *e, size_t doubles);
// d_result[c] = sum over rows of d_partial[row][c], fixed order (defined in engine.cu)
int reduce_columns(sbmbp_engine *e, const double *d_partial, unsigned nrows, unsigned ncols, double *d_result);
