/* sbmbp.h -- C ABI of libsbmbp.so, the B200 belief-propagation engine for (degree-corrected) SBM
 * inference and EM learning.
 *
 * The reference (junipertcy/sbm-bp) has no FFI: its boundary is the call sequence that
 * src/main.cpp:277-365 makes on graph_utilities / blockmodel / belief_propagation.  Each entry point
 * below names the reference interface it replaces.  Conventions: every function returns an int status
 * (0 = SBMBP_OK); no exception crosses the boundary; handles are opaque and NOT thread-safe; host
 * buffers are caller-owned and copied; device buffers are engine-owned.  Message state crosses the
 * boundary in the reference's own order: msg[(row_ptr[i]+l)*Q+q] == mmap_[i][l][q]
 * (belief_propagation.h:65-66), marg[i*Q+q] == real_psi_[i][q] (belief_propagation.h:45).
 * There is no CPU fallback: engine calls fail with SBMBP_ERR_NODEVICE when no sm_100 device is usable.
 */
#ifndef SBMBP_H
#define SBMBP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBMBP_OK 0
#define SBMBP_ERR_ARG 1         /* bad argument */
#define SBMBP_ERR_IO 2          /* file could not be opened */
#define SBMBP_ERR_PARSE 3       /* malformed edge-list line (the reference would invent an edge) */
#define SBMBP_ERR_RANGE 4       /* vertex id >= N (the reference indexes out of bounds) */
#define SBMBP_ERR_CUDA 5        /* a CUDA call failed; see sbmbp_last_error() */
#define SBMBP_ERR_NODEVICE 6    /* no usable CUDA device */
#define SBMBP_ERR_STATE 7       /* call order violated (e.g. sweep before set_params / state) */
#define SBMBP_ERR_UNSUPPORTED 8 /* e.g. Q > SBMBP_MAX_Q */

#define SBMBP_MAX_Q 32
#define SBMBP_F64 0 /* messages stored and combined in double */
#define SBMBP_F32 1 /* messages stored in float; node products, h, marginals and all reductions in double */

typedef struct sbmbp_graph sbmbp_graph;
typedef struct sbmbp_engine sbmbp_engine;

const char *sbmbp_version(void);
/* text of the last failure on the calling thread */
const char *sbmbp_last_error(void);

/* ---- graph: replaces load_edge_list (graph_utilities.cpp:42-58), edge_to_adj (:60-77), the
 * graph_neis_/graph_neis_inv_ flattening of bp_allocate (belief_propagation.cpp:246-266) and the degree
 * statistics of blockmodel_t (blockmodel.cpp:7-49).  Result: destination-sorted CSR, neighbours ascending,
 * duplicate edges merged, self-loops kept once.  N is sum(-n).  Blank lines are skipped; a malformed line
 * is an error (documented deviation); ids >= N are an error. */
int sbmbp_graph_from_edgelist(const char *path, uint32_t N, sbmbp_graph **g);
int sbmbp_graph_from_pairs(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N, sbmbp_graph **g);
int sbmbp_graph_destroy(sbmbp_graph *g);
/* N, M = directed edges (2E), E as blockmodel_t::get_E reports it, max degree (get_graph_max_degree) */
int sbmbp_graph_info(const sbmbp_graph *g, uint32_t *N, uint64_t *M, uint64_t *E, uint32_t *max_degree);
/* bit-exact views (owned by g): row_ptr[N+1]; col[M] == graph_neis_; rev[M] = global slot of the reverse
 * edge, i.e. row_ptr[col[e]] + graph_neis_inv_; deg[N].  Any pointer may be NULL. */
int sbmbp_graph_csr(const sbmbp_graph *g, const uint64_t **row_ptr, const uint32_t **col, const uint32_t **rev,
                    const uint32_t **deg);
/* the raw pair list exactly as the loader parsed it (for loader parity tests); returns count via n */
int sbmbp_parse_edgelist(const char *path, uint32_t *u, uint32_t *v, uint64_t cap, uint64_t *n);

/* Host-side view of the degree-class (ELL) message layout the small-Q sweep kernel uses (csrc/sweep_ell.cuh), for
 * tests; needs no GPU.  region_slots: in-slots per destination bucket (0 = one bucket).  Outputs (any may be NULL):
 * pos[M] / gather[M] = buffer position of the message out of / into each slot; classes: 5 words per class
 * (degree, nodes, node_first, chunk_first, index base), at most cls_cap classes written; node[n_node]; the index
 * arrays rev_idx / pos_idx (at most idx_cap words written, count in n_idx). */
int sbmbp_ell_layout(const sbmbp_graph *g, uint64_t region_slots, uint32_t *pos, uint32_t *gather, uint32_t *classes,
                     uint32_t cls_cap, uint32_t *n_cls, uint32_t *node, uint32_t *n_node, uint32_t *rev_idx,
                     uint32_t *pos_idx, uint64_t idx_cap, uint64_t *n_idx, uint32_t *n_chunks, uint32_t *n_buckets);

/* ---- parameters: bp_param_from_direct (blockmodel.cpp:274-302), bp_param_from_epsilon_c (:229-272).
 * na[Q], cab[Q*Q] row-major out.  cab_upper is the --cab vector (upper triangle, row-major). */
int sbmbp_params_from_direct(uint32_t N, uint32_t Q, const double *pa, const double *cab_upper, uint32_t *na,
                             double *cab);
int sbmbp_params_from_epsilon_c(uint32_t N, uint32_t Q, double epsilon, double c, uint32_t *na, double *cab);

/* ---- engine: replaces class belief_propagation (belief_propagation.h:18-178) for one graph on one GPU.
 * device < 0 selects the current device. */
int sbmbp_create(const sbmbp_graph *g, uint32_t Q, uint32_t deg_corr_flag, int precision, int device,
                 sbmbp_engine **e);
int sbmbp_destroy(sbmbp_engine *e);
/* run on this cudaStream_t (default: the legacy default stream) so callers can time with their own events */
int sbmbp_set_stream(sbmbp_engine *e, void *cuda_stream);
/* expand_bp_params + set_beta (belief_propagation.cpp:290-317, :417-419) */
int sbmbp_set_params(sbmbp_engine *e, const uint32_t *na, const double *cab, double beta);
int sbmbp_get_params(sbmbp_engine *e, uint32_t *na, double *cab, double *eta);
/* init_messages flag 0 (belief_propagation.cpp:110-131): identical draws to std::mt19937(seed) */
int sbmbp_init_random(sbmbp_engine *e, uint32_t seed);
/* init_messages, flags 0-3 (belief_propagation.cpp:101-215): conf[N] = the beliefs vector of main.cpp:325-336
 * (--beliefs_path / -f), -1 = unknown; ignored for flag 0.  The reference's quirks are kept (see engine.cu); where
 * its assert(conf != 1) would abort (flags 2, 3) the call fails with SBMBP_ERR_UNSUPPORTED. */
int sbmbp_init_messages(sbmbp_engine *e, uint32_t flag, const int32_t *conf, uint32_t seed);
/* bp_conditional (on, default; -m infer, main.cpp:322): nodes with a belief != -1 and degree < 50 are frozen
 * (belief_propagation.cpp:1100-1126); bp_basic (off; -m learn): they are updated like any other node */
int sbmbp_set_conditional(sbmbp_engine *e, int on);
/* Update schedule.  SBMBP_SCHED_SYNC (default): every node from the previous sweep's messages.  SBMBP_SCHED_COLORED:
 * graph-coloured asynchronous sweeps -- the nodes of one colour (no two adjacent) are updated together, colour after
 * colour, each pass seeing the messages and the field h the previous pass left (the reference updates one random node
 * at a time, belief_propagation.cpp:392-405; this is the parallel form of that).  Same fixed point, fewer sweeps, each
 * sweep costs one pass over the graph per colour; runs on the general kernel. */
#define SBMBP_SCHED_SYNC 0
#define SBMBP_SCHED_COLORED 1
/* SBMBP_SCHED_REPLAY: the reference's own schedule, draw for draw -- N draws with replacement per sweep,
 * i = unsigned(int(U * N)) from std::mt19937 (belief_propagation.cpp:394-395), one node updated in place per draw with h
 * maintained incrementally (:1088-1095), arithmetic in the reference's order.  Serial by construction (one warp,
 * ~2 us per draw): for laying niter and trajectories next to the reference's on graphs up to ~1e5 nodes, not for
 * throughput.  The generator is the one sbmbp_init_messages / sbmbp_init_random left behind (as main.cpp passes one
 * engine to init_messages and inference), or sbmbp_seed_schedule. */
#define SBMBP_SCHED_REPLAY 2
int sbmbp_set_schedule(sbmbp_engine *e, int schedule);
/* The engine's generator: main.cpp:236 seeds ONE std::mt19937 and hands it to blockmodel_t::shuffle (--mb_rand),
 * init_messages and inference / learning in turn.  sbmbp_seed_schedule: std::mt19937(seed).  sbmbp_rng_shuffle: the
 * draws of std::shuffle over n memberships (blockmodel.cpp:103-106).  sbmbp_init_messages_continue: init_messages
 * drawing from the generator where it stands (sbmbp_init_messages(.., seed) == seed + continue); the replay schedule
 * goes on from there. */
int sbmbp_seed_schedule(sbmbp_engine *e, uint32_t seed);
int sbmbp_rng_shuffle(sbmbp_engine *e, uint32_t n);
int sbmbp_init_messages_continue(sbmbp_engine *e, uint32_t flag, const int32_t *conf);
/* greedy colouring used by SBMBP_SCHED_COLORED (host only, for tests): color[N], returns the number of colours */
int sbmbp_graph_coloring(const sbmbp_graph *g, uint8_t *color, uint32_t *n_colors);
/* same distribution from a counter-based generator on the device (for graphs too large to seed serially) */
int sbmbp_init_random_device(sbmbp_engine *e, uint64_t seed);
/* host state in reference order; either pointer may be NULL.  h is derived (init_h, :320-332). */
int sbmbp_set_state(sbmbp_engine *e, const double *msg, const double *marg);
int sbmbp_get_state(sbmbp_engine *e, double *msg, double *marg, double *h);
/* marginals only (what inference() prints with --if_output_marginals, :94) */
int sbmbp_get_marginals(sbmbp_engine *e, double *marg);

/* one synchronous sweep = M directed-edge updates (the body of converge(), :392-405, all nodes at once
 * from the previous sweep's messages); maxdiff as norm_m_at_i defines it (:1059-1063) */
int sbmbp_sweep(sbmbp_engine *e, double damping, double *maxdiff);
/* n sweeps back to back without host synchronisation or convergence test (throughput measurement) */
int sbmbp_sweeps_async(sbmbp_engine *e, uint32_t n, double damping);
int sbmbp_sync(sbmbp_engine *e);
/* one sweep; kernel_ms = device time of the sweep kernel alone (events around that single launch) */
int sbmbp_time_sweep_kernel(sbmbp_engine *e, double damping, float *kernel_ms);
/* converge() (:386-415): sweeps until maxdiff < crit (float compare as :406); niter = sweep index or -1 */
int sbmbp_converge(sbmbp_engine *e, float crit, uint32_t max_sweeps, float damping, int *niter);
/* compute_free_energy (:744-750) = -f_site + f_edge + f_non_edge; parts may be NULL */
int sbmbp_free_energy(sbmbp_engine *e, double *f, double *f_site, double *f_edge, double *f_non_edge);
/* compute_f_non_edge / compute_entropy_non_edge (:675-741) are O(N^2) in the reference.  Up to exact_pairs_max_n nodes
 * (default 2^17; SBMBP_EXACT_PAIRS_MAX_N in the environment at create time) the engine sums the pairs exactly on the
 * device; beyond it evaluates the moment series  sum_k <W1^(x)k, T_k (x) T_k> / k,  T_k = sum_i psi_i^(x)k, minus the exact
 * sum over the edges.  This setter lets tests pin the series at golden sizes (0 = always the series). */
int sbmbp_set_exact_pairs_max_n(sbmbp_engine *e, uint32_t n);
/* The series' host arithmetic for callers that hold all-reduced moment tensors (multi-GPU, sbm-bp_b200/dist.py); no
 * device needed.  cab[Q*Q] row-major; order K as the engine chooses it (remainder / 2N below 1e-14, Q^K <= 2^20);
 * term = -<W1^(x)k, T (x) T> / k with W1 = 1 - pow(1 - c/N, beta) (the reference's rounded weight) and T the order-k moment tensor (first digit fastest);
 * k = 0: T[0] = sum_i log(sum_q psi_i^q), term = 2 N T[0] -- what the marginals' rounding-level normalisation defect adds
 * to the pair sum (systematic per node, so O(N ulp) in total); the series runs over k = 0 .. K. */
int sbmbp_non_edge_series_order(uint32_t Q, double N, double beta, const double *cab, uint32_t *K);
int sbmbp_non_edge_series_term(uint32_t Q, double N, double beta, const double *cab, uint32_t k, const double *T,
                               double *term);
/* compute_entropy (:752-758), the first field of the infer stdout line */
int sbmbp_entropy(sbmbp_engine *e, double *entropy);
/* compute_overlap (:775-811); true_conf[N] */
int sbmbp_overlap(sbmbp_engine *e, const uint32_t *true_conf, double *overlap);
/* compute_na_expect + compute_cab_expect (:428-440, :892-989) */
int sbmbp_em_stats(sbmbp_engine *e, double *na_expect, double *nna_expect, double *cab_expect);
/* learning() (:14-51) with learning_step (:53-75); outputs the learned na, cab, eta; em_iters optional */
int sbmbp_learn(sbmbp_engine *e, float crit, uint32_t max_time, float learning_rate, float damping,
                uint32_t *na_out, double *cab_out, double *eta_out, int *em_iters);

/* tuning aid: with SBMBP_ELL_TRACE=1 in the environment at create time the degree-class sweep kernel leaves 16
 * globaltimer stamps per warp (entry, start of work, after each of its first 12 chunks, end of work, exit) */
int sbmbp_debug_trace(sbmbp_engine *e, uint64_t *out, uint64_t cap_words, uint64_t *n_words);

/* name of the sweep kernel the engine launches for its current graph / parameters (reporting only) */
int sbmbp_sweep_kernel_name(sbmbp_engine *e, char *buf, uint32_t cap);

/* counters since creation: directed-edge updates, sweeps, kernel launches, algorithmic bytes per edge update
 * (SURVEY.md 8d), device seconds spent in sweeps as measured by events around sbmbp_converge */
int sbmbp_stats(sbmbp_engine *e, uint64_t *edge_updates, uint64_t *sweeps, uint64_t *launches,
                double *bytes_per_edge, double *sweep_seconds);

/* Edge updates since creation in which some b_l[q] = sum_t K_tq psi_t fell below EPS = 1e-50 (belief_propagation.h:69).
 * There the reference drops eta_q x field from that component (belief_propagation.cpp:1029-1042) and, for b == 0, reads
 * scratch left by an earlier update (:1013-1016): its result is not a function of the inputs.  The engine evaluates
 * the exact leave-one-out product instead, so parity with the reference is NOT claimed for runs where this counter
 * is non-zero (bin/bp says so on stderr).  Such states need exact zeros: init flags 1/3 with a zero c_ab entry. */
int sbmbp_tiny_events(sbmbp_engine *e, uint64_t *n);

/* ---- multi-GPU: one process per GPU, node-range partition.  The reference has nothing to mirror
 * here.  Rank p owns the nodes [range_starts[p], range_starts[p+1]), their in-slots, marginals and the buffers
 * holding every message INTO them.  The sweep kernel collects its remote out-messages in a local outbox ordered so
 * that what a super-tile of source nodes sends to one owner is contiguous at both ends, and ships each completed
 * super-tile to the owners' buffers with TMA bulk copies over the CUDA-IPC mappings (NVLink) while the other CTAs keep
 * computing; ranks synchronise on the device (per-rank flags + rows of Q+1 doubles in IPC-mapped sync blocks), so a
 * batch of sweeps needs neither the host nor NCCL (csrc/dist_exchange.cuh).  Only init_h, the plan exchange and the
 * reductions of the free energy / EM statistics go through the caller's collectives (sbm-bp_b200/dist.py).
 * Supported: Q in {2,4,8,16,32}, deg_corr_flag 0/1 (sweeps, free energy, EM), beta = 1, at most 8 ranks, < 2^29 in-edges
 * and < 2^31 remote out-edges per rank, at least one node per rank. */
typedef struct sbmbp_plan sbmbp_plan;
/* rows of the nodes [lo, hi) of an N_global-node graph; col holds global ids */
int sbmbp_graph_from_pairs_range(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N_global,
                                 uint32_t lo, uint32_t hi, sbmbp_graph **g);
/* layout of this rank's buffer + the positions it must tell each producer rank (sendlist, in (source node,
 * destination node) order); feed what the peers sent with plan_recv (peer == rank included), then plan_finish */
int sbmbp_plan_create(const sbmbp_graph *g, uint32_t Q, int precision, int rank, int world,
                      const uint32_t *range_starts, sbmbp_plan **p);
int sbmbp_plan_sendlist(sbmbp_plan *p, int peer, const uint32_t **data, uint64_t *n);
int sbmbp_plan_expect(sbmbp_plan *p, int peer, uint64_t *n);
int sbmbp_plan_recv(sbmbp_plan *p, int peer, const uint32_t *data, uint64_t n);
int sbmbp_plan_finish(sbmbp_plan *p);
/* host views for tests: gather[M] (where in-slot e's message sits), pos[M] (tile-sorted kernel words: bit 31 set =
 * index into the outbox of remote out-messages, else position in this rank's own buffer), info[M] (tile-local slot |
 * node << 16 | log-domain flag << 31), pos_slot[M] (owner << 29 | position at the owner, in slot order) */
int sbmbp_plan_layout(sbmbp_plan *p, const uint32_t **gather, const uint32_t **pos, const uint32_t **info,
                      const uint32_t **pos_slot, uint64_t *M, uint32_t *ntiles);
/* halo-exchange tables (tests, accounting): tile-sorted owner << 29 | position (rpos[M]); tiles per super-tile;
 * outbox range per super-tile (out_start[nsuper + 1]); shipping descriptors grouped by super-tile (ship_start[nsuper + 1];
 * ship[4 * n_ship] = src outbox index, dst position, length, owner).  Any pointer may be NULL. */
int sbmbp_plan_exchange_tables(sbmbp_plan *p, const uint32_t **rpos, uint32_t *tiles_per_super, uint32_t *nsuper,
                               const uint32_t **out_start, const uint32_t **ship_start, const uint32_t **ship,
                               uint64_t *n_ship, uint64_t *n_remote);
int sbmbp_plan_destroy(sbmbp_plan *p);
int sbmbp_create_dist(sbmbp_plan *p, uint32_t deg_corr_flag, int device, sbmbp_engine **e);
/* handles: 192 bytes = the cudaIpcMemHandle_t of the two message buffers and of the sync block */
int sbmbp_dist_ipc_export(sbmbp_engine *e, void *handles);
int sbmbp_dist_ipc_import(sbmbp_engine *e, int peer, const void *handles);
/* after every rank holds a state and a barrier: fetch this rank's out-messages from their owners */
int sbmbp_dist_sync_mirror(sbmbp_engine *e);
/* this rank's row of init_h / of one sweep: device pointer to ncols doubles [field partials (Q) .. max-diff] */
int sbmbp_dist_field_local(sbmbp_engine *e, void **row_dev, uint32_t *ncols);
int sbmbp_dist_arm(sbmbp_engine *e, float crit, uint32_t max_sweeps);
/* n sweeps back to back with no host in between (device-side barrier between them); the last one stays open */
int sbmbp_dist_sweeps(sbmbp_engine *e, uint32_t n, double damping);
/* closes the open sweep on the device (all ranks' rows -> field, max-diff, convergence); sync != 0 reads the result back */
int sbmbp_dist_close(sbmbp_engine *e, int sync, double *maxdiff, int *converged, int *niter);
/* init_h over the ranks: gathered_dev = device pointer to world x ncols doubles, rank-major (the all-gathered
 * sbmbp_dist_field_local rows); advance must be 0 */
int sbmbp_dist_finalize(sbmbp_engine *e, const void *gathered_dev, int advance, int sync, double *maxdiff,
                        int *converged, int *niter);
/* deg_global[N_global]: the degrees of all nodes (the ranks' degree arrays, concatenated by the caller); needed before
 * sbmbp_dist_energy_local when deg_corr_flag != 0 (d_i d_l per edge, l possibly remote) */
int sbmbp_dist_set_degrees(sbmbp_engine *e, const uint32_t *deg_global);
/* this rank's share of the edge pass (free energy, entropy, EM two-point sums), of the moment tensors and of the
 * edge correction of the non-edge term; the caller all-reduces (sbm-bp_b200/dist.py). */
int sbmbp_dist_energy_local(sbmbp_engine *e, int which, double *row, uint32_t cap, uint32_t *ncols);
int sbmbp_dist_moment_local(sbmbp_engine *e, uint32_t order, double *T, uint64_t cap);
int sbmbp_dist_edge_pairs_local(sbmbp_engine *e, const void *marg_global_dev, int mode, double *result);
/* local node sums for overlap / EM expectations (to be all-reduced); row has ncols doubles */
int sbmbp_dist_node_stats(sbmbp_engine *e, const uint32_t *true_conf_local, double *row, uint32_t *ncols);

#ifdef __cplusplus
}
#endif
#endif
