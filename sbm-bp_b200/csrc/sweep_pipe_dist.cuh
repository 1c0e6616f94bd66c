// Multi-GPU instantiation of the two-stage pipeline sweep: the PREVIOUS generation of the kernel body, kept for the
// DIST path only.  The unified body (sweep_tile.cuh: bulk-copy staging, descriptor ring, per-thread accumulators, 12 %
// faster on one GPU) measured SLOWER as a DIST instantiation -- 4.0 ms against 3.2 ms per step on 2 GPUs x 12.5M nodes,
// 4.4 ms with per-thread cp.async staging -- although the DIST-specific code (outbox, super-tile order, shipping, lazy
// close) is the same in both; the cause is not identified yet (no multi-GPU profiler pass was possible), so the
// multi-GPU engine stays on this body until it is.  Same arithmetic, same tiles, same results as sweep_tile.cuh.
//
// Per CTA (persistent, whole super-tiles of tps tiles), in iteration i:
//   * the messages of tile i -- gathered in-messages AND old out-messages -- are already in shared memory: they were
//     fetched with cp.async (LDGSTS, no staging registers) during iteration i-1;
//   * the index arrays / row offsets of tile i+1 are in shared memory as well (per-thread cp.async), so the first thing
//     iteration i does is to put tile i+1's message fetches in flight, and the index fetches of tile i+2;
//   * then it computes tile i entirely out of shared memory (contract in place, node combine, leave-one-out);
//   * after the last tile of a super-tile the CTA ships that super-tile's part of the outbox to the owners.
#pragma once
#include "bp_device.cuh"
#include "sweep_tile.cuh"

namespace sbmbp {

template <typename T, int QT>
struct PipeDistSmem {
    using Cfg = TileCfg<T, QT>;
    static constexpr size_t msg_bytes = sizeof(T) * QT * Cfg::TE;                            // one tile of messages
    static constexpr size_t off_num = 0;                                                      // double[QT*TN]
    static constexpr size_t off_red = off_num + sizeof(double) * QT * Cfg::TN;                // double[8*(QT+2)]
    static constexpr size_t off_par = off_red + sizeof(double) * (kThreads / 32) * (QT + 2);  // double[5*QT]
    static constexpr size_t off_k = off_par + sizeof(double) * 5 * QT;                        // T[QT*QT]
    static constexpr size_t off_off = (off_k + sizeof(KernT<T, QT>) * QT * QT + 15) & ~size_t(15);       // u32[TN+4]
    static constexpr size_t off_row = (off_off + sizeof(unsigned) * (Cfg::TN + 4) + 15) & ~size_t(15);  // u64[2][TN+8]
    static constexpr size_t off_idx = off_row + 2 * sizeof(unsigned long long) * (Cfg::TN + 8);         // u32[2][3][TE]
    static constexpr size_t off_msg = (off_idx + 2 * 3 * sizeof(unsigned) * Cfg::TE + 15) & ~size_t(15);  // T[2][2][QT*TE]: in | old
    static constexpr size_t bytes = off_msg + 4 * msg_bytes;
};

template <typename T, int QT>
__global__ void __launch_bounds__(kThreads, 2) bp_sweep_pipe_dist_kernel(const SweepArgs<T> a) {
    constexpr bool DIST = true;
    using Cfg = TileCfg<T, QT>;
    using Lay = PipeDistSmem<T, QT>;
    constexpr int TE = Cfg::TE, TN = Cfg::TN;
    constexpr int EPT = TE / kThreads;
    constexpr int NPT = (TN + 1 + kThreads - 1) / kThreads;
    constexpr unsigned Q = QT;
    extern __shared__ __align__(16) unsigned char smem[];
    double *snum = reinterpret_cast<double *>(smem + Lay::off_num);
    double *sred = reinterpret_cast<double *>(smem + Lay::off_red);
    double *seta = reinterpret_cast<double *>(smem + Lay::off_par);
    double *slogeta = seta + QT;
    double *sh = seta + 2 * QT;
    double *sexph = seta + 3 * QT;
    KernT<T, QT> *sK = reinterpret_cast<KernT<T, QT> *>(smem + Lay::off_k);
    unsigned *soff = reinterpret_cast<unsigned *>(smem + Lay::off_off);
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

    // stage A: index arrays + row offsets of a tile -> ring slot `s` (each thread copies what it will read)
    auto fetch_idx = [&](const Tile &t, int s) {
        if (t.ne <= unsigned(TE)) {
            unsigned *dst = sidx + size_t(s) * 3 * TE;
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < t.ne) {
                    cp_async4(dst + k, a.rev + t.e0 + k);
                    cp_async4(dst + TE + k, a.pos + t.e0 + k);
                    cp_async4(dst + 2 * TE + k, a.info + t.e0 + k);
                }
            }
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                const unsigned n = j * kThreads + tid;
                if (n <= t.nn) cp_async8(srow + size_t(s) * (TN + 8) + n, a.row_ptr + t.n0 + n);
            }
        }
        cp_async_commit();
    };
    // stage B: a tile's messages -> ring slot `s`: in-messages by gather index (slot order), old out-messages by
    // own position (buffer order; the local mirror in multi-GPU mode).  Returns own / info of the thread's entries.
    auto fetch_msgs = [&](const Tile &t, int s, unsigned (&own)[EPT], unsigned (&inf)[EPT]) {
        if (t.ne <= unsigned(TE)) {
            const unsigned *src = sidx + size_t(s) * 3 * TE;
            T *min = smsg + size_t(s) * 2 * QT * TE;
            T *mold = min + QT * TE;
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < t.ne) {
                    const unsigned g = src[k];
                    own[u] = src[TE + k];
                    inf[u] = src[2 * TE + k];
                    cp_async_vec<T, QT>(min + size_t(k) * QT, Sold + size_t(g) * Q);
                    bool remote = DIST && (own[u] & kRemoteBit);
#ifdef SBMBP_TUNING
                    if (DIST && (a.dx.dbg & 8u) && remote) {  // timing only: the outbox is not touched
                        remote = false;
                        own[u] = 0u;
                    }
#endif
                    cp_async_vec<T, QT>(mold + size_t(k) * QT,
                                        remote ? a.mirror + size_t(own[u] & ~kRemoteBit) * Q : Sold + size_t(own[u]) * Q);
                } else {
                    own[u] = 0u;
                    inf[u] = 0u;
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
    // tiles, i.e. in the same destination bucket).  Multi-GPU: whole super-tiles of `tps` consecutive tiles, strided by
    // super-tile, so that a CTA ships what it computed (dist_exchange.cuh); a wave still spans only gridDim * tps tiles.
    const unsigned G = gridDim.x;
    const unsigned tps = DIST ? a.dx.tps : 1u;
    auto nth = [&](unsigned j) -> unsigned {
        const unsigned long long t = ((unsigned long long)(j / tps) * G + blockIdx.x) * tps + (j % tps);
        return t < a.ntiles ? unsigned(t) : 0xffffffffu;
    };
    unsigned jt = 0;
    unsigned tile_id = nth(0);
    if (tile_id == 0xffffffffu) return;
    Tile t0 = a.tiles[tile_id];                                                  // tile i
    Tile t1 = (nth(1) != 0xffffffffu) ? a.tiles[nth(1)] : t0;                     // tile i+1
    Tile t2 = (nth(2) != 0xffffffffu) ? a.tiles[nth(2)] : t0;                     // tile i+2
    unsigned own[EPT], inf[EPT], own_n[EPT], inf_n[EPT];
    fetch_idx(t0, 0);
    if (nth(1) != 0xffffffffu) fetch_idx(t1, 1);
    cp_async_wait_all();
    fetch_msgs(t0, 0, own, inf);
    double cta_acc = 0.0;
    int ring = 0;  // slot of tile i; tile i+1 uses ring ^ 1

    for (; tile_id != 0xffffffffu; tile_id = nth(++jt), ring ^= 1) {
        const Tile tile = t0;
        const unsigned long long e0 = tile.e0;
        const unsigned n0 = tile.n0, nn = tile.nn, ne = tile.ne;
        const unsigned id3 = nth(jt + 3);
        const bool have1 = nth(jt + 1) != 0xffffffffu, have2 = nth(jt + 2) != 0xffffffffu;
        T *sb = smsg + size_t(ring) * 2 * QT * TE;  // in-messages of tile i; contracted in place into b_e
        const T *sold = sb + QT * TE;

        double wsum[QT];
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
        double mydiff = 0.0;

        // everything issued one iteration ago has landed: messages of tile i, indices / rows of tile i+1
        cp_async_wait_all();
        if (ne <= unsigned(TE)) {
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                const unsigned n = j * kThreads + tid;
                if (n <= nn) soff[n] = unsigned(srow[size_t(ring) * (TN + 8) + n] - e0);
            }
        }
        // put the next tile's messages in flight (ring ^ 1: its previous user, tile i-1, finished before the barrier
        // that ended the last iteration), then the indices of the tile after next into the slot tile i just vacated
        if (have1) fetch_msgs(t1, ring ^ 1, own_n, inf_n);
        if (have2) fetch_idx(t2, ring);
        Tile t3 = t0;
        if (id3 != 0xffffffffu) t3 = a.tiles[id3];

        if (ne <= unsigned(TE)) {
            // =============================================================== regular tile
            // ---- phase 1: contract in place (each thread reads and rewrites only its own slots)
            if (jt == 0) __syncthreads();  // first iteration: parameters in smem
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < ne) {
                    MsgVec<T, QT> m;
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) m.v[q] = sb[size_t(k) * QT + q];
                    T b[QT];
                    contract<T, QT>(m, sK, b);
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) sb[size_t(k) * QT + q] = b[q];
                }
            }
            __syncthreads();

            // ---- phase 2a: one thread per node of degree < 32 (product domain)
            for (unsigned n = tid; n < nn; n += kThreads) {
                const unsigned k0 = soff[n], d = soff[n + 1] - k0;
                if (d >= 32) continue;
                double tot[QT];
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) tot[q] = 1.0;
                for (unsigned k = k0; k < k0 + d; ++k) {
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) tot[q] *= double(sb[size_t(k) * QT + q]);
                }
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
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = tot[q] * rsum;
                    snum[q * TN + n] = mg.v[q];
                    wsum[q] += w * mg.v[q];
                }
                st_vec<double, QT>(mg, a.marg + size_t(n0 + n) * Q);
            }
            // ---- phase 2b: one warp per node of degree >= 32 (product below 50, log domain from 50 on)
            for (unsigned n = warp; n < (tile.nbig ? nn : 0u); n += kThreads / 32) {
                const unsigned k0 = soff[n], d = soff[n + 1] - k0;
                if (d < 32) continue;
                const bool logdom = d >= kLargeDegree;
                double acc[QT];
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) acc[q] = logdom ? 0.0 : 1.0;
                for (unsigned k = k0 + lane; k < k0 + d; k += 32) {
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        const double bv = double(sb[size_t(k) * QT + q]);
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
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned t = u * kThreads + tid;
                if (t >= ne) continue;
                const unsigned k = inf[u] & 0xffffu, n = (inf[u] >> 16) & 0x7fffu;
                T b[QT], cav[QT], oldv[QT];
                bool tiny = false;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    b[q] = sb[size_t(k) * QT + q];
                    oldv[q] = sold[size_t(t) * QT + q];
                    tiny = tiny || !(double(b[q]) >= kEps);
                }
                if (!(inf[u] & kInfLarge)) {
                    if (!tiny) {
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
                        atomicAdd(&a.ctl->tiny_count, 1ull);
                        const unsigned k0 = soff[n], d = soff[n + 1] - k0;
SBMBP_UNROLL_Q
                        for (int q = 0; q < QT; ++q) {
                            double p = 1.0;
                            for (unsigned kk = k0; kk < k0 + d; ++kk)
                                if (kk != k) p *= double(sb[size_t(kk) * QT + q]);
                            const double F = dc ? exp(-1.0 * double(d) * sh[q] / Nd) : sexph[q];
                            cav[q] = T(p * seta[q] * F);
                        }
                    }
                } else {
                    double v[QT], mx = -1.0e300;
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        v[q] = snum[q * TN + n] - log(double(b[q]));  // belief_propagation.cpp:859
                        mx = fmax(mx, v[q]);
                    }
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) cav[q] = T(exp(v[q] - mx));
                }
                T s = T(0);
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) s += cav[q];
                const T inv = fast_rcp(s);
                if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;  // non-finite message: make it visible
                MsgVec<T, QT> out;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const T nv = cav[q] * inv;
                    mydiff = fmax(mydiff, fabs(double(oldv[q]) - double(nv)));
                    out.v[q] = damp * nv + keep * oldv[q];
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
            __syncthreads();  // parameters in smem
            for (unsigned k = tid; k < ne; k += kThreads) {
                MsgVec<T, QT> m;
                ld_vec<T, QT>(m, Sold + size_t(__ldg(a.rev + e0 + k)) * Q);
                T b[QT];
                contract<T, QT>(m, sK, b);
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
                MsgVec<T, QT> m, old;
                const size_t o = size_t(__ldg(a.pos + e0 + k));  // hub tiles keep slot order
                ld_vec<T, QT>(m, Sold + size_t(__ldg(a.rev + e0 + k)) * Q);
                const bool remote = DIST && (o & kRemoteBit);
                ld_vec<T, QT>(old, remote ? a.mirror + size_t(o & ~size_t(kRemoteBit)) * Q : Sold + o * Q);
                T b[QT];
                contract<T, QT>(m, sK, b);
                double v[QT], vmx = -1.0e300;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    v[q] = (acc[q] - mx) - log(double(b[q]));
                    vmx = fmax(vmx, v[q]);
                }
                T cav[QT], s = T(0);
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    cav[q] = T(exp(v[q] - vmx));
                    s += cav[q];
                }
                const T inv = fast_rcp(s);
                MsgVec<T, QT> out;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    const T nv = cav[q] * inv;
                    mydiff = fmax(mydiff, fabs(double(old.v[q]) - double(nv)));
                    out.v[q] = damp * nv + keep * old.v[q];
                }
                if (remote) st_vec<T, QT>(out, a.mirror + size_t(o & ~size_t(kRemoteBit)) * Q);  // outbox
                else st_vec<T, QT>(out, Snew + o * Q);
            }
        }

        // ---- tile epilogue: the tile's row (field partials, max-diff) into the CTA's running row.  The barriers here
        // also close the tile: nobody reads its b_e / old values / soff afterwards, so the ring slot can be refilled.
        mydiff = warp_max(mydiff);
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) wsum[q] = warp_sum(wsum[q]);
        __syncthreads();
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
            cta_acc = (tid < QT) ? cta_acc + v : fmax(cta_acc, v);
        }
        if constexpr (DIST) {
            // the tile's remote out-messages are in the outbox (every thread's stores precede the barrier above)
            if ((jt + 1) % tps == 0 || !have1) {  // last tile of one of this CTA's super-tiles: carry it to the owners
                const unsigned sp = tile_id / tps;
                if constexpr ((QT * sizeof(T)) % 16 == 0) {
                    if (a.dx.ship_tma)
                        dist_ship_supertile_tma<T, QT, kThreads>(a.dx, sp, a.mirror, s_peer, reinterpret_cast<unsigned char *>(sb), unsigned(2 * Lay::msg_bytes));
                    else dist_ship_range<T, QT, kThreads>(a.dx, a.dx.out_start[sp], a.dx.out_start[sp + 1], a.mirror, s_peer);
                } else {
                    dist_ship_range<T, QT, kThreads>(a.dx, a.dx.out_start[sp], a.dx.out_start[sp + 1], a.mirror, s_peer);
                }
            }
        }
        // rotate the pipeline registers
        t0 = t1;
        t1 = t2;
        t2 = t3;
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            own[u] = own_n[u];
            inf[u] = inf_n[u];
        }
    }
    cp_async_wait_all();
    if constexpr (DIST) {
        dist_ship_drain();  // what this CTA shipped has landed at its owners
    }
    if (tid <= QT) a.partial[size_t(blockIdx.x) * (QT + 1) + tid] = cta_acc;  // one row per CTA
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

}  // namespace sbmbp
