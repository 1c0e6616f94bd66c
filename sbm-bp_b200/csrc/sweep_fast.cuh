// Lean instantiation of the BP sweep for the common case: Q equals the compiled width QT, deg_corr_flag 0 or 1,
// and one Q x Q kernel for every degree class (beta == 1, or dc == 1 where beta never enters).  Same algorithm,
// same tiles and same smem carve-up as bp_sweep_kernel (sweep_kernel.cuh) -- that one stays as the general
// path (padded Q, dc == 2, beta != 1) -- but with the run-time generality stripped from the per-edge code:
//   * per buffer entry ONE packed word says which tile-local slot and node it belongs to and whether that node
//     updates in the log domain, so phase 3 needs no edge->node table and no degree lookups
//   * the leave-one-out division becomes a product for Q <= 4:  psi~_q / b_q  ~  psi~_q * prod_{q' != q} b_q'
//     (the common factor prod_q b_q cancels in the normalisation), leaving one reciprocal per edge instead of
//     Q + Q long FP64 division chains
//   * all index loads of a thread are issued before anything waits, the old values ride with the gathers
// Reference: sum_all_messages_to_i / norm_m_at_i / bp_iter_update_psi_large_degree
// (belief_propagation.cpp:991-1071, :813-890), evaluated synchronously.
#pragma once
#include "bp_device.cuh"
#include "sweep_kernel.cuh"

namespace sbmbp {

constexpr unsigned kInfLarge = 0x80000000u;  // info word: bit 31 = node updates in the log domain (degree >= 50)

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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// extra shared memory of the persistent fast kernel, after the TileSmem carve-up
template <typename T, int QT>
struct FastSmem {
    using Cfg = TileCfg<T, QT>;
    static constexpr size_t off_idx = (TileSmem<T, QT>::bytes + 15) & ~size_t(15);        // u32[3][TE]: gather, own, info
    static constexpr size_t off_row = off_idx + sizeof(unsigned) * 3 * Cfg::TE;           // u64[TN + 8]: row_ptr slice
    static constexpr size_t bytes = off_row + sizeof(unsigned long long) * (Cfg::TN + 8);
};

// Persistent: gridDim.x CTAs stride over the tiles.  While a CTA works on tile t it has the index arrays and row
// offsets of its next tile in flight (cp.async into shared memory) and that tile's descriptor in registers, so the
// only memory round trip left on a tile's critical path is the message gather itself.
constexpr unsigned kPosBits = 29;  // multi-GPU plan (host side, mirror pull): owner rank << 29 | position on the owner
constexpr unsigned kPosMask = (1u << kPosBits) - 1u;

// DIST = false: single GPU, old values read from S_old[pos], new values written to S_new[pos].
// DIST = true : the message buffers belong to the DESTINATION's rank.  New values go straight into the owner's
//   S_new through its CUDA-IPC mapping (NVLink peer stores, posted -- they overlap the rest of the tile), the old
//   values come from a local mirror of this rank's out-messages (coalesced), which is updated in place.
template <typename T, int QT, bool DIST>
__global__ void __launch_bounds__(kThreads, SBMBP_MINB) bp_sweep_fast_kernel(const SweepArgs<T> a) {
    using Cfg = TileCfg<T, QT>;
    using Lay = TileSmem<T, QT>;
    using Ext = FastSmem<T, QT>;
    constexpr int TE = Cfg::TE, TN = Cfg::TN;
    constexpr int EPT = TE / kThreads;
    constexpr int NPT = (TN + 1 + kThreads - 1) / kThreads;  // row-offset entries per thread
    constexpr unsigned Q = QT;
    extern __shared__ __align__(16) unsigned char smem[];
    double *snum = reinterpret_cast<double *>(smem + Lay::off_num);
    double *sred = reinterpret_cast<double *>(smem + Lay::off_red);
    double *seta = reinterpret_cast<double *>(smem + Lay::off_par);
    double *slogeta = seta + QT;
    double *sh = seta + 2 * QT;
    double *sexph = seta + 3 * QT;
    KernT<T, QT> *sK = reinterpret_cast<KernT<T, QT> *>(smem + Lay::off_ks);
    T *sb = reinterpret_cast<T *>(smem + Lay::off_b);
    unsigned *soff = reinterpret_cast<unsigned *>(smem + Lay::off_off);
    unsigned *sidx = reinterpret_cast<unsigned *>(smem + Ext::off_idx);
    unsigned long long *srow = reinterpret_cast<unsigned long long *>(smem + Ext::off_row);

    Ctl *ctl = a.ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // multi-GPU, inside a batch: close the previous sweep from all ranks' rows first (see sweep_pipe.cuh / dist_exchange.cuh)
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

    // issue the cp.async prefetch of one tile's index arrays and row offsets (each thread copies what it will read)
    auto prefetch = [&](const Tile &t) {
        if (t.ne <= unsigned(TE)) {
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < t.ne) {
                    cp_async4(sidx + k, a.rev + t.e0 + k);
                    cp_async4(sidx + TE + k, a.pos + t.e0 + k);
                    cp_async4(sidx + 2 * TE + k, a.info + t.e0 + k);
                }
            }
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                const unsigned n = j * kThreads + tid;
                if (n <= t.nn) cp_async8(srow + n, a.row_ptr + t.n0 + n);
            }
        }
        cp_async_commit();
    };

    // multi-GPU shipping state (dist_exchange.cuh)
    __shared__ T *s_peer[kMaxRanks];
    if constexpr (DIST) {
        if (tid < kMaxRanks) s_peer[tid] = (par ? a.peer[0] : a.peer[1])[tid];
    }
    // the j-th tile of this CTA: strided on one GPU, whole super-tiles in multi-GPU mode (see sweep_pipe.cuh)
    const unsigned tps = DIST ? a.dx.tps : 1u;
    auto nth = [&](unsigned j) -> unsigned {
        const unsigned long long t = ((unsigned long long)(j / tps) * gridDim.x + blockIdx.x) * tps + (j % tps);
        return t < a.ntiles ? unsigned(t) : 0xffffffffu;
    };
    unsigned jt = 0;
    unsigned tile_id = nth(0);
    if (tile_id == 0xffffffffu) return;
    Tile cur = a.tiles[tile_id];
    Tile nxt = cur;
    if (nth(1) != 0xffffffffu) nxt = a.tiles[nth(1)];
    prefetch(cur);
    double cta_acc = 0.0;  // threads 0..QT: this CTA's running row over its tiles (fixed order: bitwise reproducible)

    for (; tile_id != 0xffffffffu; tile_id = nth(++jt)) {
    const Tile tile = cur;
    const unsigned long long e0 = tile.e0;
    const unsigned n0 = tile.n0, nn = tile.nn, ne = tile.ne;
    const bool have_next = nth(jt + 1) != 0xffffffffu;
    const unsigned id2 = nth(jt + 2);

    double wsum[QT];
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;

    if (ne <= unsigned(TE)) {
        // =================================================================== regular tile
        unsigned gat[EPT], own[EPT], inf[EPT];
        cp_async_wait_all();
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const unsigned k = u * kThreads + tid;
            const bool live = k < ne;
            gat[u] = live ? sidx[k] : 0u;
            own[u] = live ? sidx[TE + k] : 0u;
            inf[u] = live ? sidx[2 * TE + k] : 0u;
        }
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            const unsigned n = j * kThreads + tid;
            if (n <= nn) soff[n] = unsigned(srow[n] - e0);
        }
        // the staging slots this thread just read are free again: start the next tile's prefetch
        if (have_next) prefetch(nxt);

        // ---- phase 1: gather (slot order) + contract; the old values of phase 3 (buffer order) ride along
        MsgVec<T, QT> oldv[EPT];
        {
            MsgVec<T, QT> m[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (u * kThreads + tid < ne) ld_vec<T, QT>(m[u], Sold + size_t(gat[u]) * Q);
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (u * kThreads + tid < ne)
                    ld_vec<T, QT>(oldv[u], (DIST && (own[u] & kRemoteBit)) ? a.mirror + size_t(own[u] & ~kRemoteBit) * Q : Sold + size_t(own[u]) * Q);
            if (id2 != 0xffffffffu) cur = a.tiles[id2];  // used next iteration
            __syncthreads();  // row offsets (and, first time round, the parameters) in smem
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const unsigned k = u * kThreads + tid;
                if (k < ne) {
                    T b[QT];
                    contract<T, QT>(m[u], sK, b);
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
                    const double bv = double(sb[q * TE + k]);
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
            if (u * kThreads + tid >= ne) continue;
            const unsigned k = inf[u] & 0xffffu, n = (inf[u] >> 16) & 0x7fffu;
            T b[QT], cav[QT];
            bool tiny = false;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                b[q] = sb[q * TE + k];
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
                    // a vanishing b_e[q]: leave-one-out product taken directly (see sweep_kernel.cuh)
                    atomicAdd(&a.ctl->tiny_count, 1ull);
                    const unsigned k0 = soff[n], d = soff[n + 1] - k0;
SBMBP_UNROLL_Q
                    for (int q = 0; q < QT; ++q) {
                        double p = 1.0;
                        for (unsigned kk = k0; kk < k0 + d; ++kk)
                            if (kk != k) p *= double(sb[q * TE + kk]);
                        const double F = dc ? exp(-1.0 * double(d) * sh[q] / Nd) : sexph[q];
                        cav[q] = T(p * seta[q] * F);
                    }
                }
            } else {
                double v[QT], mx = -1.0e300;
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) {
                    v[q] = snum[q * TN + n] - log(double(b[q]));  // :859
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
                mydiff = fmax(mydiff, fabs(double(oldv[u].v[q]) - double(nv)));
                out.v[q] = damp * nv + keep * oldv[u].v[q];
            }
            if (DIST && (own[u] & kRemoteBit)) st_vec<T, QT>(out, a.mirror + size_t(own[u] & ~kRemoteBit) * Q);  // outbox
            else st_vec<T, QT>(out, Snew + size_t(own[u]) * Q);
        }
    } else {
        // =================================================================== hub node (degree > TE): log domain
        cp_async_wait_all();
        if (have_next) prefetch(nxt);
        if (id2 != 0xffffffffu) cur = a.tiles[id2];
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

    // ---- tile epilogue: reduce the field partials and the max-diff over the CTA into the CTA's running row
    // (stored once per CTA at the end; bp_finalize_kernel closes the sweep)
    mydiff = warp_max(mydiff);
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = warp_sum(wsum[q]);
    __syncthreads();  // phase 3 is done with sb / snum / soff; sred is free
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
        // last tile of one of this CTA's super-tiles: carry its remote out-messages to the owners (dist_exchange.cuh)
        if ((jt + 1) % tps == 0 || !have_next) {
            const unsigned sp = tile_id / tps;
            dist_ship_range<T, QT, kThreads>(a.dx, a.dx.out_start[sp], a.dx.out_start[sp + 1], a.mirror, s_peer);
        }
    }
    {   // rotate the descriptors: `cur` already holds the tile after next (loaded above)
        const Tile after = cur;
        cur = nxt;
        nxt = after;
    }
    }  // tile loop
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
