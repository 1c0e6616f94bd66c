// Warp-autonomous sweep kernel for small Q (QT <= 4): same arithmetic as the tile kernels (sweep_tile.cuh /
// sweep_pipe_dist.cuh), but the unit of work is a WARP tile -- a node-aligned run of <= 32 nodes and <= WE = 32 * EPL
// edge slots -- and a warp carries its tile from gather to store on its own: no CTA-wide barrier inside the
// sweep, 21 KB of static shared memory per CTA, and therefore 3-4x the resident warps of the cp.async pipeline
// kernel (which sat at 16 warps / SM waiting on five barriers per tile, profiles/ncu_sweep_cfg2_f64_r01.md).
//
// Per warp tile (reference: sum_all_messages_to_i / norm_m_at_i / bp_iter_update_psi_large_degree,
// belief_propagation.cpp:991-1071, :813-890, evaluated synchronously):
//   gather   lane l loads the in-messages of slots l, l+32, ... (the one random access, rev[] = buffer position)
//   contract b_e[q] = sum_t K[t][q] psi_e[t]  -> the warp's b-array in shared memory (slot order)
//   node     kind 0: one lane per node (all degrees < 32, product domain);
//            kind 1: one node of degree 32..WE, the whole warp (product below 50, sum of logs from 50 on)
//            -> marginal (coalesced write), w_i psi_i into the lane's running field sum, node total to smem
//   edge     in BUFFER order (pos[] ascending inside the tile, one 16-bit word per entry says which slot / node it
//            is): leave-one-out, normalise, max |old - new|, damped write.  Old values are loaded right after the
//            contraction so the node phase hides their latency.
// While a warp computes tile i, the index words of tile i+1 and the descriptor of tile i+2 are in flight.
// Nodes with more than WE edges are hubs: bp_sweep_hub_kernel (one CTA each, two passes in the log domain),
// launched just before this kernel; both leave one row of field partials per CTA and the last CTA of THIS kernel
// closes the sweep over all rows in a fixed order (bitwise reproducible run to run).
#pragma once
#include "bp_device.cuh"
#include "sweep_tile.cuh"

namespace sbmbp {

#ifndef SBMBP_WARP_MINB
#define SBMBP_WARP_MINB 3
#endif

template <typename T>
struct WarpSweepArgs {
    const WTile *tiles;
    unsigned ntiles;
    const Tile *hubs;  // nodes with more than WE edges (nn == 1)
    unsigned nhubs;
    unsigned hub_rows;      // CTAs of the hub kernel = rows it leaves
    unsigned warp_rows;     // CTAs of the warp kernel = rows it leaves
    unsigned row_base;      // first row of the warp kernel in partial[]
    unsigned hub_row_base;  // first row of the hub kernel
    int close;              // the warp kernel is the last launch of the sweep: its last CTA closes it (rows 0 .. warp_rows + hub_rows)
    const unsigned long long *row_ptr;
    const unsigned *rev;           // slot order: buffer position of the message INTO row(e) along e
    const unsigned *pos;           // buffer positions of the out-messages, ascending inside each warp tile (hubs: slot order)
    const unsigned short *info;    // per pos entry: tile-local slot | tile-local node << 7
    T *S[2];
    double *marg;
    const DevParams *prm;
    Field *field[2];
    Ctl *ctl;
    double *partial;  // [warp_rows + hub_rows][QT + 1]
    unsigned dc;
    double damping;
};

__device__ __forceinline__ WTile ld_wtile(const WTile *p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    WTile w;
    w.e0lo = v.x;
    w.e0hi = v.y;
    w.n0 = v.z;
    w.packed = v.w;
    return w;
}

template <typename T, int QT>
__device__ __forceinline__ void lds_vec(T (&v)[QT], const T *p) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i) reinterpret_cast<uint4 *>(v)[i] = reinterpret_cast<const uint4 *>(p)[i];
    } else {
        static_assert(bytes == 8, "Q x sizeof(T) must be 8 or a multiple of 16");
        *reinterpret_cast<uint2 *>(v) = *reinterpret_cast<const uint2 *>(p);
    }
}
template <typename T, int QT>
__device__ __forceinline__ void sts_vec(T *p, const T (&v)[QT]) {
    constexpr int bytes = QT * int(sizeof(T));
    if constexpr (bytes % 16 == 0) {
#pragma unroll
        for (int i = 0; i < bytes / 16; ++i) reinterpret_cast<uint4 *>(p)[i] = reinterpret_cast<const uint4 *>(v)[i];
    } else {
        *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(v);
    }
}

// per-warp shared memory: index words of this tile and the next (ring of 2), b-array, old values, node totals, offsets
template <typename T, int QT>
struct WarpSmem {
    using Cfg = WarpCfg<T, QT>;
    static constexpr int WE = Cfg::WE;
    static constexpr size_t idx_bytes = sizeof(unsigned) * 2 * WE + sizeof(unsigned long long) * 32;  // rev, pos, row_ptr slice
    static constexpr size_t off_idx = 0;                                          // [2][idx_bytes]
    static constexpr size_t off_b = off_idx + 2 * idx_bytes;                      // T[WE * QT]
    static constexpr size_t off_old = off_b + sizeof(T) * WE * QT;                // T[WE * QT]
    static constexpr size_t off_num = off_old + sizeof(T) * WE * QT;              // double[32 * QT]
    static constexpr size_t off_off = off_num + sizeof(double) * 32 * QT;         // u32[36]
    static constexpr size_t per_warp = (off_off + sizeof(unsigned) * 36 + 15) & ~size_t(15);
    static constexpr size_t bytes = per_warp * (kThreads / 32);
};

template <typename T, int QT>
__global__ void __launch_bounds__(kThreads, SBMBP_WARP_MINB) bp_sweep_warp_kernel(const WarpSweepArgs<T> a) {
    using Cfg = WarpCfg<T, QT>;
    using Lay = WarpSmem<T, QT>;
    constexpr int EPL = Cfg::EPL, WE = Cfg::WE, NW = kThreads / 32;
    constexpr unsigned Q = QT;
    static_assert(QT <= 4, "the warp kernel is the small-Q path");
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(16) T s_K[QT * QT];
    __shared__ double s_par[4 * QT];
    __shared__ double s_rows[NW][QT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nwt = gridDim.x * NW, ntiles = a.ntiles;
    unsigned t = blockIdx.x * NW + warp;
    unsigned char *wbase = smem + size_t(warp) * Lay::per_warp;
    T *sb = reinterpret_cast<T *>(wbase + Lay::off_b);
    T *sold = reinterpret_cast<T *>(wbase + Lay::off_old);
    double *snum = reinterpret_cast<double *>(wbase + Lay::off_num);
    unsigned *soff = reinterpret_cast<unsigned *>(wbase + Lay::off_off);

    // index words and row offsets of a tile -> ring slot s (each lane copies exactly what it reads back)
    auto fetch_idx = [&](const WTile &w, int s) {
        unsigned *dst = reinterpret_cast<unsigned *>(wbase + Lay::off_idx + size_t(s) * Lay::idx_bytes);
        const unsigned long long e0 = w.e0();
        const unsigned ne = w.ne();
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            const unsigned k = u * 32 + lane;
            if (k < ne) {
                cp_async4(dst + k, a.rev + e0 + k);
                cp_async4(dst + WE + k, a.pos + e0 + k);
            }
        }
        if (unsigned(lane) < w.nn())
            cp_async8(reinterpret_cast<unsigned long long *>(dst + 2 * WE) + lane, a.row_ptr + w.n0 + lane);
        cp_async_commit();
    };

    // ---- loads that do not depend on the control block go first: tile descriptors and the first tile's index words
    WTile cur, nxt;
    cur.e0lo = cur.e0hi = cur.n0 = cur.packed = 0u;  // an empty tile: nothing to load, nothing to do
    nxt = cur;
    if (ntiles) {
        cur = ld_wtile(a.tiles + min(t, ntiles - 1u));
        nxt = ld_wtile(a.tiles + min(t + nwt, ntiles - 1u));
    }
    fetch_idx(cur, 0);

    Ctl *ctl = a.ctl;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) {  // uniform over the grid
        cp_async_wait_all();
        return;
    }
    const int par = int(sweeps_done & 1u);
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const Field *fld = par ? a.field[1] : a.field[0];
    const bool dc = a.dc != 0;
    const double Nd = a.prm->N;
    const T damp = T(a.damping), keep = T(1.0 - a.damping);

    for (int i = tid; i < QT * QT; i += kThreads) s_K[i] = T(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) {
        s_par[tid] = a.prm->eta[tid];
        s_par[QT + tid] = a.prm->logeta[tid];
        s_par[2 * QT + tid] = fld->h[tid];
        s_par[3 * QT + tid] = fld->exph[tid];
    }
    __syncthreads();
    const double *seta = s_par, *slogeta = s_par + QT, *sh = s_par + 2 * QT, *sexph = s_par + 3 * QT;

    double wsum[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;
    int ring = 0;

    for (; t < ntiles; t += nwt, ring ^= 1) {
        const unsigned ne = cur.ne(), nn = cur.nn(), n0 = cur.n0;
        const unsigned long long e0 = cur.e0();
        const bool wide = cur.kind() != 0u;
        const unsigned *sidx = reinterpret_cast<const unsigned *>(wbase + Lay::off_idx + size_t(ring) * Lay::idx_bytes);

        cp_async_wait_all();  // this tile's index words (issued one tile ago) have landed

        // ---- gather (slot order) into registers; old values (buffer order) and this tile's slot/node words ride along
        MsgVec<T, QT> m[EPL];
        unsigned inf[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            const unsigned k = u * 32 + lane;
            inf[u] = 0u;
            if (k < ne) {
                ld_vec<T, QT>(m[u], Sold + size_t(sidx[k]) * Q);
                cp_async_vec<T, QT>(sold + size_t(k) * QT, Sold + size_t(sidx[WE + k]) * Q);
                inf[u] = __ldg(a.info + e0 + k);
            }
        }
        cp_async_commit();  // group: old values

        // ---- keep the pipeline fed: index words of the next tile, descriptor two tiles ahead
        if (t + nwt < ntiles) fetch_idx(nxt, ring ^ 1);
        else cp_async_commit();
        WTile nx2 = cur;
        if (t + 2u * nwt < ntiles) nx2 = ld_wtile(a.tiles + t + 2u * nwt);

        // ---- row offsets; contract -> b (shared, slot order)
        if (unsigned(lane) < nn) soff[lane] = unsigned(reinterpret_cast<const unsigned long long *>(sidx + 2 * WE)[lane] - e0);
        if (lane == 0) soff[nn] = ne;
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            const unsigned k = u * 32 + lane;
            if (k < ne) {
                T b[QT];
                contract<T, QT>(m[u], s_K, b);
                sts_vec<T, QT>(sb + size_t(k) * QT, b);
            }
        }
        __syncwarp();

        // ---- node phase
        bool logdom = false;
        if (!wide) {
            if (unsigned(lane) < nn) {
                const unsigned k0 = soff[lane], d = soff[lane + 1] - k0;
                double tot[QT];
#pragma unroll
                for (int q = 0; q < QT; ++q) tot[q] = 1.0;
                for (unsigned k = k0; k < k0 + d; ++k) {
                    T bv[QT];
                    lds_vec<T, QT>(bv, sb + size_t(k) * QT);
#pragma unroll
                    for (int q = 0; q < QT; ++q) tot[q] *= double(bv[q]);
                }
                double sum = 0.0;
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    const double F = dc ? exp(-1.0 * double(d) * sh[q] / Nd) : sexph[q];
                    tot[q] = tot[q] * seta[q] * F;
                    sum += tot[q];
                }
                const double w = dc ? double(d) : 1.0;
                const double rsum = fast_rcp(sum);
                MsgVec<double, QT> mg;
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = tot[q] * rsum;
                    wsum[q] += w * mg.v[q];
                }
                sts_vec<double, QT>(snum + lane * QT, mg.v);
                st_vec<double, QT>(mg, a.marg + size_t(n0 + lane) * Q);
            }
        } else {
            const unsigned d = ne;
            logdom = d >= kLargeDegree;
            double acc[QT];
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[q] = logdom ? 0.0 : 1.0;
            for (unsigned k = lane; k < d; k += 32) {
                T bv[QT];
                lds_vec<T, QT>(bv, sb + size_t(k) * QT);
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    if (logdom) acc[q] += log(double(bv[q]));
                    else acc[q] *= double(bv[q]);
                }
            }
            double mx = -1.0e300, sum = 0.0;
#pragma unroll
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
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    mg.v[q] = exp(acc[q] - mx);
                    sum += mg.v[q];
                }
            }
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                mg.v[q] = (logdom ? mg.v[q] : acc[q]) / sum;
                if (lane == 0) snum[q] = logdom ? acc[q] - mx : mg.v[q];
            }
            if (lane == 0) {
                const double w = dc ? double(d) : 1.0;
#pragma unroll
                for (int q = 0; q < QT; ++q) wsum[q] += w * mg.v[q];
                st_vec<double, QT>(mg, a.marg + size_t(n0) * Q);
            }
        }
        cp_async_wait_group1();  // the old values (the group before the next tile's index words) have landed
        __syncwarp();

        // ---- edge phase (buffer order): leave-one-out, normalise, max-diff, damped write
#pragma unroll
        for (int u = 0; u < EPL; ++u) {
            const unsigned tt = u * 32 + lane;
            if (tt >= ne) continue;
            const unsigned k = inf[u] & kWInfoSlotMask, n = (inf[u] >> kWInfoNodeShift) & 31u;
            T b[QT], cav[QT], oldv[QT];
            lds_vec<T, QT>(b, sb + size_t(k) * QT);
            lds_vec<T, QT>(oldv, sold + size_t(tt) * QT);
            double tot[QT];
            lds_vec<double, QT>(tot, snum + n * QT);
            bool tiny = false;
#pragma unroll
            for (int q = 0; q < QT; ++q) tiny = tiny || !(double(b[q]) >= kEps);
            if (!logdom) {
                if (!tiny) {
#pragma unroll
                    for (int q = 0; q < QT; ++q) {
                        T c = T(tot[q]);
#pragma unroll
                        for (int r = 0; r < QT; ++r)
                            if (r != q) c *= b[r];
                        cav[q] = c;
                    }
                } else {
                    // a vanishing b_e[q]: leave-one-out product taken directly (see sweep_kernel.cuh)
                    atomicAdd(&a.ctl->tiny_count, 1ull);
                    const unsigned k0 = soff[n], d = soff[n + 1] - k0;
#pragma unroll
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
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    v[q] = tot[q] - log(double(b[q]));  // belief_propagation.cpp:859
                    mx = fmax(mx, v[q]);
                }
#pragma unroll
                for (int q = 0; q < QT; ++q) cav[q] = T(exp(v[q] - mx));
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
                mydiff = fmax(mydiff, fabs(double(oldv[q]) - double(nv)));
                out.v[q] = damp * nv + keep * oldv[q];
            }
            st_vec<T, QT>(out, Snew + size_t(sidx[WE + tt]) * Q);
        }
        __syncwarp();  // the warp's b / totals / offsets are free for the next tile

        cur = nxt;
        nxt = nx2;
    }
    cp_async_wait_all();

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
        a.partial[size_t(a.row_base + blockIdx.x) * (QT + 1) + tid] = v;
    }
    if (a.close) {
        SweepArgsBase base;
        base.prm = a.prm;
        base.field[0] = a.field[0];
        base.field[1] = a.field[1];
        base.ctl = a.ctl;
        base.partial = a.partial;
        close_sweep_last_cta<QT>(base, gridDim.x + a.hub_rows, sweeps_done, nullptr, gridDim.x);
    }
}

// Hubs (degree > WE): one CTA per node at a time, two passes in the log domain (belief_propagation.cpp:813-890).
// CTA c leaves its row at partial[warp_rows + c]; the warp kernel launched next on the same stream closes the sweep.
template <typename T, int QT>
__global__ void __launch_bounds__(kThreads) bp_sweep_hub_kernel(const WarpSweepArgs<T> a) {
    constexpr unsigned Q = QT;
    __shared__ __align__(16) T s_K[QT * QT];
    __shared__ double s_par[4 * QT];
    __shared__ double sred[kThreads / 32];

    Ctl *ctl = a.ctl;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;
    const int par = int(sweeps_done & 1u);
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const Field *fld = par ? a.field[1] : a.field[0];
    const bool dc = a.dc != 0;
    const int tid = threadIdx.x;
    const double Nd = a.prm->N;
    const T damp = T(a.damping), keep = T(1.0 - a.damping);
    for (int i = tid; i < QT * QT; i += kThreads) s_K[i] = T(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) {
        s_par[tid] = a.prm->eta[tid];
        s_par[QT + tid] = a.prm->logeta[tid];
        s_par[2 * QT + tid] = fld->h[tid];
        s_par[3 * QT + tid] = fld->exph[tid];
    }
    __syncthreads();
    const double *slogeta = s_par + QT, *sh = s_par + 2 * QT;

    double wsum[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;

    for (unsigned hb = blockIdx.x; hb < a.nhubs; hb += gridDim.x) {
        const Tile tile = a.hubs[hb];
        const unsigned long long e0 = tile.e0;
        const unsigned ne = tile.ne;
        const double dd = double(ne);
        double acc[QT];
#pragma unroll
        for (int q = 0; q < QT; ++q) acc[q] = 0.0;
        for (unsigned k = tid; k < ne; k += kThreads) {
            MsgVec<T, QT> m;
            ld_vec<T, QT>(m, Sold + size_t(__ldg(a.rev + e0 + k)) * Q);
            T b[QT];
            contract<T, QT>(m, s_K, b);
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[q] += log(double(b[q]));
        }
        double mx = -1.0e300;
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            acc[q] = block_sum(acc[q], sred) + slogeta[q] - (dc ? 1.0 * dd * sh[q] / Nd : sh[q] / Nd);
            mx = fmax(mx, acc[q]);
        }
        double sum = 0.0;
        MsgVec<double, QT> mg;
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            mg.v[q] = exp(acc[q] - mx);
            sum += mg.v[q];
        }
#pragma unroll
        for (int q = 0; q < QT; ++q) mg.v[q] /= sum;
        if (tid == 0) {
            const double w = dc ? dd : 1.0;
#pragma unroll
            for (int q = 0; q < QT; ++q) wsum[q] += w * mg.v[q];
            st_vec<double, QT>(mg, a.marg + size_t(tile.n0) * Q);
        }
        for (unsigned k = tid; k < ne; k += kThreads) {
            MsgVec<T, QT> m, old;
            const size_t o = size_t(__ldg(a.pos + e0 + k));  // hubs keep slot order
            ld_vec<T, QT>(m, Sold + size_t(__ldg(a.rev + e0 + k)) * Q);
            ld_vec<T, QT>(old, Sold + o * Q);
            T b[QT];
            contract<T, QT>(m, s_K, b);
            double v[QT], vmx = -1.0e300;
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                v[q] = (acc[q] - mx) - log(double(b[q]));
                vmx = fmax(vmx, v[q]);
            }
            T cav[QT], s = T(0);
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                cav[q] = T(exp(v[q] - vmx));
                s += cav[q];
            }
            const T inv = fast_rcp(s);
            if (!(inv == inv) || !(double(inv) <= 1.0e300)) mydiff = 1.0e300;
            MsgVec<T, QT> out;
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                const T nv = cav[q] * inv;
                mydiff = fmax(mydiff, fabs(double(old.v[q]) - double(nv)));
                out.v[q] = damp * nv + keep * old.v[q];
            }
            st_vec<T, QT>(out, Snew + o * Q);
        }
    }
    mydiff = block_max(mydiff, sred);
    if (tid == 0) {
        double *row = a.partial + size_t(a.hub_row_base + blockIdx.x) * (QT + 1);
#pragma unroll
        for (int q = 0; q < QT; ++q) row[q] = wsum[q];
        row[QT] = mydiff;
    }
}

}  // namespace sbmbp
