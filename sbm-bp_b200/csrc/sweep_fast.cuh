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

template <typename T, int QT>
__global__ void __launch_bounds__(kThreads) bp_sweep_fast_kernel(const SweepArgs<T> a) {
    using Cfg = TileCfg<T, QT>;
    using Lay = TileSmem<T, QT>;
    constexpr int TE = Cfg::TE, TN = Cfg::TN;
    constexpr int EPT = TE / kThreads;
    constexpr unsigned Q = QT;
    extern __shared__ __align__(16) unsigned char smem[];
    double *snum = reinterpret_cast<double *>(smem + Lay::off_num);
    double *sred = reinterpret_cast<double *>(smem + Lay::off_red);
    double *seta = reinterpret_cast<double *>(smem + Lay::off_par);
    double *slogeta = seta + QT;
    double *sh = seta + 2 * QT;
    double *sexph = seta + 3 * QT;
    T *sK = reinterpret_cast<T *>(smem + Lay::off_ks);
    T *sb = reinterpret_cast<T *>(smem + Lay::off_b);
    unsigned *soff = reinterpret_cast<unsigned *>(smem + Lay::off_off);
    __shared__ int s_last;

    Ctl *ctl = a.ctl;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (ctl->converged || sweeps_done >= ctl->max_sweeps) return;  // uniform over the grid
    const int par = int(sweeps_done & 1u);
    const T *__restrict__ Sold = par ? a.S[1] : a.S[0];
    T *__restrict__ Snew = par ? a.S[0] : a.S[1];
    const Field *fld = par ? a.field[1] : a.field[0];
    Field *fld_next = par ? a.field[0] : a.field[1];
    const bool dc = a.dc != 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double Nd = a.prm->N;
    const T damp = T(a.damping), keep = T(1.0 - a.damping);

    const Tile tile = a.tiles[blockIdx.x];
    const unsigned long long e0 = tile.e0;
    const unsigned n0 = tile.n0, nn = tile.nn, ne = tile.ne;

    for (int i = tid; i < QT * QT; i += kThreads) sK[i] = T(a.prm->Ks[(i / QT) * kMaxQ + (i % QT)]);
    if (tid < QT) {
        seta[tid] = a.prm->eta[tid];
        slogeta[tid] = a.prm->logeta[tid];
        sh[tid] = fld->h[tid];
        sexph[tid] = fld->exph[tid];
    }

    double wsum[QT];
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) wsum[q] = 0.0;
    double mydiff = 0.0;

    if (ne <= unsigned(TE)) {
        // =================================================================== regular tile
        unsigned gat[EPT], own[EPT], inf[EPT];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const unsigned k = u * kThreads + tid;
            const bool live = k < ne;
            gat[u] = live ? __ldg(a.rev + e0 + k) : 0u;
            own[u] = live ? __ldg(a.pos + e0 + k) : 0u;
            inf[u] = live ? __ldg(a.info + e0 + k) : 0u;
        }
        for (unsigned n = tid; n <= nn; n += kThreads) soff[n] = unsigned(a.row_ptr[n0 + n] - e0);

        // ---- phase 1: gather (slot order) + contract; the old values of phase 3 (buffer order) ride along
        MsgVec<T, QT> oldv[EPT];
        {
            MsgVec<T, QT> m[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (u * kThreads + tid < ne) ld_vec<T, QT>(m[u], Sold + size_t(gat[u]) * Q);
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (u * kThreads + tid < ne) ld_vec<T, QT>(oldv[u], Sold + size_t(own[u]) * Q);
            __syncthreads();  // parameters and row offsets in smem
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
            MsgVec<double, QT> mg;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                mg.v[q] = tot[q] / sum;
                snum[q * TN + n] = mg.v[q];
                wsum[q] += w * mg.v[q];
            }
            st_vec<double, QT>(mg, a.marg + size_t(n0 + n) * Q);
        }
        // ---- phase 2b: one warp per node of degree >= 32 (product below 50, log domain from 50 on)
        for (unsigned n = warp; n < nn; n += kThreads / 32) {
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
            const T inv = T(1) / s;
            MsgVec<T, QT> out;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                const T nv = cav[q] * inv;
                mydiff = fmax(mydiff, fabs(double(oldv[u].v[q]) - double(nv)));
                out.v[q] = damp * nv + keep * oldv[u].v[q];
            }
            st_vec<T, QT>(out, Snew + size_t(own[u]) * Q);
        }
    } else {
        // =================================================================== hub node (degree > TE): log domain
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
            ld_vec<T, QT>(old, Sold + o * Q);
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
            const T inv = T(1) / s;
            MsgVec<T, QT> out;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                const T nv = cav[q] * inv;
                mydiff = fmax(mydiff, fabs(double(old.v[q]) - double(nv)));
                out.v[q] = damp * nv + keep * old.v[q];
            }
            st_vec<T, QT>(out, Snew + o * Q);
        }
    }

    // ---- CTA epilogue: one barrier for the field partials and the max-diff, then last-CTA finalisation
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
    if (tid < QT) {  // fixed order over the warps: bitwise reproducible
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += sred[w * (QT + 1) + tid];
        a.partial[size_t(blockIdx.x) * QT + tid] = v;
        __threadfence();
    }
    if (tid == QT) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v = fmax(v, sred[w * (QT + 1) + QT]);
        if (!(v == v) || v > 1.0e300) atomicAdd(&ctl->nan_count, 1ull);
        atomicMax(&ctl->maxdiff_bits, (unsigned long long)__double_as_longlong(v));
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&ctl->done, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[QT];
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) {
        double p = 0.0;
        for (unsigned b = tid; b < a.ntiles; b += kThreads) p += __ldcg(a.partial + size_t(b) * QT + q);
        tot[q] = block_sum(p, sred);
    }
    if (tid == 0) {
        publish_field(a.prm, Q, tot, fld_next);
        const double md = __longlong_as_double((long long)atomicExch(&ctl->maxdiff_bits, 0ull));
        ctl->last_maxdiff = md;
        ctl->done = 0;
        ctl->sweeps_done = sweeps_done + 1;
        if (md < ctl->crit) {  // double < float, as belief_propagation.cpp:406
            ctl->converged = 1;
            ctl->niter = int(sweeps_done - ctl->sweep_base);
        }
    }
}

}  // namespace sbmbp
