// Reductions hanging off the sweep, fused into ONE extra pass over the edges (same tiles as the sweep):
//   f_site        compute_f_site        belief_propagation.cpp:442-504
//   f_edge        compute_f_edge        :562-612
//   cab_expect    compute_cab_expect    :892-966 (the un-scaled two-point sums; the scaling of :967-988 is host-side)
//   entropy_site  compute_entropy_site  :506-560   (dc == 0; its "smaller term" is identically 0 in the reference
//                                                   because the leave-one-out loop of :538-546 always meets log(0))
//   entropy_edge  compute_entropy_edge  :614-672
// plus the node-only reductions (na/nna expectations :428-440, overlap confusion matrix :798-808) and the
// non-edge free-energy term (:675-709) as an exact tiled N^2 kernel for small N and a moment series otherwise.
// Everything accumulates in double and reduces in a fixed order (per-thread -> warp -> CTA -> per-tile
// partial -> one final CTA), so results are bitwise reproducible.
#pragma once
#include "bp_device.cuh"

namespace sbmbp {


template <typename T>
struct EnergyArgs {
    const Tile *tiles;
    const unsigned long long *row_ptr;
    const unsigned *rev;
    const unsigned *pos;     // see SweepArgs: tile-sorted positions of the out-messages ...
    const unsigned *info;    // ... and the tile-local slot (low 16 bits) each one belongs to
    const T *mirror;         // multi-GPU: the outbox of this rank's remote out-messages (the owner holds the real copy); else nullptr
    const unsigned *degsrc;  // read when dc != 0
    const T *S;              // current messages
    const DevParams *prm;
    const double *Kmat;      // device pointer to the Q x Q kernel in use ([t*kMaxQ+q]): prm->Ks or prm->C
    const Field *field;
    double *partial;         // [ntiles][kEnergyHead + QT*QT]
    unsigned Q;
    unsigned dc;
    double fcoef;            // beta with Ks (f_site :474), 1 with C (entropy :551); dc != 0 uses d_i instead
};

template <typename T, int QT>
struct EnergySmem {
    using Cfg = TileCfg<T, QT>;
    static constexpr size_t off_red = 0;                                                         // double[8*(QT+2)]
    static constexpr size_t off_par = off_red + sizeof(double) * (kThreads / 32) * (QT + 2);     // double[3*QT]
    static constexpr size_t off_k = off_par + sizeof(double) * 3 * QT;                           // double[QT*QT] K
    static constexpr size_t off_kl = off_k + sizeof(double) * QT * QT;                           // double[QT*QT] K*log c / log c
    static constexpr size_t off_p = off_kl + sizeof(double) * QT * QT;                           // double[QT*QT] P
    static constexpr size_t off_cab = off_p + sizeof(double) * QT * QT;                          // double[8*QT*QT] warp-private
    static constexpr size_t off_b = off_cab + sizeof(double) * (QT > 8 ? (kThreads / 32) * QT * QT : 1);
    static constexpr size_t off_off = off_b + sizeof(double) * QT * Cfg::TE;                     // u32[TN+4]
    static constexpr size_t off_node = off_off + sizeof(unsigned) * (Cfg::TN + 4);               // u16[TE]
    static constexpr size_t bytes = off_node + sizeof(unsigned short) * Cfg::TE;
};

template <typename T, int QT>
__global__ void __launch_bounds__(kThreads) bp_energy_kernel(const EnergyArgs<T> a) {
    using Cfg = TileCfg<T, QT>;
    using Lay = EnergySmem<T, QT>;
    constexpr int TE = Cfg::TE;
    constexpr int NW = kThreads / 32;
    extern __shared__ __align__(16) unsigned char smem[];
    double *sred = reinterpret_cast<double *>(smem + Lay::off_red);
    double *slogeta = reinterpret_cast<double *>(smem + Lay::off_par);
    double *sh = slogeta + QT;
    double *sK = reinterpret_cast<double *>(smem + Lay::off_k);
    double *sKL = reinterpret_cast<double *>(smem + Lay::off_kl);
    double *sP = reinterpret_cast<double *>(smem + Lay::off_p);
    double *scab = reinterpret_cast<double *>(smem + Lay::off_cab);
    double *sb = reinterpret_cast<double *>(smem + Lay::off_b);
    unsigned *soff = reinterpret_cast<unsigned *>(smem + Lay::off_off);
    unsigned short *snode = reinterpret_cast<unsigned short *>(smem + Lay::off_node);

    const unsigned Q = a.Q, dc = a.dc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double Nd = a.prm->N;

    for (int i = tid; i < QT * QT; i += kThreads) {
        const int t = i / QT, q = i % QT;
        const bool in = unsigned(t) < Q && unsigned(q) < Q;
        const double c = in ? a.prm->C[t * kMaxQ + q] : 0.0;
        sK[i] = in ? a.Kmat[t * kMaxQ + q] : 0.0;
        // entropy-edge numerator weight: c log c (dc 0/1, :633,:636); log c alone for dc 2 (:642)
        sKL[i] = in ? ((dc == 2) ? log(c) : c * log(c)) : 0.0;
        sP[i] = in ? a.prm->P[t * kMaxQ + q] : 0.0;
    }
    if (QT > 8)
        for (int i = tid; i < NW * QT * QT; i += kThreads) scab[i] = 0.0;
    if (tid < QT) {
        slogeta[tid] = (unsigned(tid) < Q) ? a.prm->logeta[tid] : 0.0;
        sh[tid] = (unsigned(tid) < Q) ? a.field->h[tid] : 0.0;
    }

    const Tile tile = a.tiles[blockIdx.x];
    const unsigned long long e0 = tile.e0;
    const unsigned n0 = tile.n0, nn = tile.nn;
    const unsigned long long ne64 = tile.ne;
    const bool hub = ne64 > (unsigned long long)TE;

    double f_site = 0.0, f_edge = 0.0, ent_site = 0.0, ent_edge = 0.0;
    constexpr int NP = (QT <= 8) ? QT * (QT + 1) / 2 : 1;
    double cabreg[NP];
SBMBP_UNROLL_Q
    for (int i = 0; i < NP; ++i) cabreg[i] = 0.0;
    double hubacc[QT];
SBMBP_UNROLL_Q
    for (int q = 0; q < QT; ++q) hubacc[q] = 0.0;

    if (!hub) {
        for (unsigned n = tid; n <= nn; n += kThreads) soff[n] = unsigned(a.row_ptr[n0 + n] - e0);
        __syncthreads();
        for (unsigned n = tid; n < nn; n += kThreads)
            for (unsigned k = soff[n]; k < soff[n + 1]; ++k) snode[k] = (unsigned short)n;
    }
    __syncthreads();

    // ---- edge pass (regular tile: k < ne <= TE, hub: k strides over the whole row)
    const unsigned long long kmax = hub ? ((ne64 + kThreads - 1) / kThreads) * kThreads : ((ne64 + 31) / 32) * 32;
    for (unsigned long long t = tid; t < kmax; t += kThreads) {
        const bool live = t < ne64;
        // t runs in buffer order, k is the tile-local slot it belongs to (identity for hubs and unbucketed layouts)
        const unsigned long long k = (live && !hub) ? (unsigned long long)(__ldg(a.info + e0 + t) & 0xffffu) : t;
        double bin[QT], bout[QT], mi[QT], mo[QT];
        double norm = 1.0, scale = 1.0, didl = 1.0;
        if (live) {
            MsgVec<T, QT> m_in, m_out;
            m_in.load(a.S + size_t(__ldg(a.rev + e0 + k)) * Q, Q);
            // multi-GPU: a pos word with bit 31 set is an index into the outbox (remote destination), see dist_exchange.cuh
            const unsigned pw = __ldg(a.pos + e0 + t);
            if (a.mirror && (pw & 0x80000000u)) m_out.load(a.mirror + size_t(pw & 0x7fffffffu) * Q, Q);
            else m_out.load(a.S + size_t(pw) * Q, Q);
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                mi[q] = double(m_in.v[q]);
                mo[q] = double(m_out.v[q]);
            }
            if (dc != 0) {
                double di;
                if (hub) di = double(ne64);
                else {
                    const unsigned n = snode[k];
                    di = double(soff[n + 1] - soff[n]);
                }
                didl = di * double(__ldg(a.degsrc + e0 + k));
                if (dc == 1) scale = didl;
            }
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                double ai = 0.0, ao = 0.0;
SBMBP_UNROLL_Q
                for (int t = 0; t < QT; ++t) {
                    double kv = sK[t * QT + q];
                    if (dc == 2) {
                        const double tau = didl * sP[t * QT + q];
                        kv = (unsigned(t) < Q && unsigned(q) < Q) ? tau / (1.0 + tau) : 0.0;
                    }
                    ai += kv * mi[t];
                    ao += kv * mo[t];
                }
                bin[q] = ai * scale;
                bout[q] = ao * scale;
            }
            norm = 0.0;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) norm += mi[q] * bout[q];
            f_edge += log(norm);
            double numer = 0.0;
SBMBP_UNROLL_Q
            for (int q1 = 0; q1 < QT; ++q1)
SBMBP_UNROLL_Q
                for (int q2 = 0; q2 < QT; ++q2) {
                    double w = sKL[q1 * QT + q2];
                    if (dc == 2) {
                        const double tau = didl * sP[q1 * QT + q2];
                        w = (unsigned(q1) < Q && unsigned(q2) < Q) ? tau / (1.0 + tau) * w : 0.0;
                    }
                    numer += w * scale * mi[q1] * mo[q2];
                }
            ent_edge += numer / norm;
            if (hub) {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q)
                    if (unsigned(q) < Q) hubacc[q] += log(bin[q]);
            } else {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) sb[q * TE + k] = log(bin[q]);
            }
        }
        // two-point sums of :933-964, upper triangle
        if constexpr (QT <= 8) {
            if (live) {
                int idx = 0;
SBMBP_UNROLL_Q
                for (int q1 = 0; q1 < QT; ++q1)
SBMBP_UNROLL_Q
                    for (int q2 = q1; q2 < QT; ++q2) {
                        double kv = sK[q1 * QT + q2];
                        if (dc == 2) {
                            const double tau = didl * sP[q1 * QT + q2];
                            kv = (unsigned(q1) < Q && unsigned(q2) < Q) ? tau / (1.0 + tau) : 0.0;
                        }
                        const double pair = (q1 == q2) ? mi[q1] * mo[q2] : (mi[q1] * mo[q2] + mi[q2] * mo[q1]);
                        cabreg[idx++] += 0.5 * kv * scale * pair / norm;
                    }
            }
        } else {
            for (int q1 = 0; q1 < QT; ++q1)
                for (int q2 = q1; q2 < QT; ++q2) {
                    double term = 0.0;
                    if (live) {
                        double kv = sK[q1 * QT + q2];
                        if (dc == 2) {
                            const double tau = didl * sP[q1 * QT + q2];
                            kv = (unsigned(q1) < Q && unsigned(q2) < Q) ? tau / (1.0 + tau) : 0.0;
                        }
                        const double pair = (q1 == q2) ? mi[q1] * mo[q2] : (mi[q1] * mo[q2] + mi[q2] * mo[q1]);
                        term = 0.5 * kv * scale * pair / norm;
                    }
                    term = warp_sum(term);
                    if (lane == 0) scab[(warp * QT + q1) * QT + q2] += term;
                }
        }
    }
    __syncthreads();

    // ---- node pass: log-sum-exp of the node totals (:473-499) and the entropy site term (:551-556)
    if (!hub) {
        for (unsigned n = tid; n < nn; n += kThreads) {
            const unsigned k0 = soff[n], d = soff[n + 1] - k0;
            double v[QT], mx = -1.0e300;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) v[q] = 0.0;
            for (unsigned k = k0; k < k0 + d; ++k) {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) v[q] += sb[q * TE + k];
            }
            double es_num = 0.0, es_den = 0.0;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                if (unsigned(q) < Q) {
                    const double fieldterm = (dc == 0) ? a.fcoef * sh[q] / Nd : double(d) * sh[q] / Nd;
                    v[q] = v[q] + slogeta[q] - fieldterm;
                    mx = fmax(mx, v[q]);
                }
            }
            double s = 0.0;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) {
                if (unsigned(q) < Q) {
                    const double ex = exp(v[q] - mx);
                    s += ex;
                    es_den += ex;
                    es_num += ex * (-sh[q] / Nd);
                }
            }
            f_site += mx + log(s);
            ent_site += es_num / es_den;
        }
    } else {
        double v[QT], mx = -1.0e300;
SBMBP_UNROLL_Q
        for (int q = 0; q < QT; ++q) {
            v[q] = block_sum(hubacc[q], sred);
            if (unsigned(q) < Q) {
                const double fieldterm = (dc == 0) ? a.fcoef * sh[q] / Nd : double(ne64) * sh[q] / Nd;
                v[q] = v[q] + slogeta[q] - fieldterm;
                mx = fmax(mx, v[q]);
            }
        }
        if (tid == 0) {
            double s = 0.0, es_num = 0.0;
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q)
                if (unsigned(q) < Q) {
                    const double ex = exp(v[q] - mx);
                    s += ex;
                    es_num += ex * (-sh[q] / Nd);
                }
            f_site += mx + log(s);
            ent_site += es_num / s;
        }
    }

    // ---- CTA reduction -> per-tile partial row
    double *row = a.partial + size_t(blockIdx.x) * (kEnergyHead + QT * QT);
    double r;
    r = block_sum(f_site, sred);
    if (tid == 0) row[0] = r;
    r = block_sum(f_edge, sred);
    if (tid == 0) row[1] = r;
    r = block_sum(ent_site, sred);
    if (tid == 0) row[2] = r;
    r = block_sum(ent_edge, sred);
    if (tid == 0) row[3] = r;
    if constexpr (QT <= 8) {
        int idx = 0;
SBMBP_UNROLL_Q
        for (int q1 = 0; q1 < QT; ++q1)
SBMBP_UNROLL_Q
            for (int q2 = 0; q2 < QT; ++q2) {
                if (q2 >= q1) {
                    r = block_sum(cabreg[idx++], sred);
                    if (tid == 0) row[kEnergyHead + q1 * QT + q2] = r;
                } else if (tid == 0) {
                    row[kEnergyHead + q1 * QT + q2] = 0.0;
                }
            }
    } else {
        __syncthreads();
        for (int i = tid; i < QT * QT; i += kThreads) {
            double s = 0.0;
            for (int w = 0; w < NW; ++w) s += scab[w * QT * QT + i];
            row[kEnergyHead + i] = (i / QT <= i % QT) ? s : 0.0;
        }
    }
}

}  // namespace sbmbp
