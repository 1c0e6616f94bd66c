// Reference-exact asynchronous replay (SBMBP_SCHED_REPLAY): the schedule of converge() itself
// (belief_propagation.cpp:392-405) -- N draws WITH replacement per sweep from the caller's std::mt19937, every draw
// updating one node in place (Gauss-Seidel) with the field h maintained incrementally around it (:1088-1095).
//
// Every update reads the h the previous one left, so the chain is serial by construction; what is parallel is the
// inside of one update.  One warp walks the draws of a sweep: lanes take the node's in-edges (the contraction with
// c_ab, the leave-one-out division, the scatter through the reverse index), one lane takes the order-dependent
// scalars.  Sums and products run in the reference's own order with separately rounded multiplies and adds (the
// reference is built for baseline x86-64, no FMA), so a trajectory differs from the CPU's only through the last bit of
// exp() / log().  Purpose: sweep counts (`niter`) and trajectories that can be laid next to the reference's, on graphs
// up to ~10^5 nodes -- about 2 us per draw, a checking path, not a throughput path.
//
// State is held in the reference's own order (msg[(row_ptr[i] + l) * Q + q] == mmap_[i][l][q]) in double precision:
// the engine exports its buffers into that order before the first draw and imports them afterwards.
#pragma once
#include "bp_device.cuh"

namespace sbmbp {

struct ReplayArgs {
    const unsigned long long *row_ptr;  // [N + 1]
    const unsigned *rev;                // [M] reference slot of the reverse edge: where mmap_[i2][l2] lives (:1057-1058)
    const unsigned *degn;               // [M] degree of the neighbour behind each slot (dc != 0), else nullptr
    const int *clamp;                   // conf_planted_ when bp_conditional applies (:1104), else nullptr
    const unsigned *sched;              // the draws of this launch: i = unsigned(int(U * N)) (:395)
    unsigned count;
    double *msg;   // [M * Q] updated in place
    double *marg;  // [N * Q] real_psi_
    double *h;     // [Q] the field, carried from launch to launch
    const DevParams *prm;
    // per-edge scratch of one update, each [max_degree]: b of the current component, field_iter_ (kept across
    // updates like the reference's member, :1015-1016), _mmap_total_, maxpom_psii_iter_; then _mmap_q_nb_ [Q][max_degree]
    double *scratch;
    unsigned max_degree;
    unsigned N, Q, dc;
    double damping;
    double *out;     // [0]: maxdiffm of the sweep (:393-404)
    int init_field;  // recompute h from the marginals first (init_h, :320-332): the first sweep of every converge()
};

namespace replay {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
}  // namespace replay

__global__ void __launch_bounds__(32) bp_replay_kernel(const ReplayArgs a) {
    using namespace replay;
    __shared__ double s_K[kMaxQ * kMaxQ];  // degree < 50, dc 0: pow(c_tq, beta) (:1004), evaluated by the host's libm
    __shared__ double s_C[kMaxQ * kMaxQ];  // c_tq
    __shared__ double s_P[kMaxQ * kMaxQ];  // p_tq (dc 2)
    __shared__ double s_eta[kMaxQ], s_logeta[kMaxQ], s_h[kMaxQ], s_exph[kMaxQ], s_psi[kMaxQ];
    const unsigned lane = threadIdx.x;
    const unsigned Q = a.Q, dc = a.dc, md = a.max_degree ? a.max_degree : 1u;
    const double Nd = double(a.N), beta = a.prm->beta;
    for (unsigned i = lane; i < kMaxQ * kMaxQ; i += 32) {
        s_K[i] = a.prm->Ks[i];
        s_C[i] = a.prm->C[i];
        s_P[i] = a.prm->P[i];
    }
    if (lane < Q) {
        s_eta[lane] = a.prm->eta[lane];
        s_logeta[lane] = a.prm->logeta[lane];
        s_h[lane] = a.h[lane];
    }
    __syncwarp();
    double *bcur = a.scratch, *field = bcur + md, *tot = field + md, *maxp = tot + md, *nb = maxp + md;

    // h_[q1] +/-= w * cab_[q2][q1] * real_psi_[i][q2], q2 ascending (update_h, :334-360); lane q1 owns h_[q1]
    auto update_h = [&](unsigned i, double di, bool plus) {
        if (lane < Q) {
            double h = s_h[lane];
            for (unsigned q2 = 0; q2 < Q; ++q2) {
                const double c = s_C[q2 * kMaxQ + lane], p = a.marg[size_t(i) * Q + q2];
                const double term = (dc == 0) ? mul(c, p) : mul(mul(di, c), p);
                h = plus ? add(h, term) : sub(h, term);
            }
            s_h[lane] = h;
        }
        __syncwarp();
    };
    auto update_exph = [&]() {  // :363-368
        if (lane < Q) s_exph[lane] = exp(dvd(mul(-beta, s_h[lane]), Nd));
        __syncwarp();
    };
    // b = sum_t K(t, q) * mmap_[i][l][t], t ascending; `large`: the >= 50 routine, which has no beta (:835)
    auto contract = [&](const double *m, unsigned q, double di, double dn, bool large) {
        double b = 0.0;
        for (unsigned t = 0; t < Q; ++t) {
            double k;
            if (dc == 0) k = large ? s_C[t * kMaxQ + q] : s_K[t * kMaxQ + q];
            else if (dc == 1) k = mul(mul(di, dn), s_C[t * kMaxQ + q]);
            else {
                const double tmp = mul(mul(di, dn), s_P[t * kMaxQ + q]);
                k = dvd(tmp, add(1.0, tmp));
            }
            b = add(b, mul(k, m[t]));
        }
        return b;
    };

    if (a.init_field) {  // init_h: h = 0, then update_h(i, +1) for i ascending
        if (lane < Q) s_h[lane] = 0.0;
        __syncwarp();
        if (lane < Q) {
            double h = 0.0;
            for (unsigned i = 0; i < a.N; ++i) {
                const double di = double(unsigned(a.row_ptr[i + 1] - a.row_ptr[i]));
                for (unsigned q2 = 0; q2 < Q; ++q2) {
                    const double c = s_C[q2 * kMaxQ + lane], p = a.marg[size_t(i) * Q + q2];
                    h = add(h, (dc == 0) ? mul(c, p) : mul(mul(di, c), p));
                }
            }
            s_h[lane] = h;
        }
        __syncwarp();
    }
    update_exph();

    double maxdiffm = -100.0;
    unsigned inext = a.count ? a.sched[0] : 0u;
    for (unsigned k = 0; k < a.count; ++k) {
        const unsigned i = inext;
        if (k + 1 < a.count) inext = a.sched[k + 1];
        const unsigned long long e0 = a.row_ptr[i];
        const unsigned d = unsigned(a.row_ptr[i + 1] - e0);
        const double di = double(d);
        double diffm;
        if (d >= kLargeDegree) {
            // ---- bp_iter_update_psi_large_degree (:813-890): log domain, ignores beta and conf_planted_
            for (unsigned l = lane; l < d; l += 32) {
                tot[l] = 0.0;
                maxp[l] = -100000000.0;
            }
            double maxpom = -100000000.0;
            for (unsigned q = 0; q < Q; ++q) {
                for (unsigned l = lane; l < d; l += 32) {
                    const double dn = a.degn ? double(a.degn[e0 + l]) : 0.0;
                    const double tmp = log(contract(a.msg + (e0 + l) * Q, q, di, dn, true));
                    bcur[l] = tmp;
                    field[l] = tmp;
                }
                __syncwarp();
                double psi = 0.0;
                if (lane == 0) {
                    double acc = 0.0;
                    for (unsigned l = 0; l < d; ++l) acc = add(acc, bcur[l]);
                    const double hterm = (dc == 0) ? dvd(s_h[q], Nd) : dvd(mul(mul(1.0, di), s_h[q]), Nd);
                    psi = sub(add(acc, s_logeta[q]), hterm);
                    s_psi[q] = psi;
                }
                psi = __shfl_sync(0xffffffffu, psi, 0);
                if (psi > maxpom) maxpom = psi;
                for (unsigned l = lane; l < d; l += 32) {
                    const double v = sub(psi, field[l]);
                    nb[size_t(q) * md + l] = v;
                    if (v > maxp[l]) maxp[l] = v;
                }
                __syncwarp();
            }
            double total = 0.0;
            for (unsigned q = 0; q < Q; ++q) {
                total = add(total, exp(sub(s_psi[q], maxpom)));
                for (unsigned l = lane; l < d; l += 32) tot[l] = add(tot[l], exp(sub(nb[size_t(q) * md + l], maxp[l])));
            }
            __syncwarp();
            update_h(i, di, false);
            double mymax = -100.0;
            for (unsigned q = 0; q < Q; ++q) {
                if (lane == 0) a.marg[size_t(i) * Q + q] = dvd(exp(sub(s_psi[q], maxpom)), total);
                for (unsigned l = lane; l < d; l += 32) {
                    double *slot = a.msg + size_t(a.rev[e0 + l]) * Q + q;
                    const double thisvalue = dvd(exp(sub(nb[size_t(q) * md + l], maxp[l])), tot[l]);
                    const double old = *slot;
                    const double df = fabs(sub(old, thisvalue));
                    if (df > mymax) mymax = df;
                    *slot = add(mul(a.damping, thisvalue), mul(sub(1.0, a.damping), old));
                }
            }
            __syncwarp();
            update_h(i, di, true);
            update_exph();
            diffm = mymax;
        } else if (a.clamp && a.clamp[i] != -1) {
            diffm = 0.0;  // bp_conditional: a planted node keeps its messages (:1113-1123)
        } else {
            // ---- sum_all_messages_to_i (:991-1049)
            for (unsigned l = lane; l < d; l += 32) tot[l] = 0.0;  // clean_mmap_total_at_node_i_ (:422-426)
            double total = 0.0;
            for (unsigned q = 0; q < Q; ++q) {
                for (unsigned l = lane; l < d; l += 32) {
                    const double dn = a.degn ? double(a.degn[e0 + l]) : 0.0;
                    const double b = contract(a.msg + (e0 + l) * Q, q, di, dn, false);
                    bcur[l] = b;
                    if (b != 0.) field[l] = b;  // b == 0: `continue` leaves field_iter_[l] as it was (:1011-1016)
                }
                __syncwarp();
                double psi = 0.0;
                if (lane == 0) {
                    double acc = 1.0;
                    for (unsigned l = 0; l < d; ++l) {
                        const double b = bcur[l];
                        if (b != 0.) acc = mul(acc, b);
                    }
                    const double F = (dc == 0) ? s_exph[q] : exp(dvd(mul(mul(-1.0, di), s_h[q]), Nd));
                    psi = mul(mul(acc, s_eta[q]), F);
                    s_psi[q] = psi;
                }
                psi = __shfl_sync(0xffffffffu, psi, 0);
                total = add(total, psi);
                for (unsigned l = lane; l < d; l += 32) {
                    const double f = field[l];
                    double v;
                    if (f < kEps) {  // :1029-1042
                        v = 1.0;
                        for (unsigned lx = 0; lx < d; ++lx) {
                            if (lx == l) continue;
                            const double fx = field[lx];
                            if (fx != 0) v = mul(v, fx);
                        }
                    } else {
                        v = dvd(psi, f);
                    }
                    nb[size_t(q) * md + l] = v;
                    tot[l] = add(tot[l], v);
                }
                __syncwarp();
            }
            update_h(i, di, false);
            // ---- norm_m_at_i (:1051-1071)
            double mymax = -100.0;
            for (unsigned q = 0; q < Q; ++q) {
                if (lane == 0) a.marg[size_t(i) * Q + q] = dvd(s_psi[q], total);
                for (unsigned l = lane; l < d; l += 32) {
                    double *slot = a.msg + size_t(a.rev[e0 + l]) * Q + q;
                    const double v = nb[size_t(q) * md + l], t = tot[l];
                    const double old = *slot;
                    const double df = fabs(sub(old, dvd(v, t)));
                    if (df > mymax) mymax = df;
                    *slot = add(dvd(mul(a.damping, v), t), mul(sub(1.0, a.damping), old));
                }
            }
            __syncwarp();
            update_h(i, di, true);
            update_exph();
            diffm = mymax;
        }
        // the node's diff is the max over its (q, l); NaN compares false everywhere, as on the CPU
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, diffm, o);
            if (other > diffm) diffm = other;
        }
        if (diffm > maxdiffm) maxdiffm = diffm;
    }
    if (lane < Q) a.h[lane] = s_h[lane];
    if (lane == 0) a.out[0] = maxdiffm;
}

}  // namespace sbmbp
