// Small state kernels around the sweep: field initialisation (init_h), state import/export in the
// reference's message order, device-side random initialisation, neighbour-degree table.
#pragma once
#include "bp_device.cuh"
#include "sweep_kernel.cuh"

namespace sbmbp {

// ---- init_h (belief_propagation.cpp:320-332): wsum_t = sum_i w_i psi_i^t, w_i = 1 (dc 0) or d_i (dc 1, 2)
static __global__ void __launch_bounds__(kThreads) field_partial_kernel(const double *__restrict__ marg,
                                                                 const unsigned long long *__restrict__ row_ptr,
                                                                 unsigned N, unsigned Q, unsigned dc,
                                                                 double *__restrict__ partial) {
    __shared__ double sred[kThreads / 32];
    double acc[kMaxQ];
    for (unsigned q = 0; q < Q; ++q) acc[q] = 0.0;
    for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < N; i += gridDim.x * kThreads) {
        const double w = (dc == 0) ? 1.0 : double(row_ptr[i + 1] - row_ptr[i]);
        for (unsigned q = 0; q < Q; ++q) acc[q] += w * marg[size_t(i) * Q + q];
    }
    for (unsigned q = 0; q < Q; ++q) {
        const double v = block_sum(acc[q], sred);
        if (threadIdx.x == 0) partial[size_t(blockIdx.x) * kMaxQ + q] = v;
    }
}

// fixed-order final reduction; writes the field the next sweep will read (parity of ctl->sweeps_done)
static __global__ void __launch_bounds__(kThreads) field_final_kernel(const double *__restrict__ partial, unsigned nblocks,
                                                               const DevParams *prm, unsigned Q, Field *f0,
                                                               Field *f1, const Ctl *ctl) {
    __shared__ double sred[kThreads / 32];
    __shared__ double tot[kMaxQ];
    for (unsigned q = 0; q < Q; ++q) {
        double p = 0.0;
        for (unsigned b = threadIdx.x; b < nblocks; b += kThreads) p += partial[size_t(b) * kMaxQ + q];
        const double v = block_sum(p, sred);
        if (threadIdx.x == 0) tot[q] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) publish_field(prm, Q, tot, (ctl->sweeps_done & 1u) ? f1 : f0);
}

// ---- arms the device-resident sweep control for the next `add` sweeps (replaces a host -> device copy of Ctl, so
// launching sweeps needs no host synchronisation): clears the convergence flag, re-bases the sweep index.
static __global__ void ctl_arm_kernel(Ctl *ctl, float crit, unsigned add) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        ctl->converged = 0;
        ctl->niter = -1;
        ctl->sweep_base = ctl->sweeps_done;
        ctl->max_sweeps = ctl->sweeps_done + add;
        ctl->crit = crit;
        ctl->maxdiff_bits = 0ull;
        ctl->done = 0u;
    }
}

// ---- multi-GPU pieces ------------------------------------------------------------------------------------------
struct PeerTable {
    void *p[8];
};

// outbox entry of every remote out-message <- the value it currently has at its owner (one-off, after a state was set).
// word[t]: the kernels' pos word (bit 31: outbox index); rpos[t]: owner << 29 | position at the owner
template <typename T>
__global__ void mirror_pull_kernel(T *__restrict__ mirror, const unsigned *__restrict__ word, const unsigned *__restrict__ rpos,
                                   PeerTable peers, unsigned long long M, unsigned Q) {
    const unsigned long long total = M * Q;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long t = idx / Q;
        const unsigned q = unsigned(idx - t * Q);
        const unsigned w = word[t];
        if (!(w & 0x80000000u)) continue;
        const unsigned p = rpos[t];
        const T *src = static_cast<const T *>(peers.p[p >> 29]);
        mirror[(unsigned long long)(w & 0x7fffffffu) * Q + q] = src[(unsigned long long)(p & ((1u << 29) - 1u)) * Q + q];
    }
}

// out[c] = fixed-order reduction of the per-tile rows (sum for c < QT, max for c == QT): the rank's contribution
template <int QT>
__global__ void __launch_bounds__(kFinalThreads) bp_reduce_rows_kernel(const double *__restrict__ partial,
                                                                       unsigned ntiles, double *__restrict__ out) {
    constexpr int NC = QT + 1;
    __shared__ double sred[kFinalThreads / 32][NC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc[NC];
SBMBP_UNROLL_Q
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (unsigned b = tid; b < ntiles; b += kFinalThreads) {
        const double *row = partial + size_t(b) * NC;
SBMBP_UNROLL_Q
        for (int c = 0; c < NC; ++c) acc[c] = (c < QT) ? acc[c] + row[c] : fmax(acc[c], row[c]);
    }
SBMBP_UNROLL_Q
    for (int c = 0; c < NC; ++c) {
        const double v = (c < QT) ? warp_sum(acc[c]) : warp_max(acc[c]);
        if (lane == 0) sred[warp][c] = v;
    }
    __syncthreads();
    if (tid < NC) {
        double r = sred[0][tid];
        for (int w = 1; w < kFinalThreads / 32; ++w) r = (tid < QT) ? r + sred[w][tid] : fmax(r, sred[w][tid]);
        out[tid] = r;
    }
}

// closes a sweep (advance = 1) or an init_h (advance = 0) from the all-gathered per-rank rows [world][stride]:
// ranks are combined in rank order on every GPU, so all ranks publish bit-identical fields and decisions
static __global__ void bp_finalize_dist_kernel(const double *__restrict__ gathered, unsigned world, unsigned stride,
                                               unsigned Q, const DevParams *prm, Field *f0, Field *f1, Ctl *ctl,
                                               int advance) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned sweeps_done = ctl->sweeps_done;
    if (advance && (ctl->converged || sweeps_done >= ctl->max_sweeps)) return;
    double tot[kMaxQ + 1];
    for (unsigned c = 0; c <= Q; ++c) {
        const unsigned col = (c < Q) ? c : stride - 1;
        double r = 0.0;
        for (unsigned k = 0; k < world; ++k) r = (c < Q) ? r + gathered[k * stride + col] : fmax(r, gathered[k * stride + col]);
        tot[c] = r;
    }
    if (advance) {
        publish_field(prm, Q, tot, (sweeps_done & 1u) ? f0 : f1);
        const double md = tot[Q];
        ctl->last_maxdiff = md;
        ctl->sweeps_done = sweeps_done + 1;
        if (!(md == md) || md > 1.0e299) ctl->nan_count += 1;
        if (md < ctl->crit) {
            ctl->converged = 1;
            ctl->niter = int(sweeps_done - ctl->sweep_base);
        }
    } else {
        publish_field(prm, Q, tot, (sweeps_done & 1u) ? f1 : f0);
    }
}

// ---- state import / export.  ref[(row_ptr[i]+l)*Q+q] = mmap_[i][l][q] = message INTO i along slot e,
// which the engine keeps at S[rev[e]] (rev here is the engine's gather index, a bijection of the slots).
template <typename T>
__global__ void import_msgs_kernel(const double *__restrict__ ref, const unsigned *__restrict__ rev,
                                   T *__restrict__ S, unsigned long long M, unsigned Q) {
    const unsigned long long total = M * Q;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = idx / Q;
        const unsigned q = unsigned(idx - e * Q);
        S[(unsigned long long)rev[e] * Q + q] = T(ref[idx]);
    }
}

template <typename T>
__global__ void export_msgs_kernel(const T *__restrict__ S, const unsigned *__restrict__ rev,
                                   double *__restrict__ ref, unsigned long long M, unsigned Q) {
    const unsigned long long total = M * Q;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = idx / Q;
        const unsigned q = unsigned(idx - e * Q);
        ref[idx] = double(S[(unsigned long long)rev[e] * Q + q]);
    }
}

// ---- device-side random initialisation: same distribution as init_messages flag 0
// (belief_propagation.cpp:110-131: Q iid uniforms, normalised), from a counter-based generator.
__device__ __forceinline__ double unit_uniform(unsigned long long seed, unsigned long long ctr) {
    unsigned long long z = seed + 0x9e3779b97f4a7c15ull * (ctr + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return (double(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

template <typename T>
__global__ void random_init_kernel(T *__restrict__ S, double *__restrict__ marg, unsigned long long M, unsigned N,
                                   unsigned Q, unsigned long long seed) {
    const unsigned long long total = M + N;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        double v[kMaxQ], norm = 0.0;
        for (unsigned q = 0; q < Q; ++q) {
            v[q] = unit_uniform(seed, idx * Q + q);
            norm += v[q];
        }
        if (idx < N) {
            for (unsigned q = 0; q < Q; ++q) marg[idx * Q + q] = v[q] / norm;
        } else {
            const unsigned long long e = idx - N;
            for (unsigned q = 0; q < Q; ++q) S[e * Q + q] = T(v[q] / norm);
        }
    }
}

// marg[node] = marg_ell[k] for every real entry of the chunk-ordered marginal array (padding entries hold ~0u)
static __global__ void ell_scatter_marg_kernel(const double *__restrict__ marg_ell, const unsigned *__restrict__ ell_node,
                                                unsigned n_entries, unsigned Q, double *__restrict__ marg) {
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n_entries; k += gridDim.x * blockDim.x) {
        const unsigned node = ell_node[k];
        if (node == 0xffffffffu) continue;
        for (unsigned q = 0; q < Q; ++q) marg[size_t(node) * Q + q] = marg_ell[size_t(k) * Q + q];
    }
}

// ---- compact storage of normalised Q = 2 FP64 messages (sweep_ell.cuh): one double per message.
// full -> compact; bad counts the messages that are not normalised to rounding (|psi_0 + psi_1 - 1| > 1e-15 or a
// negative component): with any such message the engine stays on full storage.
static __global__ void compact_pack_kernel(const double *__restrict__ S, double *__restrict__ C, unsigned long long slots,
                                           unsigned long long *__restrict__ bad) {
    unsigned long long mybad = 0;
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < slots;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        const double2 v = reinterpret_cast<const double2 *>(S)[p];
        if (!(fabs((v.x + v.y) - 1.0) <= 1.0e-15) || !(v.x >= 0.0) || !(v.y >= 0.0)) ++mybad;
        C[p] = (v.x <= v.y) ? v.x : -v.y;
    }
    if (mybad) atomicAdd(bad, mybad);
}
static __global__ void compact_unpack_kernel(const double *__restrict__ C, double *__restrict__ S, unsigned long long slots) {
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < slots;
         p += (unsigned long long)gridDim.x * blockDim.x) {
        const double x = C[p], s = fabs(x);
        const bool second = __double_as_longlong(x) < 0;
        reinterpret_cast<double2 *>(S)[p] = second ? make_double2(1.0 - s, s) : make_double2(s, 1.0 - s);
    }
}

// degsrc[e] = degree of col[e] (the d_l of the degree-corrected kernels, belief_propagation.cpp:1006,1009)
static __global__ void degsrc_kernel(const unsigned long long *__restrict__ row_ptr, const unsigned *__restrict__ col,
                              unsigned *__restrict__ degsrc, unsigned long long M) {
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < M;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned j = col[e];
        degsrc[e] = unsigned(row_ptr[j + 1] - row_ptr[j]);
    }
}

}  // namespace sbmbp
