// Device-side data structures and small helpers shared by the sm_100a kernels of the BP engine.
//
// Layout in HBM (one GPU):
//   row_ptr  u64[N+1]      CSR row offsets, rows = nodes, neighbours ascending
//   S[2]     T[M*Q]        message buffers (double-buffered).  The message OUT of node i along its slot
//                          e = (i, l) lives at S[pos[e]], the message INTO i along e at S[rev[e]] (the
//                          reference's mmap_[i][l], belief_propagation.h:65-66).  pos orders the buffer by
//                          (destination bucket, source slot): B200 fetches a full 128-byte line from HBM for any
//                          random access (measured, tools/gather_bench.cu), so the in-messages of each bucket of
//                          consecutive nodes form one contiguous L2-sized region -- a line is fetched once and
//                          serves all the gathers that land in it -- while the writes of consecutive tiles
//                          advance sequentially inside every region.  One bucket <=> pos = identity.
//   rev, pos u32[M]        see above
//   marg     f64[N*Q]      marginals real_psi_ (always double: h and the free energy derive from them)
//   tiles    Tile[ntiles]  node-aligned work tiles (<= TE edges, <= TN nodes) or single hub nodes
//   field[2] Field         h_q and exp(-beta h_q / N), double-buffered like S
//   ctl      Ctl           device-resident sweep control: sweep counter (buffer parity), convergence flag
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

namespace sbmbp {

constexpr int kMaxQ = 32;
constexpr int kThreads = 256;
#ifndef SBMBP_MINB
#define SBMBP_MINB 1
#endif  // minimum resident CTAs per SM the fast sweep kernel is compiled for (caps registers)
constexpr unsigned kLargeDegree = 50;  // belief_propagation.h:68: nodes of degree >= 50 update in log domain
constexpr double kEps = 1.0e-50;       // belief_propagation.h:69
constexpr int kEnergyHead = 4;  // edge-pass result row: f_site, f_edge, entropy_site, entropy_edge; then QT*QT cab sums

// Loops over the Q components are fully unrolled (register-resident Q-vectors) up to QT = 8; the generic
// large-Q instantiations keep them rolled (local-memory vectors) -- they are correctness paths for now.
#define SBMBP_UNROLL_Q _Pragma("unroll (QT <= 8 ? QT : 1)")

struct Tile {
    unsigned long long e0;  // first edge slot
    unsigned n0;            // first node
    unsigned nn;            // node count; a hub tile has nn == 1 and more than TE edges
    unsigned ne;            // edge count (row_ptr[n0 + nn] - e0), so a CTA knows its extent from one 32-byte load
    unsigned nbig;          // nodes of degree >= 32 in the tile (they get a warp each; usually none)
};

// model parameters as the kernels consume them; [t * kMaxQ + q]
struct DevParams {
    double Ks[kMaxQ * kMaxQ];  // kernel of the degree<50 path: dc0 pow(c_tq, beta) (belief_propagation.cpp:1004); dc1 c_tq
    double Kl[kMaxQ * kMaxQ];  // kernel of the degree>=50 path: c_tq, beta ignored (:835)
    double P[kMaxQ * kMaxQ];   // p_tq = c_tq / N (dc2: tau = d_i d_l p_tq, :1009)
    double C[kMaxQ * kMaxQ];   // c_tq (h-field, :342; EM statistics, :911)
    double W[kMaxQ * kMaxQ];   // (1 - c/N)^beta: pair weight of the non-edge term (:689)
    double W1[kMaxQ * kMaxQ];  // 1 - W, exactly (series form of the same term: it expands the reference's rounded weight)
    double EA[kMaxQ * kMaxQ];  // (c/N) log c: numerator weight of the non-edge entropy term (:724)
    double EW[kMaxQ * kMaxQ];  // 1 - c/N: its denominator weight (:722; beta does not enter the entropy)
    double eta[kMaxQ];
    double logeta[kMaxQ];
    double beta;
    double N;
};

struct Field {
    double h[kMaxQ];     // h_q = sum_i w_i sum_t c_tq psi_i^t  (init_h / update_h, :320-360)
    double exph[kMaxQ];  // exp(-beta h_q / N)                   (update_exph_with_h, :363-368)
    double wsum[kMaxQ];  // sum_i w_i psi_i^t, kept for inspection
};

struct Ctl {
    unsigned long long maxdiff_bits;  // running max of |old - new| this sweep (non-negative doubles order like u64)
    unsigned done;                    // CTAs finished this sweep
    unsigned sweeps_done;             // completed sweeps; parity selects the source buffer
    unsigned max_sweeps;              // stop launching work at this count
    int converged;                    // set by the last CTA of a sweep once maxdiff < crit
    int niter;                        // sweep index (0-based, within the current converge call) at convergence
    unsigned sweep_base;              // sweeps_done when the current converge call started
    float crit;                       // compared as belief_propagation.cpp:406 does (double < float)
    double last_maxdiff;
    unsigned long long nan_count;     // messages that came out non-finite (diagnostic)
    unsigned long long tiny_count;    // edge updates with a b_l[q] < 1e-50 (where the reference's :1029-1042 fallback differs)
};

// ---------------------------------------------------------------- small helpers

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reciprocal to working precision without the IEEE division sequence (~30 instructions in FP64): hardware seed plus
// two Newton steps (FP64; relative error ~2 ulp), the hardware approximation itself in FP32 (1 ulp).  x = 0, a
// subnormal or a non-finite x give a non-finite result, which the caller reports through the max-diff.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// CTA-wide sum of one double per thread; result valid in every thread.  scratch: kThreads/32 doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) r += scratch[i];
    return r;
}
__device__ __forceinline__ double block_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = scratch[0];
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) r = fmax(r, scratch[i]);
    return r;
}

// 16-byte global load flavours for the random message gather (selected at run time while tuning)
__device__ __forceinline__ uint4 ld_gather16(const uint4 *p, int mode) {
    uint4 r;
    switch (mode) {
        case 1: return __ldcg(p);  // ld.global.cg: cache in L2 only
        case 2: return __ldcs(p);  // ld.global.cs: streaming, evict first
        case 3:
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
            return r;
        case 4:
            asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
            return r;
        case 5: return *p;  // plain ld.global (L1-allocating, coherent)
        case 6: return __ldcv(p);  // ld.global.cv: do not cache
        default: return __ldg(p);  // ld.global.nc
    }
}

// Q-vector of messages moved with the widest aligned access the element count allows.
template <typename T, int QT>
struct MsgVec {
    T v[QT];

    __device__ __forceinline__ void load(const T *__restrict__ p, unsigned Q) {
        if (Q == QT) {
            constexpr int bytes = QT * int(sizeof(T));
            if constexpr (bytes % 16 == 0) {
                const uint4 *s = reinterpret_cast<const uint4 *>(p);
                uint4 *d = reinterpret_cast<uint4 *>(v);
#pragma unroll
                for (int i = 0; i < bytes / 16; ++i) d[i] = __ldg(s + i);
            } else if constexpr (bytes == 8) {
                *reinterpret_cast<uint2 *>(v) = __ldg(reinterpret_cast<const uint2 *>(p));
            } else {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) v[q] = __ldg(p + q);
            }
        } else {
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q) v[q] = (unsigned(q) < Q) ? __ldg(p + q) : T(0);
        }
    }

    // gather flavour of load(): only the 16-byte-multiple fast path honours `mode`
    __device__ __forceinline__ void gather(const T *__restrict__ p, unsigned Q, int mode) {
        constexpr int bytes = QT * int(sizeof(T));
        if (Q == QT && bytes % 16 == 0) {
            const uint4 *s = reinterpret_cast<const uint4 *>(p);
            uint4 *d = reinterpret_cast<uint4 *>(v);
#pragma unroll
            for (int i = 0; i < bytes / 16; ++i) d[i] = ld_gather16(s + i, mode);
        } else {
            load(p, Q);
        }
    }

    __device__ __forceinline__ void store(T *__restrict__ p, unsigned Q) const {
        if (Q == QT) {
            constexpr int bytes = QT * int(sizeof(T));
            if constexpr (bytes % 16 == 0) {
                uint4 *d = reinterpret_cast<uint4 *>(p);
                const uint4 *s = reinterpret_cast<const uint4 *>(v);
#pragma unroll
                for (int i = 0; i < bytes / 16; ++i) d[i] = s[i];
            } else if constexpr (bytes == 8) {
                *reinterpret_cast<uint2 *>(p) = *reinterpret_cast<const uint2 *>(v);
            } else {
SBMBP_UNROLL_Q
                for (int q = 0; q < QT; ++q) p[q] = v[q];
            }
        } else {
SBMBP_UNROLL_Q
            for (int q = 0; q < QT; ++q)
                if (unsigned(q) < Q) p[q] = v[q];
        }
    }
};

// Scalar type the Q x Q kernel matrix is held in by the tile kernels: T, except for FP32 storage with a long contraction
// (Q >= 8), where it stays double -- a float c_ab is off by up to 6e-8, the SAME way on every edge, and a degree-d
// log-domain node raises that to the power d (1.4e-5 at d = 400: above FP32 mode's 1e-5 bar).
template <typename T, int QT>
using KernT = typename std::conditional<(sizeof(T) == 4 && QT >= 8), double, T>::type;

// tile geometry per instantiation: TE edges and TN nodes per CTA tile.  Sized so that several CTAs share an SM.
template <typename T, int QT>
struct TileCfg {
#ifndef SBMBP_TE_Q2
#define SBMBP_TE_Q2 1024
#endif
    static constexpr int TE = (QT <= 2) ? SBMBP_TE_Q2 : (QT <= 4) ? 512 : 256;
    static constexpr int TN = TE / 2;
};

// ---- warp tiles: the unit of work of bp_sweep_warp_kernel (sweep_warp.cuh)
struct WTile {         // 16 bytes: one warp's unit of work
    unsigned e0lo, e0hi;  // first edge slot
    unsigned n0;          // first node
    unsigned packed;      // ne (16 bits) | nn << 16 (8 bits) | kind << 24
    __host__ __device__ unsigned long long e0() const { return (static_cast<unsigned long long>(e0hi) << 32) | e0lo; }
    __host__ __device__ unsigned ne() const { return packed & 0xffffu; }
    __host__ __device__ unsigned nn() const { return (packed >> 16) & 0xffu; }
    __host__ __device__ unsigned kind() const { return packed >> 24; }
};

template <typename T, int QT>
struct WarpCfg {
    static constexpr int EPL = (QT * int(sizeof(T)) <= 16) ? 4 : 2;  // edge slots per lane
    static constexpr int WE = 32 * EPL;                              // edge slots per warp tile (<= 128: info has 7 slot bits)
};
constexpr unsigned kWInfoSlotMask = 127u;  // info word: slot in bits 0-6, node in bits 7-11
constexpr unsigned kWInfoNodeShift = 7u;

// ---- degree classes: the unit of layout of bp_sweep_ell_kernel (sweep_ell.cuh)
constexpr unsigned kEllDegrees = 32;      // degrees 0..31; higher degrees go to the warp / hub kernels
struct EllClass {
    unsigned d;            // degree of every node in the class
    unsigned n;            // nodes in the class
    unsigned node_first;   // first entry of the class in ell_node[]
    unsigned chunk_first;  // first 32-node chunk of the class
    unsigned base;         // index-array offset of the first chunk; chunk k, slot l, lane r: base + 32 d k + 32 l + r
    unsigned pad[3];
};

// dynamic shared memory carve-up of the tile kernels (sweep and energy share it)
template <typename T, int QT>
struct TileSmem {
    using Cfg = TileCfg<T, QT>;
    static constexpr size_t off_num = 0;                                               // double[QT*TN]
    static constexpr size_t off_red = off_num + sizeof(double) * QT * Cfg::TN;          // double[(kThreads/32)*(QT+2)]
    static constexpr size_t off_par = off_red + sizeof(double) * (kThreads / 32) * (QT + 2);  // double[5*QT]: eta, logeta, h, exph, spare
    static constexpr size_t off_ks = off_par + sizeof(double) * 5 * QT;                 // T[QT*QT]
    static constexpr size_t off_kl = off_ks + sizeof(KernT<T, QT>) * QT * QT;           // KernT[QT*QT]
    static constexpr size_t off_p = off_kl + sizeof(KernT<T, QT>) * QT * QT;                       // double[QT*QT] (dc2 only, but always carved)
    static constexpr size_t off_b = off_p + sizeof(double) * QT * QT;                   // T[QT*TE]
    static constexpr size_t off_off = off_b + sizeof(T) * QT * Cfg::TE;                 // u32[TN+1] (+pad)
    static constexpr size_t off_node = off_off + sizeof(unsigned) * (Cfg::TN + 4);      // u16[TE]
    static constexpr size_t bytes = off_node + sizeof(unsigned short) * Cfg::TE;
};

}  // namespace sbmbp
