// Multi-GPU halo exchange of the sweep kernels: everything a DIST sweep does besides the node updates.
//
// Rank p owns a contiguous node range and the buffers holding every message INTO its nodes.  A message i -> j whose
// destination lives on another rank is NOT stored into that rank's buffer one 16-byte st.global at a time (round 1:
// 0.32 TB/s of NVLink egress, the 8-GPU step at 0.215 of the roofline).  Instead:
//
//   * OUTBOX.  Remote out-messages are written into a local outbox (which doubles as the mirror holding their old
//     values for max-diff / damping).  Its order is chosen so that the entries of one SUPER-TILE (a run of consecutive
//     tiles, ~64k edges) that go to the same place on the same owner are contiguous: within a super-tile the outbox is
//     sorted by (owner, position at the owner), and because an owner lays out each region of its buffer in global source
//     order, such entries are contiguous AT THE OWNER too.  A super-tile therefore ships as a short list of
//     (outbox range -> owner range) descriptors (host-built, ShipDesc).
//   * SHIPPING.  The persistent grid hands every CTA whole super-tiles (tile = (wave * gridDim + cta) * tps + r), so the
//     CTA that computes a super-tile is also the one that ships it, right after its last tile: all threads carry the
//     super-tile's outbox range to the owners with coalesced 16-byte vector stores over the CUDA-IPC mapping (the
//     destination of each entry from out_rpos; consecutive entries of a run are consecutive at the owner, so a warp's
//     stores merge into full 128-byte lines on their way to NVLink).  No cross-CTA bookkeeping at all -- no per-tile
//     atomic, no fence, no queue -- and the NVLink traffic is spread evenly over all SMs and over the whole sweep: the
//     transfer of one super-tile overlaps the computation of everybody else's.  One kernel does the sweep and its
//     all-to-all.
//     How it got here (2 GPUs, 12.5M nodes each, profiles/dist_exchange_r02.md): "whoever completes a super-tile ships
//     it" (strided tiles, one atomic per tile) concentrates the traffic on the slowest CTAs -- the one that ships falls
//     behind and completes the next super-tile too: 5.3 ms per sweep at 32 tiles per super-tile, 10.4 ms at 128; a
//     device-side queue of chunks any CTA may claim balances that (4.8 ms) but pays a fence + atomic per tile and
//     claim traffic.  TMA bulk copies (cp.async.bulk.global.shared::cta, one per descriptor piece) were measured in the
//     first structure and lost to vector stores (10.7 vs 5.3 ms): a run is ~1 KB here, too small to amortise an elected
//     thread issuing them one by one.
//   * DEVICE-SIDE SWEEP BARRIER.  When a rank's last CTA has seen every CTA report in (each after its bulk copies
//     completed), it writes the rank's row (field partials, max-diff) into EVERY rank's sync block and then raises its
//     flag there (release at system scope).  The next sweep's kernel waits in its prologue until all flags show the
//     previous sweep (acquire), reduces the rows in rank order -- every CTA of every rank the same way, so all ranks
//     take bit-identical decisions -- and goes on.  A batch of sweeps needs no host and no NCCL; the flag wait also
//     keeps a fast rank from overwriting a buffer a slow rank still gathers from.
#pragma once
#include <type_traits>

#include "bp_device.cuh"

namespace sbmbp {

constexpr int kMaxRanks = 8;
constexpr unsigned kRemoteBit = 0x80000000u;  // pos word of a DIST engine: bit 31 set = outbox index, clear = local position

struct ShipDesc {
    unsigned src;   // first outbox entry
    unsigned dst;   // first position in the owner's buffer
    unsigned len;   // messages
    unsigned rank;  // owner
};

// one per rank, CUDA-IPC mapped by every other rank; written by peers, read locally
struct SyncBlock {
    unsigned flag[kMaxRanks];                 // flag[r] = sweeps rank r has completed and shipped (monotonic)
    unsigned pad[8];
    double rows[2][kMaxRanks][kMaxQ + 1];     // rows[sweep parity][r] = rank r's [field partials (Q) | max-diff]
};

struct DistArgs {
    const unsigned *out_start;   // [nsuper + 1]: outbox range of each super-tile
    const unsigned *out_rpos;    // per outbox entry: owner << 29 | position at the owner
    const ShipDesc *ship;        // maximal runs of a super-tile's outbox range that are contiguous at one owner, grouped by
    const unsigned *ship_start;  //   super-tile ([nsuper + 1]): what a TMA bulk copy can carry in one piece
    int ship_tma;                // 1: ship with TMA bulk copies through shared memory (SBMBP_SHIP_TMA), 0: vector stores
    unsigned tps;                // tiles per super-tile; CTA c computes super-tiles c, c + gridDim, ... and ships each itself
    unsigned nsuper;
    SyncBlock *sync[kMaxRanks];  // sync[r]: rank r's block (own: local pointer)
    int rank, world;
    int from_rows;               // 1: the previous sweep was left open (rows in the sync block); 0: field / ctl are current
    unsigned seq;                // sweeps completed when this kernel starts (host-known; valid when from_rows)
#ifdef SBMBP_TUNING
    unsigned dbg;                // timing experiments only: 2 no shipping
    unsigned long long *trace;   // timing experiments only: [4 * sweep + 0..3] = globaltimer at entry, after the flag wait, at publish
#endif
};


__device__ __forceinline__ unsigned long long dx_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned dx_ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dx_st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Carries outbox entries [k_lo, k_hi) to their owners: every thread loads entries (four in flight) and stores each at
// out_rpos[k] of its owner's buffer.  Whole CTA.  peer[r]: rank r's destination buffer (this sweep's S_new).
template <typename T, int QT, int NT>
__device__ __forceinline__ void dist_ship_range(const DistArgs &d, unsigned k_lo, unsigned k_hi, const T *__restrict__ outbox,
                                                T *const *s_peer) {
    constexpr unsigned MB = QT * sizeof(T);
    using V = typename std::conditional<MB % 16 == 0, uint4, uint2>::type;
    constexpr unsigned VPM = MB / sizeof(V);  // vectors per message
    constexpr int U = (VPM == 1) ? 8 : 2;
    const int tid = threadIdx.x;
#ifdef SBMBP_TUNING
    if (d.dbg & 2u) return;
#endif
    for (unsigned k0 = k_lo + tid; k0 < k_hi; k0 += U * NT) {
        unsigned rp[U];
        V val[U][VPM];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned k = k0 + u * NT;
            if (k < k_hi) {
                rp[u] = __ldg(d.out_rpos + k);
#pragma unroll
                for (unsigned w = 0; w < VPM; ++w)
                    val[u][w] = __ldcg(reinterpret_cast<const V *>(outbox + size_t(k) * QT) + w);  // L2: written by other CTAs of this kernel
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned k = k0 + u * NT;
            if (k < k_hi) {
                V *dst = reinterpret_cast<V *>(s_peer[rp[u] >> 29] + size_t(rp[u] & ((1u << 29) - 1u)) * QT);
#pragma unroll
                for (unsigned w = 0; w < VPM; ++w) dst[w] = val[u][w];
            }
        }
    }
}

// ---- TMA flavour of the shipping: shared -> global bulk copies (SASS UBLKCP.G.S), tracked by the issuing thread's bulk groups
__device__ __forceinline__ unsigned dx_smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void dx_bulk_store(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(dx_smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void dx_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void dx_bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void dx_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void dx_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Super-tile sp through `stage` (shared, stage_bytes, 16-byte aligned, free for the duration): the threads fill the stage
// from the outbox, an elected thread sends every descriptor piece inside it as one bulk copy to its owner; the copies
// are asynchronous -- the CTA only waits until the TMA unit has READ the stage before refilling or reusing it, the
// NVLink transfer itself overlaps whatever the CTA does next.  16-byte-multiple messages only.
template <typename T, int QT, int NT>
__device__ __forceinline__ void dist_ship_supertile_tma(const DistArgs &d, unsigned sp, const T *__restrict__ outbox, T *const *s_peer,
                                                        unsigned char *stage, unsigned stage_bytes) {
    constexpr unsigned MB = QT * sizeof(T);
    static_assert(MB % 16 == 0, "bulk copies need 16-byte aligned messages");
    const int tid = threadIdx.x;
    const unsigned o0 = d.out_start[sp], o1 = d.out_start[sp + 1];
    if (o0 == o1) return;
    unsigned di = d.ship_start[sp];
    const unsigned dend = d.ship_start[sp + 1];
    constexpr unsigned kWin = 128;  // descriptors staged in shared memory at a time: the first 2 KB of the stage
    ShipDesc *s_desc = reinterpret_cast<ShipDesc *>(stage);
    stage += kWin * sizeof(ShipDesc);
    stage_bytes -= kWin * sizeof(ShipDesc);
    unsigned wb = di;
    for (unsigned k = tid; k < kWin && wb + k < dend; k += NT) s_desc[k] = d.ship[wb + k];
    __syncthreads();
    ShipDesc cur = s_desc[0];
    const unsigned cap = stage_bytes / MB;
    for (unsigned c0 = o0; c0 < o1; c0 += cap) {
        const unsigned n = min(cap, o1 - c0);
        if (tid == 0) dx_bulk_wait_read_all();  // the previous fill has left the stage
        __syncthreads();
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(outbox + size_t(c0) * QT);
            uint4 *dst = reinterpret_cast<uint4 *>(stage);
            const unsigned n16 = n * (MB / 16);
            for (unsigned k = tid; k < n16; k += NT) dst[k] = __ldcg(src + k);
        }
        dx_fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            unsigned at = c0;
            while (at < c0 + n) {
                while (cur.src + cur.len <= at) {
                    ++di;
                    if (di >= wb + kWin) {
                        wb = di;
                        for (unsigned k = 0; k < kWin && wb + k < dend; ++k) s_desc[k] = d.ship[wb + k];
                    }
                    cur = s_desc[di - wb];
                }
                const unsigned take = min(cur.src + cur.len, c0 + n) - at;
                dx_bulk_store(s_peer[cur.rank] + size_t(cur.dst + (at - cur.src)) * QT, stage + size_t(at - c0) * MB, take * MB);
                at += take;
            }
            dx_bulk_commit();
        }
    }
    if (tid == 0) dx_bulk_wait_read_all();  // the stage goes back to the pipeline
    __syncthreads();
}

// before a CTA reports in: what it shipped (the vector stores of all its threads, the elected thread's bulk copies) has
// completed and is visible at system scope
__device__ __forceinline__ void dist_ship_drain() {
    __threadfence_system();
    if (threadIdx.x == 0) {
        dx_bulk_wait_all();
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence_system();
    }
}

// The rank's reduced row of sweep `seq` -> every rank's sync block, then the flag (called by the last CTA, tid < QT + 1
// hold row[tid]).  rows parity = seq & 1; flag value = seq + 1.
template <int QT>
__device__ __forceinline__ void dist_publish_row(const DistArgs &d, const double *s_tot, unsigned seq) {
    const int tid = threadIdx.x;
    if (tid <= QT) {
        const double v = s_tot[tid];
        const int col = (tid < QT) ? tid : kMaxQ;
        for (int r = 0; r < d.world; ++r) d.sync[r]->rows[seq & 1u][d.rank][col] = v;
        __threadfence_system();
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        for (int r = 0; r < d.world; ++r) dx_st_release_sys(&d.sync[r]->flag[d.rank], seq + 1u);
#ifdef SBMBP_TUNING
        if (d.trace) d.trace[4 * (seq & 63u) + 2] = dx_now();
#endif
    }
}

// Waits until every rank has completed sweep seq - 1 (flags >= seq), then combines the rows of that sweep in rank order:
// s_tot[0 .. QT) = field partials, s_tot[QT] = max-diff.  Whole CTA; identical result in every CTA of every rank.
template <int QT>
__device__ __forceinline__ void dist_wait_and_reduce(const DistArgs &d, unsigned seq, double *s_tot) {
    const int tid = threadIdx.x;
    const SyncBlock *mine = d.sync[d.rank];
#ifdef SBMBP_TUNING
    if (d.trace && blockIdx.x == 0 && tid == 0) d.trace[4 * (seq & 63u) + 0] = dx_now();
#endif
    if (tid < d.world) {
        while (dx_ld_acquire_sys(&mine->flag[tid]) < seq) __nanosleep(64);
    }
    __syncthreads();
#ifdef SBMBP_TUNING
    if (d.trace && blockIdx.x == 0 && tid == 0) d.trace[4 * (seq & 63u) + 1] = dx_now();
#endif
    if (tid <= QT) {
        const int col = (tid < QT) ? tid : kMaxQ;
        const unsigned par = (seq - 1u) & 1u;
        double r = 0.0;
        for (int k = 0; k < d.world; ++k) {
            const double v = *reinterpret_cast<const volatile double *>(&mine->rows[par][k][col]);
            r = (tid < QT) ? r + v : fmax(r, v);
        }
        s_tot[tid] = r;
    }
    __syncthreads();
}

}  // namespace sbmbp
