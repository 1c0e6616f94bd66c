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
//   * SHIPPING.  The CTA that finishes the last tile of a super-tile (one atomic per tile) streams that super-tile's
//     outbox range through shared memory and sends every descriptor piece as ONE bulk copy to the owner's buffer:
//     cp.async.bulk.global.shared::cta over the CUDA-IPC mapping (SASS UBLKCP.G.S) -- full-line NVLink writes posted by
//     the TMA unit while every other CTA keeps computing.  The transfer of a super-tile overlaps the computation of the
//     following ones: one kernel does the sweep and its all-to-all.
//   * DEVICE-SIDE SWEEP BARRIER.  When a rank's last CTA has seen every CTA report in (each after its bulk copies
//     completed), it writes the rank's row (field partials, max-diff) into EVERY rank's sync block and then raises its
//     flag there (release at system scope).  The next sweep's kernel waits in its prologue until all flags show the
//     previous sweep (acquire), reduces the rows in rank order -- every CTA of every rank the same way, so all ranks
//     take bit-identical decisions -- and goes on.  A batch of sweeps needs no host and no NCCL; the flag wait also
//     keeps a fast rank from overwriting a buffer a slow rank still gathers from.
#pragma once
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
    const ShipDesc *ship;        // descriptors, grouped by super-tile
    const unsigned *ship_start;  // [nsuper + 1]
    const unsigned *out_start;   // [nsuper + 1]: outbox range of each super-tile
    unsigned *st_done;           // [nsuper]: tiles finished, cumulative over sweeps
    unsigned tps;                // tiles per super-tile
    unsigned nsuper;
    SyncBlock *sync[kMaxRanks];  // sync[r]: rank r's block (own: local pointer)
    int rank, world;
    int from_rows;               // 1: the previous sweep was left open (rows in the sync block); 0: field / ctl are current
    unsigned seq;                // sweeps completed when this kernel starts (host-known; valid when from_rows)
};

// ---- mbarrier-free TMA bulk store primitives (shared -> global); completion by the issuing thread's bulk groups
__device__ __forceinline__ unsigned dx_smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void dx_bulk_store(void *gdst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(dx_smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void dx_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void dx_bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void dx_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void dx_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned dx_ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dx_st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Ships super-tile s: outbox entries [out_start[s], out_start[s+1]) through `stage` (shared, stage_bytes, 16-byte aligned,
// free for the duration) to their owners.  Whole CTA; returns with the bulk copies committed (not necessarily complete):
// call dist_ship_drain before the CTA reports in.  peer[r]: rank r's destination buffer (this sweep's S_new).
template <typename T, int QT, int NT>
__device__ __forceinline__ void dist_ship_supertile(const DistArgs &d, unsigned s, const T *__restrict__ outbox, T *const *peer,
                                                    unsigned char *stage, unsigned stage_bytes) {
    constexpr unsigned MB = QT * sizeof(T);
    const int tid = threadIdx.x;
    const unsigned o0 = d.out_start[s], o1 = d.out_start[s + 1];
    if (o0 == o1) return;
    unsigned di = d.ship_start[s];
    const unsigned dend = d.ship_start[s + 1];
    if constexpr (MB % 16 != 0) {
        // 8-byte messages (Q = 2 in FP32): positions at the owner are only 8-byte aligned, which a bulk copy cannot
        // address -- coalesced vector stores instead (a warp still writes 256 contiguous bytes)
        for (; di < dend; ++di) {
            const ShipDesc sd = d.ship[di];
            const uint2 *src = reinterpret_cast<const uint2 *>(outbox + size_t(sd.src) * QT);
            uint2 *dst = reinterpret_cast<uint2 *>(peer[sd.rank] + size_t(sd.dst) * QT);
            for (unsigned k = tid; k < sd.len; k += NT) dst[k] = __ldcg(src + k);  // L2: written by other CTAs of this kernel
        }
        return;
    } else {
        const unsigned cap = stage_bytes / MB;  // messages per stage fill
        ShipDesc cur = d.ship[di];
        for (unsigned c0 = o0; c0 < o1; c0 += cap) {
            const unsigned n = min(cap, o1 - c0);
            // the previous fill's bulk copies must have finished reading the stage
            if (tid == 0) dx_bulk_wait_read_all();
            __syncthreads();
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(outbox + size_t(c0) * QT);
                uint4 *dst = reinterpret_cast<uint4 *>(stage);
                const unsigned n16 = n * (MB / 16);
                for (unsigned k = tid; k < n16; k += NT) dst[k] = __ldcg(src + k);  // L2: written by other CTAs of this kernel
            }
            dx_fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                // every descriptor piece inside [c0, c0 + n): one bulk copy each (descriptors are sorted by src and tile
                // the outbox range without gaps)
                unsigned at = c0;
                while (at < c0 + n) {
                    while (cur.src + cur.len <= at) cur = d.ship[++di];
                    const unsigned take = min(cur.src + cur.len, c0 + n) - at;
                    dx_bulk_store(peer[cur.rank] + size_t(cur.dst + (at - cur.src)) * QT, stage + size_t(at - c0) * MB, take * MB);
                    at += take;
                }
                dx_bulk_commit();
            }
        }
    }
}

// before a CTA reports in: its bulk copies have completed and are visible at system scope
__device__ __forceinline__ void dist_ship_drain() {
    if (threadIdx.x == 0) {
        dx_bulk_wait_all();
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence_system();
    }
}

// One tile finished (all its outbox writes issued by this CTA): returns true (uniformly) if it was the last tile of its
// super-tile in this sweep, i.e. the caller must ship it.  seq = sweeps completed before this one.
__device__ __forceinline__ bool dist_tile_done(const DistArgs &d, unsigned tile, unsigned ntiles, unsigned seq, int *s_flag) {
    const unsigned s = tile / d.tps;
    const unsigned in_s = min(d.tps, ntiles - s * d.tps);
    __threadfence();  // this thread's outbox stores before the count (gpu scope)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned before = atomicAdd(d.st_done + s, 1u);
        *s_flag = (before + 1u == (seq + 1u) * in_s) ? 1 : 0;
        __threadfence();  // the counts seen -> the outbox entries read by the shipper
    }
    __syncthreads();
    return *s_flag != 0;
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
    }
}

// Waits until every rank has completed sweep seq - 1 (flags >= seq), then combines the rows of that sweep in rank order:
// s_tot[0 .. QT) = field partials, s_tot[QT] = max-diff.  Whole CTA; identical result in every CTA of every rank.
template <int QT>
__device__ __forceinline__ void dist_wait_and_reduce(const DistArgs &d, unsigned seq, double *s_tot) {
    const int tid = threadIdx.x;
    const SyncBlock *mine = d.sync[d.rank];
    if (tid < d.world) {
        while (dx_ld_acquire_sys(&mine->flag[tid]) < seq) __nanosleep(64);
    }
    __syncthreads();
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
