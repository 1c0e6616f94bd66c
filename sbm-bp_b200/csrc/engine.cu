// libsbmbp engine: device state, kernel dispatch and the C ABI of include/sbmbp.h.
// Host-side drivers mirror the reference's converge / inference pieces / learning
// (belief_propagation.cpp:14-99, :386-415) on device-resident state.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <random>
#include <string>
#include <atomic>
#include <thread>
#include <utility>
#include <vector>

#include "engine.hpp"
#include "reduce_kernels.cuh"
#include "state_kernels.cuh"
#include "sweep_replay.cuh"

// Destination-bucketed message layout.  Buckets are ranges of consecutive nodes whose in-slots cover about
// `region_slots` messages; region b of the buffer is exactly the in-slot range of bucket b, filled in source-slot
// order.  pos[e] = where the message out of slot e is stored; gather[e] = pos[rev[e]] = where the message into
// slot e is found.  Returns the number of buckets (1: identity layout).
static unsigned build_layout(const sbmbp_graph &g, uint64_t region_slots, std::vector<unsigned> &pos,
                             std::vector<unsigned> &gather) {
    const uint64_t M = g.M;
    gather.assign(g.rev.begin(), g.rev.end());
    pos.resize(M);
    for (uint64_t s = 0; s < M; ++s) pos[s] = unsigned(s);
    if (M == 0 || region_slots == 0 || M <= region_slots) return 1;
    std::vector<unsigned> bucket_of(g.N);
    std::vector<uint64_t> cursor;
    uint64_t next = 0;
    unsigned b = 0;
    for (uint32_t i = 0; i < g.N; ++i) {
        if (g.row_ptr[i] >= next) {  // node i opens a new bucket
            cursor.push_back(g.row_ptr[i]);
            next = g.row_ptr[i] + region_slots;
            b = unsigned(cursor.size() - 1);
        }
        bucket_of[i] = b;
    }
    if (cursor.size() <= 1) return 1;
    for (uint64_t s = 0; s < M; ++s) pos[s] = unsigned(cursor[bucket_of[g.col[s]]]++);
    for (uint64_t s = 0; s < M; ++s) gather[s] = pos[g.rev[s]];
    return unsigned(cursor.size());
}

// Per tile, reorder the out-message positions ascending and pack, per buffer entry, which tile-local slot and node
// it belongs to and whether that node updates in the log domain (hub tiles keep slot order; their info is unused).
// pos: slot order in, tile-sorted out.
static void sort_tile_positions(const sbmbp_graph &g, const std::vector<Tile> &tiles, int te,
                                std::vector<unsigned> &pos, std::vector<unsigned> &info) {
    info.assign(pos.size(), 0);
    const unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    auto work = [&](size_t lo, size_t hi) {
        std::vector<std::pair<unsigned, unsigned>> tmp;
        for (size_t b = lo; b < hi; ++b) {
            const Tile &t = tiles[b];
            if (t.ne > unsigned(te)) continue;
            tmp.resize(t.ne);
            for (unsigned n = 0; n < t.nn; ++n) {
                const uint32_t node = t.n0 + n;
                const unsigned flag = (g.deg[node] >= kLargeDegree) ? 0x80000000u : 0u;
                for (uint64_t s = g.row_ptr[node]; s < g.row_ptr[node + 1]; ++s) {
                    const unsigned k = unsigned(s - t.e0);
                    tmp[k] = {pos[s], flag | (n << 16) | k};
                }
            }
            std::sort(tmp.begin(), tmp.end());
            for (unsigned k = 0; k < t.ne; ++k) {
                pos[t.e0 + k] = tmp[k].first;
                info[t.e0 + k] = tmp[k].second;
            }
        }
    };
    std::vector<std::thread> pool;
    const size_t per = (tiles.size() + nthreads - 1) / nthreads;
    for (unsigned i = 0; i < nthreads; ++i) {
        const size_t lo = std::min(tiles.size(), size_t(i) * per), hi = std::min(tiles.size(), lo + per);
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto &th : pool) th.join();
}

int sync_marg(sbmbp_engine *e) {
    if (!e->marg_ell_dirty) return SBMBP_OK;
    const unsigned blocks = std::max(1u, std::min((e->ell_entries + 255u) / 256u, unsigned(8 * e->sm_count)));
    ell_scatter_marg_kernel<<<blocks, 256, 0, e->stream>>>(e->d_marg_ell, e->d_ell_node, e->ell_entries, e->Q, e->d_marg);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    e->marg_ell_dirty = false;
    return SBMBP_OK;
}

int ensure_full(sbmbp_engine *e) {
    if (!e->compact) return SBMBP_OK;
    const unsigned b = e->sweeps_done & 1u;
    const unsigned blocks = unsigned(std::min<uint64_t>((e->buf_slots + 255) / 256, uint64_t(e->sm_count) * 16));
    compact_unpack_kernel<<<blocks, 256, 0, e->stream>>>(static_cast<const double *>(e->d_C[b]), static_cast<double *>(e->d_S[b]),
                                                       e->buf_slots);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    e->compact = false;
    return SBMBP_OK;
}

int ensure_compact(sbmbp_engine *e, bool *now_compact) {
    *now_compact = e->compact;
    if (e->compact || !e->compact_ok || e->compact_refused) return SBMBP_OK;
    const unsigned b = e->sweeps_done & 1u;
    unsigned long long *d_bad = reinterpret_cast<unsigned long long *>(e->d_out);
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), e->stream));
    const unsigned blocks = unsigned(std::min<uint64_t>((e->buf_slots + 255) / 256, uint64_t(e->sm_count) * 16));
    compact_pack_kernel<<<blocks, 256, 0, e->stream>>>(static_cast<const double *>(e->d_S[b]), static_cast<double *>(e->d_C[b]),
                                                     e->buf_slots, d_bad);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(e->h_out, d_bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (*reinterpret_cast<unsigned long long *>(e->h_out) != 0ull) {
        e->compact_refused = true;  // e.g. init flag 2 (un-normalised in-slots, :185-191): this state stays on full storage
        return SBMBP_OK;
    }
    e->compact = true;
    *now_compact = true;
    return SBMBP_OK;
}

int ensure_scratch(sbmbp_engine *e, size_t doubles) {
    if (doubles <= e->scratch_doubles) return SBMBP_OK;
    if (e->d_scratch) cudaFree(e->d_scratch);
    e->d_scratch = nullptr;
    e->scratch_doubles = 0;
    CUDA_TRY(cudaMalloc(&e->d_scratch, doubles * sizeof(double)));
    e->scratch_doubles = doubles;
    return SBMBP_OK;
}

int reduce_columns(sbmbp_engine *e, const double *d_partial, unsigned nrows, unsigned ncols, double *d_result) {
    reduce_cols_kernel<<<ncols, kThreads, 0, e->stream>>>(d_partial, nrows, ncols, d_result);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    return SBMBP_OK;
}

namespace {

constexpr size_t kOutDoubles = 8 + 2 * kMaxQ + 2 * kMaxQ * kMaxQ;
// slack behind row_ptr / rev / pos / info: the tile kernels stage them with 16-byte aligned bulk copies (sweep_tile.cuh),
// which may start up to 12 bytes before a tile's first entry and end up to 12 bytes after its last
constexpr size_t kBulkPad = 32;

int pick_qt(uint32_t Q) {
    for (int qt : {2, 4, 8, 16, 32})
        if (Q <= uint32_t(qt)) return qt;
    return 0;
}

template <typename F>
int dispatch(const sbmbp_engine *e, F &&f) {
#define CASE(QT)                                         \
    case QT:                                             \
        if (e->prec == SBMBP_F64) return f(double(0), std::integral_constant<int, QT>()); \
        return f(float(0), std::integral_constant<int, QT>());
    switch (e->qt) {
        CASE(2)
        CASE(4)
        CASE(8)
        CASE(16)
        CASE(32)
    }
#undef CASE
    set_error("unsupported Q");
    return SBMBP_ERR_UNSUPPORTED;
}

template <typename T, int QT>
void tile_geometry(int &te, int &tn) {
    te = TileCfg<T, QT>::TE;
    tn = TileCfg<T, QT>::TN;
}

// node-aligned greedy tiling: <= te edges and <= tn nodes per tile; a node with more than te edges is a hub tile
std::vector<Tile> make_tiles(const sbmbp_graph &g, int te, int tn) {
    std::vector<Tile> tiles;
    uint32_t n = 0;
    while (n < g.N) {
        Tile t;
        t.n0 = n;
        t.e0 = g.row_ptr[n];
        t.nbig = 0;
        if (g.deg[n] > uint32_t(te)) {
            t.nn = 1;
            t.ne = g.deg[n];
            ++n;
        } else {
            uint64_t edges = 0;
            uint32_t cnt = 0;
            while (n < g.N && cnt < uint32_t(tn) && g.deg[n] <= uint32_t(te) && edges + g.deg[n] <= uint64_t(te)) {
                edges += g.deg[n];
                if (g.deg[n] >= 32) t.nbig++;
                ++cnt;
                ++n;
            }
            t.nn = cnt;
            t.ne = unsigned(edges);
        }
        tiles.push_back(t);
    }
    return tiles;
}

// Degree-class (ELL) layout (sweep_ell.cuh), destination-bucketed like build_layout: buckets are ranges of
// consecutive nodes whose in-slots cover about `region_slots` messages, region b of the buffer holds the messages
// INTO bucket b.  Inside a bucket the nodes of degree d < 32 form a class, cut into chunks of 32 nodes (one warp,
// one node per lane); the index words of slot l of lane r of a chunk sit at  chunk base + 32 l + r  (rev = where the
// in-message is, pos = where the out-message goes).  Regions are filled in processing order (chunk, slot, lane), so
// the out-messages of one (chunk, slot) that go to the same bucket are consecutive.  Nodes of degree >= 32 (warp /
// hub kernels) fill the regions last.  Returns pos / gather per slot like build_layout plus the ELL-side arrays.
unsigned build_bell_layout(const sbmbp_graph &g, uint64_t region_slots, std::vector<unsigned> &pos,
                           std::vector<unsigned> &gather, std::vector<EllClass> &cls, std::vector<unsigned> &ell_node,
                           std::vector<unsigned> &ell_rev, std::vector<unsigned> &ell_pos, unsigned &nchunks) {
    const uint64_t M = g.M;
    // buckets over consecutive nodes
    std::vector<unsigned> bucket_of(g.N);
    std::vector<uint32_t> bucket_first;
    std::vector<uint64_t> cursor;
    {
        uint64_t next = 0;
        for (uint32_t i = 0; i < g.N; ++i) {
            if (bucket_first.empty() || (region_slots != 0 && g.row_ptr[i] >= next)) {
                bucket_first.push_back(i);
                cursor.push_back(g.row_ptr[i]);
                next = g.row_ptr[i] + region_slots;
            }
            bucket_of[i] = unsigned(bucket_first.size() - 1);
        }
    }
    const unsigned nb = unsigned(bucket_first.size());
    bucket_first.push_back(g.N);
    // classes (bucket, degree) and the class-ordered node list
    cls.clear();
    ell_node.clear();
    unsigned chunk_first = 0;
    uint64_t idx_base = 0;
    for (unsigned b = 0; b < nb; ++b) {
        std::vector<std::vector<uint32_t>> by_deg(kEllDegrees);
        for (uint32_t i = bucket_first[b]; i < bucket_first[b + 1]; ++i)
            if (g.deg[i] < kEllDegrees) by_deg[g.deg[i]].push_back(i);
        for (unsigned d = 0; d < kEllDegrees; ++d) {
            if (by_deg[d].empty()) continue;
            EllClass c;
            std::memset(&c, 0, sizeof(c));
            c.d = d;
            c.n = unsigned(by_deg[d].size());
            c.node_first = unsigned(ell_node.size());
            c.chunk_first = chunk_first;
            c.base = unsigned(idx_base);
            cls.push_back(c);
            ell_node.insert(ell_node.end(), by_deg[d].begin(), by_deg[d].end());
            const unsigned nch = (c.n + 31) / 32;
            chunk_first += nch;
            idx_base += uint64_t(nch) * 32 * d;
        }
    }
    nchunks = chunk_first;
    // positions in processing order
    pos.assign(M, 0);
    for (const EllClass &c : cls)
        for (unsigned r0 = 0; r0 < c.n; r0 += 32)
            for (unsigned l = 0; l < c.d; ++l)
                for (unsigned r = r0; r < std::min(c.n, r0 + 32); ++r) {
                    const uint64_t s = g.row_ptr[ell_node[c.node_first + r]] + l;
                    pos[s] = unsigned(cursor[bucket_of[g.col[s]]]++);
                }
    for (uint32_t i = 0; i < g.N; ++i)
        if (g.deg[i] >= kEllDegrees)
            for (uint64_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) pos[s] = unsigned(cursor[bucket_of[g.col[s]]]++);
    gather.resize(M);
    for (uint64_t s = 0; s < M; ++s) gather[s] = pos[g.rev[s]];
    ell_rev.assign(idx_base, 0);
    ell_pos.assign(idx_base, 0);
    for (const EllClass &c : cls)
        for (unsigned r = 0; r < c.n; ++r) {
            const uint64_t s0 = g.row_ptr[ell_node[c.node_first + r]];
            const uint64_t ib = uint64_t(c.base) + uint64_t(r / 32) * 32 * c.d + (r % 32);
            for (unsigned l = 0; l < c.d; ++l) {
                ell_rev[ib + 32 * l] = gather[s0 + l];
                ell_pos[ib + 32 * l] = pos[s0 + l];
            }
        }
    return nb;
}

// Padded one-bucket degree-class layout: ONE destination bucket, classes by degree over all nodes, chunks of 32 nodes padded to 32
// lanes, and buffer position == index-word offset: the out-message of slot l of lane r of chunk k of class c sits at
// c.base + 32 d k + 32 l + r.  A chunk's old / new out-messages are therefore one contiguous block (one TMA bulk copy
// each way) and no pos array is needed.  ell_node is padded the same way (32 entries per chunk, ~0u = no node), so that
// entry 32 k' + r of the chunk-ordered marginal array belongs to lane r of chunk k'.  Nodes of degree >= 32 (warp / hub
// kernels) take the positions after the last chunk.  total_slots: slots of the message buffers (>= M: padding).
void build_ell_padded_layout(const sbmbp_graph &g, std::vector<unsigned> &pos, std::vector<unsigned> &gather,
                       std::vector<EllClass> &cls, std::vector<unsigned> &ell_node, std::vector<unsigned> &ell_rev,
                       unsigned &nchunks, uint64_t &total_slots) {
    const uint64_t M = g.M;
    std::vector<std::vector<uint32_t>> by_deg(kEllDegrees);
    for (uint32_t i = 0; i < g.N; ++i)
        if (g.deg[i] < kEllDegrees) by_deg[g.deg[i]].push_back(i);
    cls.clear();
    ell_node.clear();
    unsigned chunk_first = 0;
    uint64_t idx_base = 0;
    for (unsigned d = 0; d < kEllDegrees; ++d) {
        if (by_deg[d].empty()) continue;
        EllClass c;
        std::memset(&c, 0, sizeof(c));
        c.d = d;
        c.n = unsigned(by_deg[d].size());
        c.node_first = 32u * chunk_first;
        c.chunk_first = chunk_first;
        c.base = unsigned(idx_base);
        cls.push_back(c);
        const unsigned nch = (c.n + 31) / 32;
        ell_node.insert(ell_node.end(), by_deg[d].begin(), by_deg[d].end());
        ell_node.resize(size_t(32) * (chunk_first + nch), 0xffffffffu);
        chunk_first += nch;
        idx_base += uint64_t(nch) * 32 * d;
    }
    nchunks = chunk_first;
    pos.assign(M, 0);
    for (const EllClass &c : cls)
        for (unsigned r = 0; r < c.n; ++r) {
            const uint64_t s0 = g.row_ptr[ell_node[c.node_first + r]];
            const uint64_t ib = uint64_t(c.base) + uint64_t(r / 32) * 32 * c.d + (r % 32);
            for (unsigned l = 0; l < c.d; ++l) pos[s0 + l] = unsigned(ib + 32 * l);
        }
    uint64_t cursor = idx_base;
    for (uint32_t i = 0; i < g.N; ++i)
        if (g.deg[i] >= kEllDegrees)
            for (uint64_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) pos[s] = unsigned(cursor++);
    total_slots = std::max<uint64_t>(cursor, 1);
    gather.resize(M);
    for (uint64_t s = 0; s < M; ++s) gather[s] = pos[g.rev[s]];
    ell_rev.assign(idx_base, 0);  // padding lanes gather slot 0 (any valid address) and are never used
    for (const EllClass &c : cls)
        for (unsigned r = 0; r < c.n; ++r) {
            const uint64_t s0 = g.row_ptr[ell_node[c.node_first + r]];
            const uint64_t ib = uint64_t(c.base) + uint64_t(r / 32) * 32 * c.d + (r % 32);
            for (unsigned l = 0; l < c.d; ++l) ell_rev[ib + 32 * l] = gather[s0 + l];
        }
}

// Work lists of the ELL kernel: every chunk (32 lanes of one class) goes to one warp of the persistent grid.  Chunks cost
// differently (degree 8 is several times degree 1; degrees above the unrolled ones run the two-pass loop), so the
// chunks are dealt out longest-first to the least loaded warp (LPT) and each warp's list is then put back in chunk
// order.  Static, so the field sums stay bitwise reproducible.  sched[w * len + i]: x = index-array offset of the chunk,
// y = offset into ell_node, z = degree | lanes << 8; padding entries are all zero.
void build_ell_schedule(const std::vector<EllClass> &cls, unsigned nwarps, unsigned unroll_degree, std::vector<uint4> &sched,
                        unsigned &len) {
    struct Chunk {
        unsigned x, y, z, order;
        unsigned long long weight;
    };
    std::vector<Chunk> chunks;
    unsigned order = 0;
    for (const EllClass &c : cls)
        for (unsigned k = 0; k * 32 < c.n; ++k) {
            Chunk ch;
            ch.x = c.base + k * 32 * c.d;
            ch.y = c.node_first + k * 32;
            ch.z = c.d | (std::min(32u, c.n - k * 32) << 8);
            ch.order = order++;
            ch.weight = 64ull + 32ull * c.d * (c.d <= unroll_degree ? 1ull : 3ull);
            chunks.push_back(ch);
        }
    std::vector<unsigned> by_weight(chunks.size());
    for (unsigned i = 0; i < chunks.size(); ++i) by_weight[i] = i;
    std::stable_sort(by_weight.begin(), by_weight.end(),
                     [&](unsigned p, unsigned q) { return chunks[p].weight > chunks[q].weight; });
    typedef std::pair<unsigned long long, unsigned> Load;  // (load, warp)
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (unsigned w = 0; w < nwarps; ++w) heap.push(Load(0ull, w));
    std::vector<std::vector<unsigned>> lists(nwarps);
    for (unsigned i : by_weight) {
        Load top = heap.top();
        heap.pop();
        lists[top.second].push_back(i);
        heap.push(Load(top.first + chunks[i].weight, top.second));
    }
    len = 0;
    for (auto &l : lists) {
        std::sort(l.begin(), l.end());  // chunk ids are in processing order
        len = std::max<unsigned>(len, unsigned(l.size()));
    }
    sched.assign(size_t(nwarps) * len, make_uint4(0u, 0u, 0u, 0u));
    for (unsigned w = 0; w < nwarps; ++w)
        for (unsigned i = 0; i < lists[w].size(); ++i) {
            const Chunk &ch = chunks[lists[w][i]];
            sched[size_t(w) * len + i] = make_uint4(ch.x, ch.y, ch.z, ch.order);
        }
}

// Warp tiles (sweep_warp.cuh): node-aligned runs of <= 32 nodes of degree < 32 and <= we edge slots (kind 0), single
// nodes of degree 32..we (kind 1); nodes with more than we edges are hubs and get a CTA each (hub kernel).
void make_wtiles(const sbmbp_graph &g, int we, uint32_t min_degree, std::vector<WTile> &wt, std::vector<Tile> &hubs) {
    wt.clear();
    hubs.clear();
    auto emit = [&](uint64_t e0, uint32_t n0, unsigned ne, unsigned nn, unsigned kind) {
        WTile w;
        w.e0lo = unsigned(e0 & 0xffffffffull);
        w.e0hi = unsigned(e0 >> 32);
        w.n0 = n0;
        w.packed = ne | (nn << 16) | (kind << 24);
        wt.push_back(w);
    };
    uint32_t n = 0;
    while (n < g.N) {
        const uint32_t d = g.deg[n];
        if (d < min_degree) {  // left to the degree-class kernel
            ++n;
            continue;
        }
        if (d > uint32_t(we)) {
            Tile t;
            t.e0 = g.row_ptr[n];
            t.n0 = n;
            t.nn = 1;
            t.ne = d;
            t.nbig = 1;
            hubs.push_back(t);
            ++n;
        } else if (d >= 32) {
            emit(g.row_ptr[n], n, d, 1, 1);
            ++n;
        } else {
            const uint32_t n0 = n;
            unsigned ne = 0, nn = 0;
            while (n < g.N && nn < 32 && g.deg[n] < 32 && g.deg[n] >= min_degree && ne + g.deg[n] <= unsigned(we)) {
                ne += g.deg[n];
                ++nn;
                ++n;
            }
            emit(g.row_ptr[n0], n0, ne, nn, 0);
        }
    }
}

// pos: slot order in; out: ascending inside each warp tile, with info = tile-local slot | tile-local node << 7
void sort_wtile_positions(const sbmbp_graph &g, const std::vector<WTile> &wt, std::vector<unsigned> &pos,
                          std::vector<unsigned short> &info) {
    info.assign(pos.size(), 0);
    const unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    auto work = [&](size_t lo, size_t hi) {
        std::pair<unsigned, unsigned short> tmp[128];
        for (size_t b = lo; b < hi; ++b) {
            const WTile &t = wt[b];
            const uint64_t e0 = t.e0();
            const unsigned ne = t.ne();
            for (unsigned n = 0; n < t.nn(); ++n) {
                const uint32_t node = t.n0 + n;
                for (uint64_t s = g.row_ptr[node]; s < g.row_ptr[node + 1]; ++s) {
                    const unsigned k = unsigned(s - e0);
                    tmp[k] = {pos[s], (unsigned short)(k | (n << kWInfoNodeShift))};
                }
            }
            std::sort(tmp, tmp + ne);
            for (unsigned k = 0; k < ne; ++k) {
                pos[e0 + k] = tmp[k].first;
                info[e0 + k] = tmp[k].second;
            }
        }
    };
    std::vector<std::thread> pool;
    const size_t per = (wt.size() + nthreads - 1) / nthreads;
    for (unsigned i = 0; i < nthreads; ++i) {
        const size_t lo = std::min(wt.size(), size_t(i) * per), hi = std::min(wt.size(), lo + per);
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto &th : pool) th.join();
}

int arm_ctl(sbmbp_engine *e, float crit, unsigned add_sweeps) {
    ctl_arm_kernel<<<1, 32, 0, e->stream>>>(e->d_ctl, crit, add_sweeps);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    return SBMBP_OK;
}

int download_ctl(sbmbp_engine *e) {
    CUDA_TRY(cudaMemcpyAsync(e->h_ctl, e->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->sweeps_done = e->h_ctl->sweeps_done;
    return SBMBP_OK;
}

int need(sbmbp_engine *e, bool params, bool state) {
    if (!e) {
        set_error("null engine");
        return SBMBP_ERR_ARG;
    }
    if (params && !e->have_params) {
        set_error("set_params has not been called");
        return SBMBP_ERR_STATE;
    }
    if (state && !e->have_state) {
        set_error("no message state: call init_random or set_state first");
        return SBMBP_ERR_STATE;
    }
    CUDA_TRY(cudaSetDevice(e->device));
    return SBMBP_OK;
}

// init_h: recompute h from the current marginals into the field the next sweep reads
int ensure_field(sbmbp_engine *e) {
    if (e->field_valid) return SBMBP_OK;
    if (e->dist) {
        set_error("multi-GPU engine: the field needs sbmbp_dist_field_local + all-gather + sbmbp_dist_finalize(advance=0)");
        return SBMBP_ERR_STATE;
    }
    const unsigned blocks = std::max(1u, std::min((e->N + kThreads - 1) / kThreads, unsigned(8 * e->sm_count)));
    TRY(ensure_scratch(e, size_t(blocks) * kMaxQ));
    TRY(sync_marg(e));
    field_partial_kernel<<<blocks, kThreads, 0, e->stream>>>(e->d_marg, e->d_row_ptr, e->N, e->Q, e->dc,
                                                            e->d_scratch);
    field_final_kernel<<<1, kThreads, 0, e->stream>>>(e->d_scratch, blocks, e->d_prm, e->Q, e->d_field[0],
                                                      e->d_field[1], e->d_ctl);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 2;
    e->field_valid = true;
    return SBMBP_OK;
}

int run_sweeps(sbmbp_engine *e, unsigned count, double damping) {
    if (e->ntiles == 0 || count == 0) return SBMBP_OK;
    return dispatch(e, [&](auto t, auto qt) { return launch_sweeps<decltype(t), decltype(qt)::value>(e, count, damping); });
}

// which = 0: kernel Ks / beta (free energy); which = 1: kernel C / 1 (entropy, EM statistics).
// They coincide unless dc == 0 and beta != 1.
int energy_pass(sbmbp_engine *e, int which, const std::vector<double> **out) {
    if (which == 1 && !(e->dc == 0 && e->beta != 1.0)) which = 0;
    if (e->energy_version[which] != e->state_version) {
        TRY(ensure_field(e));
        TRY(dispatch(e, [&](auto t, auto qt) {
            return launch_energy<decltype(t), decltype(qt)::value>(e, which, e->energy_out[which]);
        }));
        e->energy_version[which] = e->state_version;
    }
    *out = &e->energy_out[which];
    return SBMBP_OK;
}

int ensure_col(sbmbp_engine *e) {
    if (e->d_col || e->M == 0) return SBMBP_OK;
    CUDA_TRY(cudaMalloc(&e->d_col, e->M * sizeof(unsigned)));
    CUDA_TRY(cudaMemcpyAsync(e->d_col, e->g->col.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice, e->stream));
    return SBMBP_OK;
}

int node_stats(sbmbp_engine *e, const uint32_t *true_conf, std::vector<double> &row) {
    const unsigned blocks = std::max(1u, std::min((e->N + kThreads - 1) / kThreads, unsigned(4 * e->sm_count)));
    TRY(ensure_scratch(e, size_t(blocks) * kNodeCols + kNodeCols));
    if (true_conf) {
        if (!e->d_true) CUDA_TRY(cudaMalloc(&e->d_true, std::max<size_t>(e->N, 1) * sizeof(unsigned)));
        CUDA_TRY(cudaMemcpyAsync(e->d_true, true_conf, size_t(e->N) * sizeof(unsigned), cudaMemcpyHostToDevice,
                                 e->stream));
    }
    CUDA_TRY(cudaMemsetAsync(e->d_scratch, 0, (size_t(blocks) * kNodeCols + kNodeCols) * sizeof(double), e->stream));
    const size_t smem = size_t(kThreads / 32) * e->Q * e->Q * sizeof(double);
    static bool attr_by_device[kMaxDevices] = {};  // cudaFuncSetAttribute is per device
    bool &attr_set = attr_by_device[e->device];
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(node_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(size_t(kThreads / 32) * kMaxQ * kMaxQ * sizeof(double))));
        attr_set = true;
    }
    TRY(sync_marg(e));
    node_stats_kernel<<<blocks, kThreads, smem, e->stream>>>(e->d_marg, e->d_row_ptr, true_conf ? e->d_true : nullptr,
                                                            e->N, e->Q, e->d_scratch);
    double *d_res = e->d_scratch + size_t(blocks) * kNodeCols;
    reduce_cols_kernel<<<kNodeCols, kThreads, 0, e->stream>>>(e->d_scratch, blocks, kNodeCols, d_res);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 2;
    row.assign(kNodeCols, 0.0);
    CUDA_TRY(cudaMemcpyAsync(row.data(), d_res, kNodeCols * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}

// sum of a device vector of partials, fixed order
int reduce_vector(sbmbp_engine *e, const double *d_partial, unsigned n, double *result) {
    reduce_cols_kernel<<<1, kThreads, 0, e->stream>>>(d_partial, n, 1, e->d_out);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(e->h_out, e->d_out, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    *result = e->h_out[0];
    return SBMBP_OK;
}

// Largest N whose non-edge terms are summed pair by pair (nonedge_pairs_kernel, O(N^2)); beyond it the moment series
// below.  Default 2^17; SBMBP_EXACT_PAIRS_MAX_N in the environment at create time or sbmbp_set_exact_pairs_max_n
// override it, so that the series -- the path every BASELINE size takes -- can be pinned at golden sizes.
uint32_t exact_pairs_default() {
    if (const char *env = std::getenv("SBMBP_EXACT_PAIRS_MAX_N")) return uint32_t(std::strtoul(env, nullptr, 10));
    return 1u << 17;
}

// sum over directed edges of the pair term (see nonedge_edges_kernel)
int edge_pairs_sum(sbmbp_engine *e, const double *A, const double *B, int mode, double *result,
                   const double *marg_nb = nullptr) {
    *result = 0.0;
    if (e->M == 0) return SBMBP_OK;
    TRY(ensure_col(e));
    const unsigned blocks = std::max(1u, std::min((e->N + kThreads - 1) / kThreads, unsigned(8 * e->sm_count)));
    TRY(ensure_scratch(e, blocks));
    TRY(sync_marg(e));
    nonedge_edges_kernel<<<blocks, kThreads, 2 * e->Q * e->Q * sizeof(double), e->stream>>>(
        e->d_marg, marg_nb ? marg_nb : e->d_marg, e->d_row_ptr, e->d_col, e->N, e->Q, A, B, mode, e->d_scratch);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    return reduce_vector(e, e->d_scratch, blocks, result);
}

int all_pairs_exact(sbmbp_engine *e, const double *A, const double *B, int mode, double *result) {
    const unsigned gx = (e->N + kThreads - 1) / kThreads, gy = (e->N + kPairTile - 1) / kPairTile;
    TRY(ensure_scratch(e, size_t(gx) * gy));
    const size_t smem = (size_t(kPairTile) * e->Q + 2 * e->Q * e->Q) * sizeof(double);
    static bool attr_by_device[kMaxDevices] = {};  // cudaFuncSetAttribute is per device
    bool &attr_set = attr_by_device[e->device];
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(nonedge_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int((size_t(kPairTile) * kMaxQ + 2 * kMaxQ * kMaxQ) * sizeof(double))));
        attr_set = true;
    }
    TRY(sync_marg(e));
    nonedge_pairs_kernel<<<dim3(gx, gy), kThreads, smem, e->stream>>>(e->d_marg, e->N, e->Q, A, B, mode, e->d_scratch);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    return reduce_vector(e, e->d_scratch, gx * gy, result);
}

// T_k = sum_i psi_i^{(x) k} (Q^k entries, first digit fastest); order 0: the one number sum_i log(sum_q psi_i^q)
// (see lognorm_kernel)
int moment_tensor(sbmbp_engine *e, unsigned order, std::vector<double> &T) {
    if (order == 0) {
        T.assign(1, 0.0);
        if (e->N == 0) return SBMBP_OK;
        TRY(sync_marg(e));
        const unsigned blocks = std::max(1u, std::min((e->N + kThreads - 1) / kThreads, unsigned(4 * e->sm_count)));
        TRY(ensure_scratch(e, blocks));
        lognorm_kernel<<<blocks, kThreads, 0, e->stream>>>(e->d_marg, e->N, e->Q, e->d_scratch);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
        return reduce_vector(e, e->d_scratch, blocks, T.data());
    }
    size_t len = 1;
    for (unsigned o = 0; o < order; ++o) len *= e->Q;
    T.assign(len, 0.0);
    const unsigned blocks = std::max(1u, std::min((e->N + kThreads - 1) / kThreads, unsigned(2 * e->sm_count)));
    const unsigned chunk = 16 * kThreads;
    static bool attr_by_device[kMaxDevices] = {};  // cudaFuncSetAttribute is per device
    bool &attr_set = attr_by_device[e->device];
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(size_t(kThreads) * kMaxQ * sizeof(double))));
        attr_set = true;
    }
    TRY(sync_marg(e));
    for (size_t idx0 = 0; idx0 < len; idx0 += chunk) {
        const unsigned cur = unsigned(std::min<size_t>(chunk, len - idx0));
        TRY(ensure_scratch(e, size_t(blocks) * cur + cur));
        moments_kernel<<<blocks, kThreads, size_t(kThreads) * e->Q * sizeof(double), e->stream>>>(
            e->d_marg, e->N, e->Q, order, unsigned(idx0), cur, e->d_scratch);
        double *d_res = e->d_scratch + size_t(blocks) * cur;
        reduce_cols_kernel<<<cur, kThreads, 0, e->stream>>>(e->d_scratch, blocks, cur, d_res);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 2;
        CUDA_TRY(cudaMemcpyAsync(T.data() + idx0, d_res, cur * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return SBMBP_OK;
}

// <M_0 (x) M_1 (x) ... , T (x) T> for an order-k tensor T: apply M_j along mode j, then dot with T
double contract_moments(const std::vector<double> &T, unsigned order, unsigned Q,
                        const std::vector<const double *> &mats /* each Q*Q row-major [a][b] */) {
    std::vector<double> U = T, V(T.size());
    size_t stride = 1;
    for (unsigned j = 0; j < order; ++j) {
        const double *Mj = mats[j];
        for (size_t idx = 0; idx < T.size(); ++idx) {
            const size_t dig = (idx / stride) % Q;
            const size_t base = idx - dig * stride;
            double s = 0.0;
            for (unsigned b = 0; b < Q; ++b) s += Mj[dig * Q + b] * U[base + b * stride];
            V[idx] = s;
        }
        U.swap(V);
        stride *= Q;
    }
    double r = 0.0;
    for (size_t idx = 0; idx < T.size(); ++idx) r += T[idx] * U[idx];
    return r;
}

// number of series terms for an N-node, Q-group model with pair weights |y| <= ymax: the remainder of
// sum_pairs sum_{k>K} y^k / k over N^2 pairs, divided by 2N, is driven below tol (K <= 8 and Q^K <= 2^20 entries)
unsigned series_order(uint32_t Q, double N, double ymax, double tol) {
    unsigned best = 1;
    for (unsigned K = 1; K <= 8; ++K) {
        double len = std::pow(double(Q), double(K));
        if (len > double(1u << 20)) break;
        best = K;
        const double rem = 0.5 * N * std::pow(ymax, double(K + 1)) / double(K + 1) / (1.0 - ymax);
        if (rem <= tol) break;
    }
    return best;
}

// W1_ab = 1 - W_ab with W_ab = pow(1 - c_ab / N, beta) ROUNDED TO DOUBLE as the reference evaluates it
// (belief_propagation.cpp:689): the pair weight the reference sums is that double, whose distance from 1 carries a
// relative error of ulp(1) N / c -- systematic per (a, b), so it does not average out over the pairs.  The series must
// expand the same number (the subtraction is exact), not the mathematically exact 1 - (1 - c/N)^beta (3.6e-12 apart in
// f_non_edge at N = 20 000).  Returns max |W1|.
double series_weights(uint32_t Q, double N, double beta, const double *cab, std::vector<double> &W1) {
    W1.assign(size_t(Q) * Q, 0.0);
    double ymax = 0.0;
    for (uint32_t a = 0; a < Q; ++a)
        for (uint32_t b = 0; b < Q; ++b) {
            W1[a * Q + b] = 1.0 - std::pow(1 - cab[a * Q + b] / N, beta);
            ymax = std::max(ymax, std::fabs(W1[a * Q + b]));
        }
    return ymax;
}

// compute_f_non_edge (belief_propagation.cpp:675-709)
int f_non_edge(sbmbp_engine *e, double *out) {
    *out = 0.0;
    if (e->dc != 0 || e->N == 0) return SBMBP_OK;  // :692-697: the dc branches add nothing
    const uint32_t Q = e->Q;
    double all = 0.0, edges = 0.0;
    if (e->N <= e->exact_pairs_max_n) {
        TRY(all_pairs_exact(e, e->d_prm->W, nullptr, 0, &all));
        TRY(edge_pairs_sum(e, e->d_prm->W, nullptr, 0, &edges));
    } else {
        std::vector<double> W1;
        const double ymax = series_weights(Q, double(e->N), e->beta, e->cab.data(), W1);
        const unsigned K = series_order(Q, double(e->N), ymax, 1e-14);
        for (unsigned k = 0; k <= K; ++k) {
            std::vector<double> T;
            TRY(moment_tensor(e, k, T));
            if (k == 0) {  // the normalisation defect of the marginals: 2 N sum_i log s_i
                all += 2.0 * double(e->N) * T[0];
                continue;
            }
            std::vector<const double *> mats(k, W1.data());
            all -= contract_moments(T, k, Q, mats) / double(k);
        }
        TRY(edge_pairs_sum(e, e->d_prm->W1, nullptr, 1, &edges));
    }
    *out = (all - edges) / (2.0 * double(e->N));
    return SBMBP_OK;
}

// compute_entropy_non_edge (:711-741)
int entropy_non_edge(sbmbp_engine *e, double *out) {
    *out = 0.0;
    if (e->N == 0) return SBMBP_OK;
    const uint32_t Q = e->Q;
    double all = 0.0, edges = 0.0;
    if (e->N <= e->exact_pairs_max_n) {
        TRY(all_pairs_exact(e, e->d_prm->EA, e->d_prm->EW, 2, &all));
    } else {
        // x / (1 - y) = sum_k x y^k with x = psi^T EA psi, y = psi^T (C/N) psi
        std::vector<double> A(size_t(Q) * Q), Y(size_t(Q) * Q);
        double ymax = 0.0;
        for (uint32_t a = 0; a < Q; ++a)
            for (uint32_t b = 0; b < Q; ++b) {
                const double c = e->cab[a * Q + b];
                A[a * Q + b] = c / double(e->N) * std::log(c);
                Y[a * Q + b] = c / double(e->N);
                ymax = std::max(ymax, Y[a * Q + b]);
            }
        const unsigned K = series_order(Q, double(e->N), ymax, 1e-14);
        for (unsigned k = 1; k <= K; ++k) {
            std::vector<double> T;
            TRY(moment_tensor(e, k, T));
            std::vector<const double *> mats(k, Y.data());
            mats[0] = A.data();
            all += contract_moments(T, k, Q, mats);
        }
    }
    TRY(edge_pairs_sum(e, e->d_prm->EA, e->d_prm->EW, 2, &edges));
    *out = (all - edges) / (2.0 * double(e->N));
    return SBMBP_OK;
}

void build_dev_params(const sbmbp_engine *e, DevParams &p) {
    std::memset(&p, 0, sizeof(p));
    const uint32_t Q = e->Q;
    const double N = double(e->dist ? e->N_global : e->N);
    for (uint32_t t = 0; t < Q; ++t)
        for (uint32_t q = 0; q < Q; ++q) {
            const double c = e->cab[t * Q + q];
            const int i = int(t) * kMaxQ + int(q);
            p.C[i] = c;
            p.Ks[i] = (e->dc == 0) ? std::pow(c, e->beta) : c;
            p.Kl[i] = c;
            p.P[i] = c / N;
            p.W[i] = std::pow(1 - c / N, e->beta);
            p.W1[i] = 1.0 - p.W[i];  // exact: the series expands the reference's own rounded weight (see series_weights)
            p.EA[i] = (c / N) * std::log(c);
            p.EW[i] = 1 - c / N;
        }
    for (uint32_t q = 0; q < Q; ++q) {
        p.eta[q] = e->eta[q];
        p.logeta[q] = std::log(e->eta[q]);
    }
    p.beta = e->beta;
    p.N = N;
}

int apply_params(sbmbp_engine *e) {
    DevParams p;
    build_dev_params(e, p);
    // synchronous copy from pageable memory: p may go out of scope right after
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaMemcpy(e->d_prm, &p, sizeof(p), cudaMemcpyHostToDevice));
    e->have_params = true;
    e->field_valid = false;  // h depends on c_ab
    e->state_version++;
    return SBMBP_OK;
}

template <typename T>
int import_state(sbmbp_engine *e, const double *msg, const double *marg) {
    if (msg && e->M) {
        e->compact = false;  // the new messages arrive in full storage
        e->compact_refused = false;
        const size_t n = size_t(e->M) * e->Q;
        TRY(ensure_scratch(e, n));
        CUDA_TRY(cudaMemcpyAsync(e->d_scratch, msg, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        const unsigned blocks = unsigned(std::min<size_t>((n + 255) / 256, size_t(e->sm_count) * 16));
        import_msgs_kernel<T><<<blocks, 256, 0, e->stream>>>(e->d_scratch, e->d_rev,
                                                             static_cast<T *>(e->d_S[e->sweeps_done & 1u]), e->M, e->Q);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
    }
    if (marg && e->N) {
        CUDA_TRY(cudaMemcpyAsync(e->d_marg, marg, size_t(e->N) * e->Q * sizeof(double), cudaMemcpyHostToDevice,
                                 e->stream));
        e->marg_ell_dirty = false;
    } else {
        TRY(sync_marg(e));  // the marginals stay: bring them to node order before the next sweep makes them stale
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}

template <typename T>
int export_msgs(sbmbp_engine *e, double *msg) {
    if (!e->M) return SBMBP_OK;
    TRY(ensure_full(e));
    const size_t n = size_t(e->M) * e->Q;
    TRY(ensure_scratch(e, n));
    const unsigned blocks = unsigned(std::min<size_t>((n + 255) / 256, size_t(e->sm_count) * 16));
    export_msgs_kernel<T><<<blocks, 256, 0, e->stream>>>(static_cast<const T *>(e->d_S[e->sweeps_done & 1u]), e->d_rev,
                                                         e->d_scratch, e->M, e->Q);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(msg, e->d_scratch, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}

// learning_step (belief_propagation.cpp:53-75): n_a is an unsigned int, truncated every step
void learning_step_host(sbmbp_engine *e, float learning_rate, const double *na_expect, const double *cab_expect) {
    const uint32_t Q = e->Q;
    uint32_t rest = e->N;
    for (uint32_t i = 0; i + 1 < Q; ++i) {
        e->na[i] = unsigned(int(learning_rate * na_expect[i] + (1.0 - learning_rate) * e->na[i]));
        rest -= e->na[i];
    }
    e->na[Q - 1] = rest;
    for (uint32_t i = 0; i < Q; ++i) {
        e->eta[i] = double(e->na[i]) / e->N;
        for (uint32_t j = 0; j < Q; ++j)
            e->cab[i * Q + j] = learning_rate * cab_expect[i * Q + j] + (1.0 - learning_rate) * e->cab[i * Q + j];
    }
}

}  // namespace

// =============================================================================================== C ABI

extern "C" {

const char *sbmbp_version(void) { return "sbmbp-b200 0.1 (sm_100a)"; }
const char *sbmbp_last_error(void) { return get_error(); }

int sbmbp_graph_from_pairs(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N, sbmbp_graph **g) {
    if (!g || (n_pairs && (!u || !v))) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    auto *gr = new sbmbp_graph();
    int rc = build_graph(u, v, n_pairs, N, *gr);
    if (rc != SBMBP_OK) {
        delete gr;
        return rc;
    }
    *g = gr;
    return SBMBP_OK;
}

int sbmbp_graph_from_edgelist(const char *path, uint32_t N, sbmbp_graph **g) {
    if (!path || !g) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    std::vector<uint32_t> u, v;
    TRY(parse_edgelist(path, u, v));
    return sbmbp_graph_from_pairs(u.data(), v.data(), u.size(), N, g);
}

int sbmbp_parse_edgelist(const char *path, uint32_t *u, uint32_t *v, uint64_t cap, uint64_t *n) {
    if (!path || !n) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    std::vector<uint32_t> uu, vv;
    TRY(parse_edgelist(path, uu, vv));
    *n = uu.size();
    for (uint64_t k = 0; k < uu.size() && k < cap; ++k) {
        if (u) u[k] = uu[k];
        if (v) v[k] = vv[k];
    }
    return SBMBP_OK;
}

int sbmbp_ell_layout(const sbmbp_graph *g, uint64_t region_slots, uint32_t *pos, uint32_t *gather, uint32_t *classes,
                     uint32_t cls_cap, uint32_t *n_cls, uint32_t *node, uint32_t *n_node, uint32_t *rev_idx,
                     uint32_t *pos_idx, uint64_t idx_cap, uint64_t *n_idx, uint32_t *n_chunks, uint32_t *n_buckets) {
    if (!g) {
        set_error("null graph");
        return SBMBP_ERR_ARG;
    }
    std::vector<unsigned> vpos, vgather, vnode, vrev, vpidx;
    std::vector<EllClass> cls;
    unsigned nchunks = 0;
    const unsigned nb = build_bell_layout(*g, region_slots, vpos, vgather, cls, vnode, vrev, vpidx, nchunks);
    if (pos) std::copy(vpos.begin(), vpos.end(), pos);
    if (gather) std::copy(vgather.begin(), vgather.end(), gather);
    if (classes)
        for (size_t i = 0; i < cls.size() && i < cls_cap; ++i) {
            classes[5 * i + 0] = cls[i].d;
            classes[5 * i + 1] = cls[i].n;
            classes[5 * i + 2] = cls[i].node_first;
            classes[5 * i + 3] = cls[i].chunk_first;
            classes[5 * i + 4] = cls[i].base;
        }
    if (n_cls) *n_cls = uint32_t(cls.size());
    if (node) std::copy(vnode.begin(), vnode.end(), node);
    if (n_node) *n_node = uint32_t(vnode.size());
    if (rev_idx) std::copy(vrev.begin(), vrev.begin() + std::min<uint64_t>(vrev.size(), idx_cap), rev_idx);
    if (pos_idx) std::copy(vpidx.begin(), vpidx.begin() + std::min<uint64_t>(vpidx.size(), idx_cap), pos_idx);
    if (n_idx) *n_idx = vrev.size();
    if (n_chunks) *n_chunks = nchunks;
    if (n_buckets) *n_buckets = nb;
    {   // self-check of the work lists the kernel would walk: every chunk exactly once, on some warp
        std::vector<uint4> sched;
        unsigned len = 0;
        const unsigned nwarps = 37;
        build_ell_schedule(cls, nwarps, 8, sched, len);
        std::vector<unsigned> seen(nchunks, 0);
        uint64_t lanes = 0;
        for (const uint4 &s : sched) {
            if ((s.z >> 8) == 0) continue;
            if (s.w >= nchunks) {
                set_error("ELL schedule: chunk id out of range");
                return SBMBP_ERR_STATE;
            }
            seen[s.w]++;
            lanes += s.z >> 8;
        }
        for (unsigned k = 0; k < nchunks; ++k)
            if (seen[k] != 1) {
                set_error("ELL schedule: a chunk is missing or duplicated");
                return SBMBP_ERR_STATE;
            }
        if (lanes != vnode.size() || sched.size() != size_t(nwarps) * len) {
            set_error("ELL schedule: lane count mismatch");
            return SBMBP_ERR_STATE;
        }
    }
    return SBMBP_OK;
}

int sbmbp_debug_trace(sbmbp_engine *e, uint64_t *out, uint64_t cap_words, uint64_t *n_words) {
    if (!e || !n_words) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    const uint64_t n = e->d_trace ? uint64_t(e->trace_warps) * 16 : 0;
    *n_words = n;
    if (out && n) {
        CUDA_TRY(cudaSetDevice(e->device));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        CUDA_TRY(cudaMemcpy(out, e->d_trace, std::min(n, cap_words) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    return SBMBP_OK;
}

int sbmbp_graph_destroy(sbmbp_graph *g) {
    delete g;
    return SBMBP_OK;
}

int sbmbp_graph_info(const sbmbp_graph *g, uint32_t *N, uint64_t *M, uint64_t *E, uint32_t *max_degree) {
    if (!g) {
        set_error("null graph");
        return SBMBP_ERR_ARG;
    }
    if (N) *N = g->N;
    if (M) *M = g->M;
    if (E) *E = g->E;
    if (max_degree) *max_degree = g->max_degree;
    return SBMBP_OK;
}

int sbmbp_graph_csr(const sbmbp_graph *g, const uint64_t **row_ptr, const uint32_t **col, const uint32_t **rev,
                    const uint32_t **deg) {
    if (!g) {
        set_error("null graph");
        return SBMBP_ERR_ARG;
    }
    if (row_ptr) *row_ptr = g->row_ptr.data();
    if (col) *col = g->col.data();
    if (rev) *rev = g->rev.data();
    if (deg) *deg = g->deg.data();
    return SBMBP_OK;
}

// bp_param_from_direct (blockmodel.cpp:274-302): na[q] = unsigned(int(pa[q] * N)) for EVERY q -- the remainder
// fix-up of :282-284 is overwritten at :286 -- and --cab is the upper triangle in row-major order (:295-297)
int sbmbp_params_from_direct(uint32_t N, uint32_t Q, const double *pa, const double *cab_upper, uint32_t *na,
                             double *cab) {
    if (!pa || !cab_upper || !na || !cab || Q == 0) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    for (uint32_t q = 0; q < Q; ++q) na[q] = unsigned(int(pa[q] * N));
    for (uint32_t q = 0; q < Q; ++q) {
        const uint32_t base = q * Q - q * (q - 1) / 2;
        cab[q * Q + q] = cab_upper[base];
        for (uint32_t t = q + 1; t < Q; ++t) {
            cab[q * Q + t] = cab_upper[base + t - q];
            cab[t * Q + q] = cab[q * Q + t];
        }
    }
    return SBMBP_OK;
}

// bp_param_from_epsilon_c (blockmodel.cpp:229-272)
int sbmbp_params_from_epsilon_c(uint32_t N, uint32_t Q, double epsilon, double c, uint32_t *na, double *cab) {
    if (!na || !cab || Q == 0) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    for (uint32_t q = 0; q < Q; ++q) {
        const double pa = 1.0 / Q;
        na[q] = unsigned(int(pa * N));
    }
    double cin, co;
    if (epsilon < 0) {
        cin = 0;
        co = c * Q / (Q - 1);
    } else {
        cin = c * Q / ((Q - 1) * epsilon + 1);
        co = epsilon * cin;
    }
    for (uint32_t q = 0; q < Q; ++q) {
        cab[q * Q + q] = cin;
        for (uint32_t t = q + 1; t < Q; ++t) cab[q * Q + t] = cab[t * Q + q] = co;
    }
    return SBMBP_OK;
}

int sbmbp_create(const sbmbp_graph *g, uint32_t Q, uint32_t deg_corr_flag, int precision, int device,
                 sbmbp_engine **out) {
    if (!g || !out) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    if (Q < 1 || Q > SBMBP_MAX_Q) {
        set_error("Q must be in [1, " + std::to_string(SBMBP_MAX_Q) + "]");
        return SBMBP_ERR_UNSUPPORTED;
    }
    if (deg_corr_flag > 2 || (precision != SBMBP_F64 && precision != SBMBP_F32)) {
        set_error("bad deg_corr_flag or precision");
        return SBMBP_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the engine has no CPU fallback");
        return SBMBP_ERR_NODEVICE;
    }
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= ndev || device >= kMaxDevices) {
        set_error("device index out of range");
        return SBMBP_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    {
        // The sweep's only random access is one 16/32-byte message gather per edge.  With the default L2 fetch
        // granularity every such miss drags a full 128-byte line out of HBM (measured: 138 B read per edge update
        // at Q=2 FP64 against 53 B needed); 32-byte sectors cut that traffic by more than half.
        int gran = 32;
        if (const char *env = std::getenv("SBMBP_L2_FETCH")) gran = std::atoi(env);
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(gran));
        cudaGetLastError();
    }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("device '") + prop.name + "' is not sm_100: this library carries sm_100a code only");
        return SBMBP_ERR_NODEVICE;
    }
    auto *e = new sbmbp_engine();
    e->g = g;
    e->N = g->N;
    e->M = g->M;
    e->Q = Q;
    e->dc = deg_corr_flag;
    e->prec = precision;
    e->qt = pick_qt(Q);
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->exact_pairs_max_n = exact_pairs_default();
    if (const char *env = std::getenv("SBMBP_GATHER_MODE")) e->gather_mode = std::atoi(env);
    e->fast_path = true;
    if (const char *env = std::getenv("SBMBP_NO_FAST")) e->fast_path = std::atoi(env) == 0;
    if (const char *env = std::getenv("SBMBP_NO_PIPE")) e->pipe_path = std::atoi(env) == 0;
    int te = 0, tn = 0;
    dispatch(e, [&](auto t, auto qt) {
        tile_geometry<decltype(t), decltype(qt)::value>(te, tn);
        return SBMBP_OK;
    });
    std::vector<Tile> tiles = make_tiles(*g, te, tn);
    e->ntiles = unsigned(tiles.size());
    const size_t elt = (precision == SBMBP_F64) ? 8 : 4;
    auto fail = [&](int rc) {
        sbmbp_destroy(e);
        return rc;
    };
#define CREATE_TRY(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _err = (expr);                                                           \
        if (_err != cudaSuccess) {                                                           \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_err));                 \
            return fail(SBMBP_ERR_CUDA);                                                     \
        }                                                                                    \
    } while (0)
    CREATE_TRY(cudaMalloc(&e->d_row_ptr, (size_t(e->N) + 1) * sizeof(unsigned long long) + kBulkPad));
    CREATE_TRY(cudaMalloc(&e->d_rev, std::max<size_t>(e->M, 1) * sizeof(unsigned) + kBulkPad));
    CREATE_TRY(cudaMalloc(&e->d_marg, std::max<size_t>(size_t(e->N) * Q, 1) * sizeof(double)));
    CREATE_TRY(cudaMalloc(&e->d_tiles, std::max<size_t>(e->ntiles, 1) * sizeof(Tile)));
    CREATE_TRY(cudaMalloc(&e->d_prm, sizeof(DevParams)));
    CREATE_TRY(cudaMalloc(&e->d_field[0], sizeof(Field)));
    CREATE_TRY(cudaMalloc(&e->d_field[1], sizeof(Field)));
    CREATE_TRY(cudaMalloc(&e->d_ctl, sizeof(Ctl)));
    // one row per tile (general kernel) or per CTA (persistent kernels; the warp path adds up to 4 hub CTAs per SM)
    CREATE_TRY(cudaMalloc(&e->d_partial,
                          std::max<size_t>(size_t(e->ntiles) + size_t(prop.multiProcessorCount) * 48, 1) * (e->qt + 1) * sizeof(double)));
    CREATE_TRY(cudaMalloc(&e->d_out, kOutDoubles * sizeof(double)));
    CREATE_TRY(cudaMallocHost(&e->h_ctl, sizeof(Ctl)));
    CREATE_TRY(cudaMallocHost(&e->h_out, kOutDoubles * sizeof(double)));
    CREATE_TRY(cudaEventCreate(&e->ev0));
    CREATE_TRY(cudaEventCreate(&e->ev1));
    CREATE_TRY(cudaMemcpy(e->d_row_ptr, g->row_ptr.data(), (size_t(e->N) + 1) * sizeof(unsigned long long),
                          cudaMemcpyHostToDevice));
    {
        // region size of the bucketed layout: a fraction of the 126 MB L2 (SBMBP_REGION_MB while tuning; 0 = one bucket)
        double region_mb = 16.0;
        // very large graphs (configs[3] at its stated 100M nodes on one GPU: 16 GB per buffer): with 16 MiB regions a tile's
        // out-messages scatter over ~1000 buckets and no two of them share a line; keep the bucket count near 128
        // (the multi-GPU plan makes the same trade, sbmbp_plan_create)
        {
            const double buf_mb = double(e->M) * Q * elt / 1048576.0;
            if (buf_mb / region_mb > 128.0) region_mb = std::min(128.0, buf_mb / 128.0);
        }
        if (const char *env = std::getenv("SBMBP_REGION_MB")) region_mb = std::atof(env);
        const uint64_t region_slots = uint64_t(region_mb * 1048576.0 / double(Q * elt));
        std::vector<unsigned> pos, gather;
        // Small Q, sparse, message buffers up to a few hundred MB: degree-class layout + bp_sweep_ell_kernel.  Otherwise the
        // destination-bucketed layout of the tile kernels (SBMBP_ELL_MAX_MB: largest single buffer that takes the ELL path,
        // SBMBP_ELL_MIN_LOW: smallest share of edge slots on unrolled degrees).
        e->wide_path = e->qt == 32 && Q == 32 && e->dc != 2 && e->N > 0;
        if (const char *env = std::getenv("SBMBP_NO_WIDE")) e->wide_path = e->wide_path && std::atoi(env) == 0;
        const bool small_q = (e->qt <= 4) && e->Q == uint32_t(e->qt) && e->dc != 2 && e->N > 0;
        // ... measured on the configs[1] family (c = 3): ELL beats the CTA-tile kernel from 1M to 8M nodes (24 regions of
        // 16 MiB: 0.54-0.57 of the roofline against 0.48-0.51) and loses from 16M on (48 regions: the 32 out-messages of a
        // (chunk, slot) no longer share lines: 1.25 ms against 0.86 ms); beyond the unrolled degrees it gathers twice, so
        // it also needs most edge slots to sit on nodes of degree <= 8 (configs[3], c = 10: 5.3 ms against 2.1 ms)
        double ell_max_mb = 400.0;
        if (const char *env = std::getenv("SBMBP_ELL_MAX_MB")) ell_max_mb = std::atof(env);
        uint64_t low_slots = 0;
        const uint32_t du = (Q * elt <= 16) ? 8u : 4u;  // EllUnroll<T, QT>::DU
        for (uint32_t i = 0; i < g->N; ++i)
            if (g->deg[i] <= du) low_slots += g->deg[i];
        double min_low = 0.6;
        if (const char *env = std::getenv("SBMBP_ELL_MIN_LOW")) min_low = std::atof(env);
        e->ell_path = small_q && double(e->M) * Q * elt <= ell_max_mb * 1048576.0 &&
                      double(low_slots) >= min_low * double(e->M);
        if (const char *env = std::getenv("SBMBP_NO_ELL")) e->ell_path = e->ell_path && std::atoi(env) == 0;
        e->warp_path = e->ell_path;  // degrees >= 32 of the ELL path
        if (const char *env = std::getenv("SBMBP_WARP_MAIN")) e->warp_path = e->warp_path || (small_q && std::atoi(env) != 0);
        // ... and when a buffer fits the L2 with room to spare (SBMBP_ELL_PADDED_MAX_MB, default 64), the one-bucket padded variant of that layout (build_ell_padded_layout): no pos words, contiguous old / new blocks
        double padded_max_mb = 64.0;
        if (const char *env = std::getenv("SBMBP_ELL_PADDED_MAX_MB")) padded_max_mb = std::atof(env);
        e->ell_padded = e->ell_path && double(e->M) * Q * elt <= padded_max_mb * 1048576.0;
        if (const char *env = std::getenv("SBMBP_NO_ELL_PADDED")) e->ell_padded = e->ell_padded && std::atoi(env) == 0;
        e->buf_slots = std::max<uint64_t>(e->M, 1);
        if (e->ell_path) {
            std::vector<EllClass> cls;
            std::vector<unsigned> ell_node, ell_rev, ell_pos;
            if (e->ell_padded) {
                build_ell_padded_layout(*g, pos, gather, cls, ell_node, ell_rev, e->ell_nchunks, e->buf_slots);
                e->nbuckets = 1;
                e->ell_entries = unsigned(ell_node.size());
                CREATE_TRY(cudaMalloc(&e->d_marg_ell, std::max<size_t>(ell_node.size(), 1) * Q * sizeof(double)));
                CREATE_TRY(cudaMemset(e->d_marg_ell, 0, std::max<size_t>(ell_node.size(), 1) * Q * sizeof(double)));
            } else {
                e->nbuckets = build_bell_layout(*g, region_slots, pos, gather, cls, ell_node, ell_rev, ell_pos, e->ell_nchunks);
            }
            e->ell_ncls = unsigned(cls.size());
            e->ell_nidx = ell_rev.size();
            CREATE_TRY(cudaMalloc(&e->d_ell_cls, std::max<size_t>(cls.size(), 1) * sizeof(EllClass)));
            CREATE_TRY(cudaMalloc(&e->d_ell_node, std::max<size_t>(ell_node.size(), 1) * sizeof(unsigned)));
            CREATE_TRY(cudaMalloc(&e->d_ell_rev, std::max<size_t>(ell_rev.size(), 1) * sizeof(unsigned)));
            CREATE_TRY(cudaMalloc(&e->d_ell_pos, std::max<size_t>(ell_pos.size(), 1) * sizeof(unsigned)));
            if (!cls.empty()) CREATE_TRY(cudaMemcpy(e->d_ell_cls, cls.data(), cls.size() * sizeof(EllClass), cudaMemcpyHostToDevice));
            if (!ell_node.empty())
                CREATE_TRY(cudaMemcpy(e->d_ell_node, ell_node.data(), ell_node.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
            if (!ell_rev.empty())
                CREATE_TRY(cudaMemcpy(e->d_ell_rev, ell_rev.data(), ell_rev.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
            if (!ell_pos.empty())
                CREATE_TRY(cudaMemcpy(e->d_ell_pos, ell_pos.data(), ell_pos.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
            {
                int ctas = 1, du = 4, wpc = 8;
                dispatch(e, [&](auto t, auto qt) { return ell_kernel_config<decltype(t), decltype(qt)::value>(&ctas, &du, &wpc); });
                e->ell_wpc = unsigned(wpc);
                e->ell_grid = std::max(1u, std::min<unsigned>((e->ell_nchunks + e->ell_wpc - 1u) / e->ell_wpc, unsigned(ctas) * unsigned(e->sm_count)));
                std::vector<uint4> sched;
                build_ell_schedule(cls, e->ell_grid * e->ell_wpc, unsigned(du), sched, e->ell_sched_len);
                CREATE_TRY(cudaMalloc(&e->d_ell_sched, std::max<size_t>(sched.size(), 1) * sizeof(uint4)));
                if (!sched.empty())
                    CREATE_TRY(cudaMemcpy(e->d_ell_sched, sched.data(), sched.size() * sizeof(uint4), cudaMemcpyHostToDevice));
            }
            if (const char *env = std::getenv("SBMBP_ELL_TRACE")) {
                if (std::atoi(env)) {
                    e->trace_warps = e->ell_grid * e->ell_wpc;
                    CREATE_TRY(cudaMalloc(&e->d_trace, size_t(e->trace_warps) * 16 * sizeof(unsigned long long)));
                    CREATE_TRY(cudaMemset(e->d_trace, 0, size_t(e->trace_warps) * 16 * sizeof(unsigned long long)));
                }
            }
        } else {
            // wide Q: a message is one or two full lines, the gather needs no bucketing -> slot order
            e->nbuckets = build_layout(*g, e->wide_path ? 0 : region_slots, pos, gather);
        }
        if (e->wide_path) {
            std::vector<unsigned> wn;
            std::vector<Tile> bt;
            for (uint32_t i = 0; i < g->N; ++i) {
                if (g->deg[i] <= 32u) {
                    wn.push_back(i);
                } else {
                    Tile t;
                    t.e0 = g->row_ptr[i];
                    t.n0 = i;
                    t.nn = 1;
                    t.ne = g->deg[i];
                    t.nbig = 1;
                    bt.push_back(t);
                }
            }
            e->n_wide_nodes = unsigned(wn.size());
            e->nbtiles = unsigned(bt.size());
            CREATE_TRY(cudaMalloc(&e->d_wide_nodes, std::max<size_t>(wn.size(), 1) * sizeof(unsigned)));
            if (!wn.empty()) CREATE_TRY(cudaMemcpy(e->d_wide_nodes, wn.data(), wn.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
            if (!bt.empty()) {
                std::vector<unsigned> bpos = pos, binfo;  // slot order = buffer order here
                sort_tile_positions(*g, bt, te, bpos, binfo);
                CREATE_TRY(cudaMalloc(&e->d_btiles, bt.size() * sizeof(Tile)));
                CREATE_TRY(cudaMalloc(&e->d_bpos, e->M * sizeof(unsigned) + kBulkPad));
                CREATE_TRY(cudaMalloc(&e->d_binfo, e->M * sizeof(unsigned) + kBulkPad));
                CREATE_TRY(cudaMemcpy(e->d_btiles, bt.data(), bt.size() * sizeof(Tile), cudaMemcpyHostToDevice));
                CREATE_TRY(cudaMemcpy(e->d_bpos, bpos.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
                CREATE_TRY(cudaMemcpy(e->d_binfo, binfo.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
            }
        }
        // message buffers: one slot per directed edge, plus the lane padding of the padded degree-class layout; padding is zero-filled
        // once and never read by any kernel (bulk stores may overwrite it)
        for (int b = 0; b < 2; ++b) {
            CREATE_TRY(cudaMalloc(&e->d_S[b], e->buf_slots * Q * elt));
            CREATE_TRY(cudaMemset(e->d_S[b], 0, e->buf_slots * Q * elt));
        }
        // Compact storage of the sweeps' messages: Q = 2 in FP64 with every node on the degree-class kernel and the padded
        // layout (SBMBP_COMPACT=0 turns it off).  The lane padding of the full buffers is set to (1/2, 1/2) so that the
        // normalisation check of ensure_compact sees only genuine messages fail.
        int want_compact = 1;
        if (const char *env = std::getenv("SBMBP_COMPACT")) want_compact = std::atoi(env);
        bool no_big = true;
        for (uint32_t i = 0; i < g->N && no_big; ++i) no_big = g->deg[i] < kEllDegrees;
        e->compact_ok = want_compact && e->ell_padded && Q == 2 && precision == SBMBP_F64 && no_big;
        if (e->compact_ok) {
            std::vector<double> half(size_t(e->buf_slots) * 2, 0.5);
            for (int b = 0; b < 2; ++b) {
                CREATE_TRY(cudaMemcpy(e->d_S[b], half.data(), half.size() * sizeof(double), cudaMemcpyHostToDevice));
                CREATE_TRY(cudaMalloc(&e->d_C[b], e->buf_slots * sizeof(double)));
                CREATE_TRY(cudaMemset(e->d_C[b], 0, e->buf_slots * sizeof(double)));
            }
        }
        if (e->M) CREATE_TRY(cudaMemcpy(e->d_rev, gather.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
        if (e->warp_path) {
            const int we = (Q * elt <= 16) ? 128 : 64;  // WarpCfg<T, QT>::WE
            std::vector<WTile> wt;
            std::vector<Tile> hubs;
            make_wtiles(*g, we, e->ell_path ? kEllDegrees : 0u, wt, hubs);
            std::vector<unsigned> wpos = pos;  // slot order
            std::vector<unsigned short> winfo;
            sort_wtile_positions(*g, wt, wpos, winfo);
            e->nwtiles = unsigned(wt.size());
            e->nhubs = unsigned(hubs.size());
            CREATE_TRY(cudaMalloc(&e->d_wtiles, std::max<size_t>(wt.size(), 1) * sizeof(WTile)));
            CREATE_TRY(cudaMalloc(&e->d_hubs, std::max<size_t>(hubs.size(), 1) * sizeof(Tile)));
            CREATE_TRY(cudaMalloc(&e->d_wpos, std::max<size_t>(e->M, 1) * sizeof(unsigned)));
            CREATE_TRY(cudaMalloc(&e->d_winfo, std::max<size_t>(e->M, 1) * sizeof(unsigned short)));
            if (!wt.empty()) CREATE_TRY(cudaMemcpy(e->d_wtiles, wt.data(), wt.size() * sizeof(WTile), cudaMemcpyHostToDevice));
            if (!hubs.empty()) CREATE_TRY(cudaMemcpy(e->d_hubs, hubs.data(), hubs.size() * sizeof(Tile), cudaMemcpyHostToDevice));
            if (e->M) {
                CREATE_TRY(cudaMemcpy(e->d_wpos, wpos.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
                CREATE_TRY(cudaMemcpy(e->d_winfo, winfo.data(), e->M * sizeof(unsigned short), cudaMemcpyHostToDevice));
            }
        }
        if (!pos.empty()) {
            std::vector<unsigned> info;
            sort_tile_positions(*g, tiles, te, pos, info);
            CREATE_TRY(cudaMalloc(&e->d_info, e->M * sizeof(unsigned) + kBulkPad));
            CREATE_TRY(cudaMemcpy(e->d_info, info.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc(&e->d_pos, e->M * sizeof(unsigned) + kBulkPad));
            CREATE_TRY(cudaMemcpy(e->d_pos, pos.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
        }
    }
    if (e->ntiles)
        CREATE_TRY(cudaMemcpy(e->d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemset(e->d_ctl, 0, sizeof(Ctl)));
    CREATE_TRY(cudaMemset(e->d_field[0], 0, sizeof(Field)));
    CREATE_TRY(cudaMemset(e->d_field[1], 0, sizeof(Field)));
    if (e->dc != 0 && e->M) {
        if (ensure_col(e) != SBMBP_OK) return fail(SBMBP_ERR_CUDA);
        CREATE_TRY(cudaMalloc(&e->d_degsrc, e->M * sizeof(unsigned)));
        const unsigned blocks = unsigned(std::min<uint64_t>((e->M + 255) / 256, uint64_t(e->sm_count) * 16));
        degsrc_kernel<<<blocks, 256, 0, e->stream>>>(e->d_row_ptr, e->d_col, e->d_degsrc, e->M);
        CREATE_TRY(cudaGetLastError());
        CREATE_TRY(cudaStreamSynchronize(e->stream));
    }
#undef CREATE_TRY
    e->na.assign(Q, 0);
    e->cab.assign(size_t(Q) * Q, 0.0);
    e->eta.assign(Q, 0.0);
    *out = e;
    return SBMBP_OK;
}

int sbmbp_destroy(sbmbp_engine *e) {
    if (!e) return SBMBP_OK;
    cudaSetDevice(e->device);
    cudaFree(e->d_row_ptr);
    cudaFree(e->d_rev);
    cudaFree(e->d_pos);
    cudaFree(e->d_info);
    cudaFree(e->d_col);
    cudaFree(e->d_degsrc);
    cudaFree(e->d_true);
    cudaFree(e->d_S[0]);
    cudaFree(e->d_S[1]);
    cudaFree(e->d_marg);
    cudaFree(e->d_tiles);
    cudaFree(e->d_wtiles);
    cudaFree(e->d_hubs);
    cudaFree(e->d_wpos);
    cudaFree(e->d_winfo);
    cudaFree(e->d_clamp);
    cudaFree(e->d_color);
    cudaFree(e->d_rp_rev);
    cudaFree(e->d_rp_degn);
    cudaFree(e->d_rp_sched);
    cudaFree(e->d_rp_h);
    cudaFree(e->d_rp_scratch);
    cudaFree(e->d_wide_nodes);
    cudaFree(e->d_btiles);
    cudaFree(e->d_bpos);
    cudaFree(e->d_binfo);
    cudaFree(e->d_ell_cls);
    cudaFree(e->d_ell_node);
    cudaFree(e->d_ell_rev);
    cudaFree(e->d_ell_pos);
    cudaFree(e->d_marg_ell);
    cudaFree(e->d_C[0]);
    cudaFree(e->d_C[1]);
    cudaFree(e->d_trace);
    cudaFree(e->d_ell_sched);
    cudaFree(e->d_prm);
    cudaFree(e->d_field[0]);
    cudaFree(e->d_field[1]);
    cudaFree(e->d_ctl);
    cudaFree(e->d_partial);
    cudaFree(e->d_scratch);
    cudaFree(e->d_out);
    cudaFree(e->d_mirror);
    cudaFree(e->d_rpos);
    cudaFree(e->d_out_start);
    cudaFree(e->d_out_rpos);
    cudaFree(e->d_ship);
    cudaFree(e->d_ship_start);
    cudaFree(e->d_sync);
    cudaFree(e->d_row);
    for (void *ptr : e->ipc_opened) cudaIpcCloseMemHandle(ptr);
    if (e->h_ctl) cudaFreeHost(e->h_ctl);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    delete e;
    return SBMBP_OK;
}

int sbmbp_set_stream(sbmbp_engine *e, void *cuda_stream) {
    TRY(need(e, false, false));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->stream = static_cast<cudaStream_t>(cuda_stream);
    return SBMBP_OK;
}

int sbmbp_set_params(sbmbp_engine *e, const uint32_t *na, const double *cab, double beta) {
    TRY(need(e, false, false));
    if (!na || !cab) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    e->na.assign(na, na + e->Q);
    e->cab.assign(cab, cab + size_t(e->Q) * e->Q);
    for (uint32_t q = 0; q < e->Q; ++q) e->eta[q] = 1.0 * e->na[q] / (e->dist ? e->N_global : e->N);  // belief_propagation.cpp:307
    e->beta = beta;
    return apply_params(e);
}

int sbmbp_get_params(sbmbp_engine *e, uint32_t *na, double *cab, double *eta) {
    TRY(need(e, true, false));
    if (na) std::copy(e->na.begin(), e->na.end(), na);
    if (cab) std::copy(e->cab.begin(), e->cab.end(), cab);
    if (eta) std::copy(e->eta.begin(), e->eta.end(), eta);
    return SBMBP_OK;
}

int sbmbp_set_state(sbmbp_engine *e, const double *msg, const double *marg) {
    TRY(need(e, false, false));
    if (e->prec == SBMBP_F64) TRY(import_state<double>(e, msg, marg));
    else TRY(import_state<float>(e, msg, marg));
    e->have_state = true;
    e->field_valid = false;
    e->state_version++;
    return SBMBP_OK;
}

// init_messages flag 0 (belief_propagation.cpp:110-131), draw for draw: per node Q uniforms for the marginal, then
// per neighbour (ascending) Q uniforms for the OUTGOING message, each normalised.  The outgoing message of slot
// e = (i, l) is stored by the reference at mmap_[j][idx_ji], i.e. at reference position rev[e].
static int init_random_from(sbmbp_engine *e, std::mt19937 &engine) {
    TRY(need(e, false, false));
    if (e->dist) {
        set_error("single-GPU entry point called on a multi-GPU engine (use the sbmbp_dist_* calls)");
        return SBMBP_ERR_STATE;
    }
    const uint32_t Q = e->Q;
    std::uniform_real_distribution<> random_real(0, 1);
    std::vector<double> msg(size_t(e->M) * Q), marg(size_t(e->N) * Q);
    const auto &g = *e->g;
    for (uint32_t i = 0; i < e->N; ++i) {
        double norm = 0.0;
        for (uint32_t q = 0; q < Q; ++q) {
            marg[size_t(i) * Q + q] = random_real(engine);
            norm += marg[size_t(i) * Q + q];
        }
        for (uint32_t q = 0; q < Q; ++q) marg[size_t(i) * Q + q] /= norm;
        for (uint64_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) {
            double *slot = msg.data() + size_t(g.rev[s]) * Q;
            norm = 0.0;
            for (uint32_t q = 0; q < Q; ++q) {
                slot[q] = random_real(engine);
                norm += slot[q];
            }
            for (uint32_t q = 0; q < Q; ++q) slot[q] /= norm;
        }
    }
    e->rng = engine;  // converge() goes on drawing from the same generator (main.cpp:338,362)
    e->rng_valid = true;
    return sbmbp_set_state(e, msg.data(), marg.data());
}

int sbmbp_init_random(sbmbp_engine *e, uint32_t seed) {
    std::mt19937 engine(seed);
    return init_random_from(e, engine);
}

int sbmbp_init_random_device(sbmbp_engine *e, uint64_t seed) {
    TRY(need(e, false, false));
    // every slot of the buffer, lane padding of the padded degree-class layout included (padding is never read; it just stays finite)
    const uint64_t slots = e->M ? e->buf_slots : 0;
    const uint64_t total = slots + e->N;
    const unsigned blocks = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, uint64_t(e->sm_count) * 16)));
    if (e->prec == SBMBP_F64)
        random_init_kernel<double><<<blocks, 256, 0, e->stream>>>(static_cast<double *>(e->d_S[e->sweeps_done & 1u]),
                                                                 e->d_marg, slots, e->N, e->Q, seed);
    else
        random_init_kernel<float><<<blocks, 256, 0, e->stream>>>(static_cast<float *>(e->d_S[e->sweeps_done & 1u]),
                                                                e->d_marg, slots, e->N, e->Q, seed);
    e->marg_ell_dirty = false;
    e->compact = false;  // the new state is in full storage
    e->compact_refused = false;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->stat_launches += 1;
    e->have_state = true;
    e->field_valid = false;
    e->state_version++;
    return SBMBP_OK;
}

// init_messages, all four flags (belief_propagation.cpp:101-215), draw for draw with std::mt19937(seed).  conf[N] is the
// beliefs vector main.cpp builds from --beliefs_path / -f (:325-336); -1 = unknown.  Quirks kept on purpose: flag 2
// writes the node's own in-slots, un-normalised, with q as the outer loop and a float noise constant; flag 3 advances
// its neighbour index twice per turn, so only even-ranked neighbours receive the planted message and the other slots
// stay zero.  Flags 2 and 3 assert(conf != 1) in the reference (:179,:197): reported as SBMBP_ERR_UNSUPPORTED.
static int init_messages_from(sbmbp_engine *e, uint32_t flag, const int32_t *conf, std::mt19937 &engine) {
    TRY(need(e, false, false));
    if (flag == 0) {
        e->conf_planted.clear();
        e->n_planted = 0;
        return init_random_from(e, engine);
    }
    if (e->dist) {
        set_error("single-GPU entry point called on a multi-GPU engine");
        return SBMBP_ERR_STATE;
    }
    if (flag > 3 || !conf) {
        set_error("bp_messages_init_flag must be 0..3 and flags 1-3 need a beliefs vector");
        return SBMBP_ERR_ARG;
    }
    const uint32_t Q = e->Q, N = e->N;
    for (uint32_t i = 0; i < N; ++i) {
        if (conf[i] < -1 || conf[i] >= int32_t(Q)) {
            set_error("belief out of range (must be -1 or a group index)");
            return SBMBP_ERR_ARG;
        }
        if (flag >= 2 && conf[i] == 1) {
            set_error("bp_messages_init_flag 2/3 with a belief equal to 1: the reference aborts on its assert "
                      "(belief_propagation.cpp:179,:197)");
            return SBMBP_ERR_UNSUPPORTED;
        }
    }
    std::uniform_real_distribution<> random_real(0, 1);
    std::vector<double> msg(size_t(e->M) * Q, 0.0), marg(size_t(N) * Q, 0.0);
    const auto &g = *e->g;
    const float planted_noise = 0.1;  // :177
    for (uint32_t i = 0; i < N; ++i) {
        double *mp = marg.data() + size_t(i) * Q;
        const uint64_t r0 = g.row_ptr[i];
        const uint32_t d = g.deg[i];
        if (flag == 1) {
            double norm = 0.0;
            if (conf[i] != -1) {
                for (uint32_t q = 0; q < Q; ++q) mp[q] = (q == uint32_t(conf[i])) ? 1.0 : 0.0;
            } else {
                for (uint32_t q = 0; q < Q; ++q) {
                    mp[q] = random_real(engine);
                    norm += mp[q];
                }
                for (uint32_t q = 0; q < Q; ++q) mp[q] /= norm;
            }
            for (uint32_t l = 0; l < d; ++l) {
                double *slot = msg.data() + size_t(g.rev[r0 + l]) * Q;  // the outgoing message lives in the neighbour's in-slot
                if (conf[i] != -1) {
                    for (uint32_t q = 0; q < Q; ++q) slot[q] = (q == uint32_t(conf[i])) ? 1.0 : 0.0;
                } else {
                    norm = 0.0;
                    for (uint32_t q = 0; q < Q; ++q) {
                        slot[q] = random_real(engine);
                        norm += slot[q];
                    }
                    for (uint32_t q = 0; q < Q; ++q) slot[q] /= norm;
                }
            }
        } else if (flag == 2) {
            for (uint32_t q = 0; q < Q; ++q) {
                if (q == uint32_t(conf[i])) mp[q] = planted_noise + (1.0 - planted_noise) * random_real(engine);
                else mp[q] = random_real(engine) * (1.0 - planted_noise);
                for (uint32_t l = 0; l < d; ++l) {
                    double *slot = msg.data() + size_t(r0 + l) * Q;  // mmap_[i][idxij]: the node's OWN in-slots (:185-191)
                    if (q == uint32_t(conf[i])) slot[q] = planted_noise + (1.0 - planted_noise) * random_real(engine);
                    else slot[q] = random_real(engine) * (1.0 - planted_noise);
                }
            }
        } else {
            for (uint32_t q = 0; q < Q; ++q) mp[q] = (q == uint32_t(conf[i])) ? 1.0 : 0.0;
            for (uint32_t l = 0; l < d; l += 2) {  // :203-205
                double *slot = msg.data() + size_t(g.rev[r0 + l]) * Q;
                for (uint32_t q = 0; q < Q; ++q) slot[q] = (q == uint32_t(conf[i])) ? 1.0 : 0.0;
            }
        }
    }
    e->conf_planted.assign(conf, conf + N);
    e->n_planted = 0;
    for (uint32_t i = 0; i < N; ++i) e->n_planted += conf[i] != -1;
    if (!e->d_clamp) CUDA_TRY(cudaMalloc(&e->d_clamp, std::max<size_t>(N, 1) * sizeof(int)));
    if (N) CUDA_TRY(cudaMemcpy(e->d_clamp, e->conf_planted.data(), size_t(N) * sizeof(int), cudaMemcpyHostToDevice));
    e->rng = engine;
    e->rng_valid = true;
    return sbmbp_set_state(e, msg.data(), marg.data());
}

int sbmbp_init_messages(sbmbp_engine *e, uint32_t flag, const int32_t *conf, uint32_t seed) {
    std::mt19937 engine(seed);
    return init_messages_from(e, flag, conf, engine);
}

// init_messages drawing from the engine's own generator where it stands (sbmbp_seed_schedule, sbmbp_rng_shuffle): the
// binary hands ONE std::mt19937 to blockmodel_t::shuffle, init_messages and inference / learning (main.cpp:236-365)
int sbmbp_init_messages_continue(sbmbp_engine *e, uint32_t flag, const int32_t *conf) {
    TRY(need(e, false, false));
    if (!e->rng_valid) {
        set_error("no generator state: call sbmbp_seed_schedule first");
        return SBMBP_ERR_STATE;
    }
    std::mt19937 engine = e->rng;
    return init_messages_from(e, flag, conf, engine);
}

// --mb_rand (main.cpp:299-301 -> blockmodel.cpp:103-106): std::shuffle over the n memberships with the run's generator.
// The permuted memberships themselves are not read by the BP path; the draws are what the run that follows sees.
int sbmbp_rng_shuffle(sbmbp_engine *e, uint32_t n) {
    TRY(need(e, false, false));
    if (!e->rng_valid) {
        set_error("no generator state: call sbmbp_seed_schedule first");
        return SBMBP_ERR_STATE;
    }
    std::vector<unsigned> memberships(n, 0u);
    std::shuffle(memberships.begin(), memberships.end(), e->rng);
    return SBMBP_OK;
}

// main.cpp:318-323: -m infer runs bp_conditional (planted nodes of degree < 50 keep their messages and marginal,
// belief_propagation.cpp:1100-1126), -m learn runs bp_basic (they are updated like any other node).  Default: on.
int sbmbp_set_conditional(sbmbp_engine *e, int on) {
    TRY(need(e, false, false));
    e->conditional = on != 0;
    return SBMBP_OK;
}

int sbmbp_get_marginals(sbmbp_engine *e, double *marg) {
    TRY(need(e, false, true));
    if (!marg) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    TRY(sync_marg(e));
    if (e->N)
        CUDA_TRY(cudaMemcpyAsync(marg, e->d_marg, size_t(e->N) * e->Q * sizeof(double), cudaMemcpyDeviceToHost,
                                 e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}

int sbmbp_get_state(sbmbp_engine *e, double *msg, double *marg, double *h) {
    TRY(need(e, false, true));
    if (msg) {
        if (e->prec == SBMBP_F64) TRY(export_msgs<double>(e, msg));
        else TRY(export_msgs<float>(e, msg));
    }
    if (marg) TRY(sbmbp_get_marginals(e, marg));
    if (h) {
        if (!e->have_params) {
            set_error("h needs parameters");
            return SBMBP_ERR_STATE;
        }
        TRY(ensure_field(e));
        Field f;
        CUDA_TRY(cudaMemcpyAsync(&f, e->d_field[e->sweeps_done & 1u], sizeof(Field), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        std::copy(f.h, f.h + e->Q, h);
    }
    return SBMBP_OK;
}

// Greedy colouring in node order: the smallest colour no neighbour holds (self-loops ignored).  Sparse random graphs need
// a handful; a hub only needs a colour its neighbours do not use.
static int greedy_coloring(const sbmbp_graph &g, std::vector<unsigned char> &color, unsigned &ncolors) {
    color.assign(g.N, 0);
    ncolors = g.N ? 1u : 0u;
    std::vector<uint32_t> mark(256, 0xffffffffu);
    for (uint32_t i = 0; i < g.N; ++i) {
        for (uint64_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) {
            const uint32_t j = g.col[s];
            if (j < i) mark[color[j]] = i;  // only already coloured neighbours matter
        }
        unsigned c = 0;
        while (c < 256 && mark[c] == i) ++c;
        if (c >= 255) {
            set_error("coloured schedule: more than 255 colours needed");
            return SBMBP_ERR_UNSUPPORTED;
        }
        color[i] = (unsigned char)c;
        ncolors = std::max(ncolors, c + 1);
    }
    return SBMBP_OK;
}

int sbmbp_graph_coloring(const sbmbp_graph *g, uint8_t *color, uint32_t *n_colors) {
    if (!g || !color || !n_colors) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    std::vector<unsigned char> c;
    unsigned nc = 0;
    TRY(greedy_coloring(*g, c, nc));
    std::copy(c.begin(), c.end(), color);
    *n_colors = nc;
    return SBMBP_OK;
}

int sbmbp_set_schedule(sbmbp_engine *e, int schedule) {
    TRY(need(e, false, false));
    if (schedule != SBMBP_SCHED_SYNC && schedule != SBMBP_SCHED_COLORED && schedule != SBMBP_SCHED_REPLAY) {
        set_error("unknown schedule");
        return SBMBP_ERR_ARG;
    }
    if (schedule == SBMBP_SCHED_REPLAY && e->dist) {
        set_error("the replay schedule is single-GPU only");
        return SBMBP_ERR_UNSUPPORTED;
    }
    if (schedule == SBMBP_SCHED_COLORED) {
        if (e->dist) {
            set_error("the coloured schedule is single-GPU only");
            return SBMBP_ERR_UNSUPPORTED;
        }
        if (!e->d_color) {
            std::vector<unsigned char> color;
            TRY(greedy_coloring(*e->g, color, e->ncolors));
            CUDA_TRY(cudaMalloc(&e->d_color, std::max<size_t>(e->N, 1)));
            if (e->N) CUDA_TRY(cudaMemcpy(e->d_color, color.data(), e->N, cudaMemcpyHostToDevice));
        }
    }
    e->schedule = schedule;
    return SBMBP_OK;
}

// one coloured sweep: a pass per colour through the general kernel (nodes of the other colours are carried forward), the
// field h refreshed after every pass; returns the largest max-diff of the passes
static int colored_sweep(sbmbp_engine *e, double damping, double *maxdiff) {
    double md = 0.0;
    for (unsigned c = 0; c < e->ncolors; ++c) {
        e->cur_color = c;
        TRY(arm_ctl(e, -1.0f, 1));
        TRY(run_sweeps(e, 1, damping));
        TRY(download_ctl(e));
        md = std::max(md, e->h_ctl->last_maxdiff);
    }
    e->state_version++;
    e->stat_sweeps += 1;
    e->stat_edge_updates += e->M;
    *maxdiff = md;
    return SBMBP_OK;
}

int sbmbp_seed_schedule(sbmbp_engine *e, uint32_t seed) {
    TRY(need(e, false, false));
    e->rng.seed(seed);
    e->rng_valid = true;
    return SBMBP_OK;
}

// Reference-exact replay (sweep_replay.cuh): up to max_sweeps sweeps of N with-replacement draws from e->rng, the state
// held in the reference's order in d_scratch meanwhile.  test != 0: stop after the first sweep whose maxdiffm < crit
// (:406) and report its index; the generator is left where the reference's would be.
static int replay_sweeps(sbmbp_engine *e, float crit, uint32_t max_sweeps, double damping, int test, int *niter,
                         double *last_maxdiff) {
    if (!e->rng_valid) {
        set_error("replay schedule: no generator state (call sbmbp_init_messages / sbmbp_init_random or sbmbp_seed_schedule)");
        return SBMBP_ERR_STATE;
    }
    const auto &g = *e->g;
    const uint32_t N = e->N, Q = e->Q;
    const size_t md = std::max<uint32_t>(g.max_degree, 1);
    TRY(sync_marg(e));
    TRY(ensure_full(e));
    if (!e->d_rp_rev) {
        CUDA_TRY(cudaMalloc(&e->d_rp_rev, std::max<size_t>(e->M, 1) * sizeof(unsigned)));
        if (e->M) CUDA_TRY(cudaMemcpy(e->d_rp_rev, g.rev.data(), size_t(e->M) * sizeof(unsigned), cudaMemcpyHostToDevice));
        if (e->dc != 0) {
            std::vector<unsigned> degn(e->M);
            for (uint64_t s = 0; s < e->M; ++s) degn[s] = g.deg[g.col[s]];
            CUDA_TRY(cudaMalloc(&e->d_rp_degn, std::max<size_t>(e->M, 1) * sizeof(unsigned)));
            if (e->M) CUDA_TRY(cudaMemcpy(e->d_rp_degn, degn.data(), size_t(e->M) * sizeof(unsigned), cudaMemcpyHostToDevice));
        }
        CUDA_TRY(cudaMalloc(&e->d_rp_sched, std::max<size_t>(N, 1) * sizeof(unsigned)));
        CUDA_TRY(cudaMalloc(&e->d_rp_h, (kMaxQ + 1) * sizeof(double)));
        CUDA_TRY(cudaMalloc(&e->d_rp_scratch, (4 + size_t(Q)) * md * sizeof(double)));
        CUDA_TRY(cudaMemset(e->d_rp_scratch, 0, (4 + size_t(Q)) * md * sizeof(double)));
    }
    const size_t n = size_t(e->M) * Q;
    TRY(ensure_scratch(e, std::max<size_t>(n, 1)));
    const unsigned blocks = unsigned(std::max<size_t>(1, std::min<size_t>((n + 255) / 256, size_t(e->sm_count) * 16)));
    void *S = e->d_S[e->sweeps_done & 1u];
    if (n) {
        if (e->prec == SBMBP_F64)
            export_msgs_kernel<double><<<blocks, 256, 0, e->stream>>>(static_cast<const double *>(S), e->d_rev, e->d_scratch, e->M, Q);
        else
            export_msgs_kernel<float><<<blocks, 256, 0, e->stream>>>(static_cast<const float *>(S), e->d_rev, e->d_scratch, e->M, Q);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
    }
    ReplayArgs a;
    a.row_ptr = e->d_row_ptr;
    a.rev = e->d_rp_rev;
    a.degn = e->d_rp_degn;
    a.clamp = (e->conditional && e->n_planted) ? e->d_clamp : nullptr;
    a.sched = e->d_rp_sched;
    a.count = N;
    a.msg = e->d_scratch;
    a.marg = e->d_marg;
    a.h = e->d_rp_h;
    a.prm = e->d_prm;
    a.scratch = e->d_rp_scratch;
    a.max_degree = unsigned(md);
    a.N = N;
    a.Q = Q;
    a.dc = e->dc;
    a.damping = damping;
    a.out = e->d_rp_h + kMaxQ;
    std::uniform_real_distribution<> random_real(0, 1);
    std::vector<unsigned> draws(N);
    int result = -1;
    double mdiff = -100.0;
    unsigned done = 0;
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    for (uint32_t s = 0; s < max_sweeps; ++s) {
        for (uint32_t k = 0; k < N; ++k) draws[k] = unsigned(int(random_real(e->rng) * N));  // :395
        if (N) CUDA_TRY(cudaMemcpyAsync(e->d_rp_sched, draws.data(), size_t(N) * sizeof(unsigned), cudaMemcpyHostToDevice, e->stream));
        a.init_field = (s == 0);  // converge() starts with init_h (:390)
        bp_replay_kernel<<<1, 32, 0, e->stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
        CUDA_TRY(cudaMemcpyAsync(e->h_out, a.out, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));  // also: draws may be overwritten now
        mdiff = e->h_out[0];
        ++done;
        if (test && mdiff < crit) {  // double < float, :406
            result = int(s);
            break;
        }
    }
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    if (n) {
        if (e->prec == SBMBP_F64)
            import_msgs_kernel<double><<<blocks, 256, 0, e->stream>>>(e->d_scratch, e->d_rev, static_cast<double *>(S), e->M, Q);
        else
            import_msgs_kernel<float><<<blocks, 256, 0, e->stream>>>(e->d_scratch, e->d_rev, static_cast<float *>(S), e->M, Q);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->stat_seconds += ms * 1e-3;
    e->stat_sweeps += done;
    e->stat_edge_updates += uint64_t(done) * e->M;  // in expectation: N draws with replacement touch M edges
    e->field_valid = false;  // the engine's own field is rebuilt from the marginals on demand
    e->state_version++;
    if (niter) *niter = result;
    if (last_maxdiff) *last_maxdiff = mdiff;
    return SBMBP_OK;
}

int sbmbp_sweep(sbmbp_engine *e, double damping, double *maxdiff) {
    TRY(need(e, true, true));
    if (e->dist) {
        set_error("single-GPU entry point called on a multi-GPU engine (use the sbmbp_dist_* calls)");
        return SBMBP_ERR_STATE;
    }
    if (e->schedule == SBMBP_SCHED_REPLAY) return replay_sweeps(e, 0.f, 1, damping, 0, nullptr, maxdiff);
    TRY(ensure_field(e));
    if (e->schedule == SBMBP_SCHED_COLORED && e->ntiles) {
        double md = 0.0;
        TRY(colored_sweep(e, damping, &md));
        if (maxdiff) *maxdiff = md;
        return SBMBP_OK;
    }
    TRY(arm_ctl(e, -1.0f, 1));
    TRY(run_sweeps(e, 1, damping));
    TRY(download_ctl(e));
    if (e->ntiles == 0) e->h_ctl->last_maxdiff = 0.0;
    e->state_version++;
    e->stat_sweeps += 1;
    e->stat_edge_updates += e->M;
    if (maxdiff) *maxdiff = e->h_ctl->last_maxdiff;
    return SBMBP_OK;
}

int sbmbp_sweeps_async(sbmbp_engine *e, uint32_t n, double damping) {
    TRY(need(e, true, true));
    if (e->dist) {
        set_error("single-GPU entry point called on a multi-GPU engine (use the sbmbp_dist_* calls)");
        return SBMBP_ERR_STATE;
    }
    if (e->schedule == SBMBP_SCHED_REPLAY) return replay_sweeps(e, 0.f, n, damping, 0, nullptr, nullptr);
    TRY(ensure_field(e));
    if (e->schedule == SBMBP_SCHED_COLORED && e->ntiles) {
        for (uint32_t s = 0; s < n; ++s) {
            double md = 0.0;
            TRY(colored_sweep(e, damping, &md));
        }
        return SBMBP_OK;
    }
    TRY(arm_ctl(e, -1.0f, n));
    TRY(run_sweeps(e, n, damping));
    if (e->ntiles) e->sweeps_done += n;  // no convergence test: the count is known without reading it back
    e->state_version++;
    e->stat_sweeps += n;
    e->stat_edge_updates += uint64_t(n) * e->M;
    return SBMBP_OK;
}

// one sweep, returning the device time of the sweep kernel alone (CUDA events on the engine's stream around that
// one launch; the arm and finalize launches are outside the bracket) -- the roofline measurement of bench.py
int sbmbp_time_sweep_kernel(sbmbp_engine *e, double damping, float *kernel_ms) {
    TRY(need(e, true, true));
    if (e->dist) {
        set_error("single-GPU entry point");
        return SBMBP_ERR_STATE;
    }
    TRY(ensure_field(e));
    TRY(arm_ctl(e, -1.0f, 1));
    e->time_kernel = true;
    int rc = run_sweeps(e, 1, damping);
    e->time_kernel = false;
    TRY(rc);
    CUDA_TRY(cudaEventSynchronize(e->ev1));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    if (e->ntiles) CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    if (e->ntiles) e->sweeps_done += 1;
    e->state_version++;
    e->stat_sweeps += 1;
    e->stat_edge_updates += e->M;
    if (kernel_ms) *kernel_ms = ms;
    return SBMBP_OK;
}

int sbmbp_sync(sbmbp_engine *e) {
    TRY(need(e, false, false));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}

// converge() (belief_propagation.cpp:386-415).  Sweeps are launched in batches; every kernel checks the
// device-side convergence flag first, so the host synchronises once per batch rather than once per sweep.
int sbmbp_converge(sbmbp_engine *e, float crit, uint32_t max_sweeps, float damping, int *niter) {
    TRY(need(e, true, true));
    if (e->dist) {
        set_error("single-GPU entry point called on a multi-GPU engine (use the sbmbp_dist_* calls)");
        return SBMBP_ERR_STATE;
    }
    if (e->schedule == SBMBP_SCHED_REPLAY) return replay_sweeps(e, crit, max_sweeps, double(damping), 1, niter, nullptr);
    e->field_valid = false;  // converge() always starts with init_h (:390)
    TRY(ensure_field(e));
    int result = -1;
    if (e->ntiles == 0) {
        // no nodes: the reference's loop body never runs and maxdiffm = -100 < crit at the first check
        if (niter) *niter = max_sweeps ? 0 : -1;
        return SBMBP_OK;
    }
    if (e->schedule == SBMBP_SCHED_COLORED) {
        // the host takes the decision of :406 once per sweep, over the passes of all colours
        CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
        for (uint32_t s = 0; s < max_sweeps; ++s) {
            double md = 0.0;
            TRY(colored_sweep(e, double(damping), &md));
            if (md < crit) {
                result = int(s);
                break;
            }
        }
        CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
        CUDA_TRY(cudaEventSynchronize(e->ev1));
        float cms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&cms, e->ev0, e->ev1));
        e->stat_seconds += cms * 1e-3;
        if (niter) *niter = result;
        return SBMBP_OK;
    }
    const unsigned start = e->sweeps_done;
    TRY(arm_ctl(e, crit, max_sweeps));
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    unsigned launched = 0;
    unsigned batch = 4;
    while (launched < max_sweeps) {
        const unsigned cur = std::min(batch, max_sweeps - launched);
        TRY(run_sweeps(e, cur, double(damping)));
        launched += cur;
        TRY(download_ctl(e));
        if (e->h_ctl->converged) {
            result = e->h_ctl->niter;
            break;
        }
        batch = std::min(batch * 2, 32u);
    }
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->stat_seconds += ms * 1e-3;
    const unsigned done = e->sweeps_done - start;
    e->stat_sweeps += done;
    e->stat_edge_updates += uint64_t(done) * e->M;
    e->state_version++;
    if (niter) *niter = result;
    return SBMBP_OK;
}

int sbmbp_free_energy(sbmbp_engine *e, double *f, double *f_site, double *f_edge, double *f_ne) {
    TRY(need(e, true, true));
    if (e->dist) {
        set_error("single-GPU entry point called on a multi-GPU engine (use the sbmbp_dist_* calls)");
        return SBMBP_ERR_STATE;
    }
    const std::vector<double> *r = nullptr;
    TRY(energy_pass(e, 0, &r));
    const double N = double(e->N);
    const double fs = (*r)[0] / N;          // :502
    const double fe = (*r)[1] / (2. * N);   // :610
    double fn = 0.0;
    TRY(f_non_edge(e, &fn));
    if (f_site) *f_site = fs;
    if (f_edge) *f_edge = fe;
    if (f_ne) *f_ne = fn;
    if (f) *f = -fs + fe + fn;  // :744-750
    return SBMBP_OK;
}

int sbmbp_set_exact_pairs_max_n(sbmbp_engine *e, uint32_t n) {
    if (!e) {
        set_error("null engine");
        return SBMBP_ERR_ARG;
    }
    e->exact_pairs_max_n = n;
    e->state_version++;
    return SBMBP_OK;
}

int sbmbp_non_edge_series_order(uint32_t Q, double N, double beta, const double *cab, uint32_t *K) {
    if (!cab || !K || Q < 1 || Q > SBMBP_MAX_Q || !(N >= 1.0)) {
        set_error("bad argument");
        return SBMBP_ERR_ARG;
    }
    std::vector<double> W1;
    const double ymax = series_weights(Q, N, beta, cab, W1);
    *K = series_order(Q, N, ymax, 1e-14);
    return SBMBP_OK;
}

int sbmbp_non_edge_series_term(uint32_t Q, double N, double beta, const double *cab, uint32_t k, const double *T,
                               double *term) {
    if (!cab || !T || !term || Q < 1 || Q > SBMBP_MAX_Q || k > 8 || !(N >= 1.0)) {
        set_error("bad argument");
        return SBMBP_ERR_ARG;
    }
    if (k == 0) {  // T[0] = sum_i log(sum_q psi_i^q), all ranks: the normalisation defect of the marginals
        *term = 2.0 * N * T[0];
        return SBMBP_OK;
    }
    std::vector<double> W1;
    series_weights(Q, N, beta, cab, W1);
    size_t len = 1;
    for (uint32_t o = 0; o < k; ++o) len *= Q;
    const std::vector<double> Tv(T, T + len);
    std::vector<const double *> mats(k, W1.data());
    *term = -contract_moments(Tv, k, Q, mats) / double(k);
    return SBMBP_OK;
}

int sbmbp_entropy(sbmbp_engine *e, double *entropy) {
    TRY(need(e, true, true));
    if (!entropy) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    if (e->dc != 0) {  // the reference's site term is 0/0 for dc != 0 (:518-523, :550-556)
        *entropy = std::nan("");
        return SBMBP_OK;
    }
    const std::vector<double> *r = nullptr;
    TRY(energy_pass(e, 1, &r));
    const double N = double(e->N);
    double ene = 0.0;
    TRY(entropy_non_edge(e, &ene));
    const double e_site = -((*r)[2] / N);
    const double e_link = (*r)[3] / (2. * N);
    *entropy = e_site + e_link - ene;  // :752-758
    return SBMBP_OK;
}

int sbmbp_overlap(sbmbp_engine *e, const uint32_t *true_conf, double *overlap) {
    TRY(need(e, false, true));
    if (!true_conf || !overlap) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    std::vector<double> row;
    TRY(node_stats(e, true_conf, row));
    const uint32_t Q = e->Q;
    std::vector<unsigned> perm(Q);
    for (uint32_t q = 0; q < Q; ++q) perm[q] = q;
    double max_ov = -1.0;
    do {  // :784-790: all relabellings for Q <= 8, the identity only above
        double ov = 0.0;
        for (uint32_t t = 0; t < Q; ++t) ov += row[2 * kMaxQ + t * kMaxQ + perm[t]];
        ov /= double(e->N);
        if (ov > max_ov) max_ov = ov;
    } while (Q <= 8 && std::next_permutation(perm.begin(), perm.end()));
    *overlap = max_ov;
    return SBMBP_OK;
}

int sbmbp_em_stats(sbmbp_engine *e, double *na_expect, double *nna_expect, double *cab_expect) {
    TRY(need(e, true, true));
    const uint32_t Q = e->Q;
    std::vector<double> row;
    TRY(node_stats(e, nullptr, row));
    const double *na = row.data(), *nna = row.data() + kMaxQ;
    if (na_expect) std::copy(na, na + Q, na_expect);
    if (nna_expect) std::copy(nna, nna + Q, nna_expect);
    if (cab_expect) {
        const std::vector<double> *r = nullptr;
        TRY(energy_pass(e, 1, &r));
        const int qt = e->qt;
        const double N = double(e->N);
        for (uint32_t q1 = 0; q1 < Q; ++q1)
            for (uint32_t q2 = q1; q2 < Q; ++q2) {
                double v = (*r)[kEnergyHead + q1 * qt + q2];
                if (na[q1] > kEps && na[q2] > kEps) {  // :970
                    const double *w = (e->dc == 0) ? na : nna;
                    v *= ((q1 == q2) ? 2. * N : N) / (w[q1] * w[q2]);  // :972-985
                }
                cab_expect[q1 * Q + q2] = cab_expect[q2 * Q + q1] = v;
            }
    }
    return SBMBP_OK;
}

// learning() (belief_propagation.cpp:14-51): -t bounds both the EM iterations and the BP sweeps per E-step,
// the BP tolerance is the learning criterion, messages are not re-initialised between EM iterations.
int sbmbp_learn(sbmbp_engine *e, float learning_conv_crit, uint32_t learning_max_time, float learning_rate,
                float dumping_rate, uint32_t *na_out, double *cab_out, double *eta_out, int *em_iters) {
    TRY(need(e, true, true));
    const uint32_t Q = e->Q;
    std::vector<double> na_e(Q), nna_e(Q), cab_e(size_t(Q) * Q);
    double fold = 0.0, fdiff = 1.0;
    int learning_time = 0;
    for (learning_time = 0; learning_time < int(learning_max_time); learning_time++) {
        if (fdiff < learning_conv_crit) learning_conv_crit *= 0.1;
        int niter = 0;
        TRY(sbmbp_converge(e, learning_conv_crit, learning_max_time, dumping_rate, &niter));
        TRY(sbmbp_em_stats(e, na_e.data(), nna_e.data(), cab_e.data()));
        double fnew = 0.0;
        TRY(sbmbp_free_energy(e, &fnew, nullptr, nullptr, nullptr));
        fdiff = std::fabs(fnew - fold);
        fold = fnew;
        if (std::isnan(fold) || std::isinf(fold)) break;
        if (fdiff < learning_conv_crit) break;
        learning_step_host(e, learning_rate, na_e.data(), cab_e.data());
        TRY(apply_params(e));
    }
    if (na_out) std::copy(e->na.begin(), e->na.end(), na_out);
    if (cab_out) std::copy(e->cab.begin(), e->cab.end(), cab_out);
    if (eta_out) std::copy(e->eta.begin(), e->eta.end(), eta_out);
    if (em_iters) *em_iters = learning_time;
    return SBMBP_OK;
}

// which sweep kernel the next sbmbp_sweep / sbmbp_converge launches (same decision as launch_sweeps in inst.cu)
int sbmbp_sweep_kernel_name(sbmbp_engine *e, char *buf, uint32_t cap) {
    if (!e || !buf || cap == 0) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    const char *t = (e->prec == SBMBP_F64) ? "double" : "float";
    const size_t elt = (e->prec == SBMBP_F64) ? 8 : 4;
    const bool can_fast = (e->qt * elt) % 16 == 0 || e->qt * elt == 8;
    const bool select_k = (e->dc == 0 && e->beta != 1.0);
    const bool clamped = (e->conditional && e->n_planted) || e->schedule == SBMBP_SCHED_COLORED;
    const bool fast = can_fast && e->fast_path && e->Q == uint32_t(e->qt) && e->dc != 2 && !select_k && !clamped;
    std::string name;
    if (e->schedule == SBMBP_SCHED_REPLAY) name = "bp_replay_kernel";
    else if (e->dist) name = "bp_sweep_pipe_dist_kernel<" + std::string(t) + "," + std::to_string(e->qt) + ">";
    else if (fast && e->wide_path)
        name = "bp_sweep_wide_kernel<" + std::string(t) + ">" + (e->nbtiles ? " (+ bp_sweep_fast_kernel for degrees > 32)" : "");
    else if (fast && e->qt <= 4 && e->ell_path)
        name = "bp_sweep_ell_kernel<" + std::string(t) + "," + std::to_string(e->qt) + (e->compact_ok && !e->compact_refused ? ",compact>" : ">") +
               (e->ell_padded ? " (padded one-bucket layout)" : "") +
               ((e->nwtiles || e->nhubs) ? " (+ bp_sweep_warp_kernel / bp_sweep_hub_kernel for degrees >= 32)" : "");
    else if (fast && e->qt <= 4 && e->warp_path && e->d_wtiles)
        name = "bp_sweep_warp_kernel<" + std::string(t) + "," + std::to_string(e->qt) + ">";
    else if (fast && e->pipe_path && !(e->qt == 32 && e->prec == SBMBP_F64))
        name = "bp_sweep_pipe_kernel<" + std::string(t) + "," + std::to_string(e->qt) + ",false>";
    else if (fast) name = "bp_sweep_fast_kernel<" + std::string(t) + "," + std::to_string(e->qt) + ",false>";
    else name = "bp_sweep_kernel<" + std::string(t) + "," + std::to_string(e->qt) + "> + bp_finalize_kernel";
    std::snprintf(buf, cap, "%s", name.c_str());
    return SBMBP_OK;
}

int sbmbp_stats(sbmbp_engine *e, uint64_t *edge_updates, uint64_t *sweeps, uint64_t *launches, double *bytes_per_edge,
                double *sweep_seconds) {
    if (!e) {
        set_error("null engine");
        return SBMBP_ERR_ARG;
    }
    if (edge_updates) *edge_updates = e->stat_edge_updates;
    if (sweeps) *sweeps = e->stat_sweeps;
    if (launches) *launches = e->stat_launches;
    if (bytes_per_edge) {
        // B = 3 Q s + 4 [+4 if dc != 0] + (8 + Q s) / (M / N)
        const double s = (e->prec == SBMBP_F64) ? 8.0 : 4.0;
        const double cbar = e->N ? double(e->M) / double(e->N) : 1.0;
        *bytes_per_edge = 3.0 * e->Q * s + 4.0 + (e->dc ? 4.0 : 0.0) + (cbar > 0 ? (8.0 + e->Q * s) / cbar : 0.0);
    }
    if (sweep_seconds) *sweep_seconds = e->stat_seconds;
    return SBMBP_OK;
}

int sbmbp_tiny_events(sbmbp_engine *e, uint64_t *n) {
    TRY(need(e, false, false));
    if (!n) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    TRY(download_ctl(e));
    *n = e->h_ctl->tiny_count;
    return SBMBP_OK;
}

}  // extern "C"

// =============================================================================================== multi-GPU
//
// One process per GPU.  Rank p owns a contiguous range of nodes, the rows (in-slots) of those nodes, their
// marginals and the message buffers holding every message INTO them; a message i -> j is produced on owner(i)
// and stored on owner(j), written there directly by the sweep kernel through a CUDA-IPC mapping (NVLink).
// The layout of a rank's buffer is the same destination-bucketed order as on one GPU: region b = in-slots of
// bucket b, filled in global source order (source node, then destination).  Because a rank holds all in-edges of
// its nodes, it can compute that order alone; what the PRODUCERS need -- where each of their out-messages goes --
// is sent to them once (plan_sendlist / plan_recv), e.g. with torch.distributed.all_to_all.

struct sbmbp_plan {
    const sbmbp_graph *g = nullptr;
    int rank = 0, world = 1;
    uint32_t Q = 0;
    int prec = SBMBP_F64, qt = 2, te = 0, tn = 0;
    std::vector<uint32_t> starts;  // world + 1 node range boundaries
    std::vector<Tile> tiles;
    std::vector<unsigned> gather;  // per in-slot: position of its message in this rank's buffer
    std::vector<std::vector<unsigned>> sendlist, recvlist;
    std::vector<uint64_t> expect;  // values expected from each peer
    std::vector<unsigned> pos, info;  // pos: the kernels' word per tile entry (bit 31: outbox index, else own-buffer position)
    std::vector<unsigned> rpos;       // per tile entry: owner << 29 | position at the owner (mirror pull, inspection)
    std::vector<unsigned> pos_slot;   // owner << 29 | position, in slot order (before the per-tile sort), kept for inspection
    // halo exchange (dist_exchange.cuh): super-tiles of tps tiles; outbox range and shipping descriptors of each
    unsigned tps = 8, nsuper = 0;
    std::vector<unsigned> out_start, ship_start;
    std::vector<ShipDesc> ship;
    std::vector<unsigned> out_rpos;  // per outbox entry: owner << 29 | position at the owner
    uint64_t n_remote = 0;
    unsigned nbuckets = 1;
    bool finished = false;
};

namespace {

int owner_of(const std::vector<uint32_t> &starts, uint32_t node) {
    return int(std::upper_bound(starts.begin(), starts.end(), node) - starts.begin()) - 1;
}

double region_mb_setting() {
    double region_mb = 16.0;
    if (const char *env = std::getenv("SBMBP_REGION_MB")) region_mb = std::atof(env);
    return region_mb;
}

}  // namespace

extern "C" {

int sbmbp_graph_from_pairs_range(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N_global,
                                 uint32_t lo, uint32_t hi, sbmbp_graph **g) {
    if (!g || (n_pairs && (!u || !v))) {
        set_error("null argument");
        return SBMBP_ERR_ARG;
    }
    auto *gr = new sbmbp_graph();
    int rc = build_graph_range(u, v, n_pairs, N_global, lo, hi, *gr);
    if (rc != SBMBP_OK) {
        delete gr;
        return rc;
    }
    *g = gr;
    return SBMBP_OK;
}

int sbmbp_plan_create(const sbmbp_graph *g, uint32_t Q, int precision, int rank, int world,
                      const uint32_t *range_starts, sbmbp_plan **out) {
    if (!g || !out || !range_starts || world < 1 || world > 8 || rank < 0 || rank >= world) {
        set_error("bad argument (world must be 1..8)");
        return SBMBP_ERR_ARG;
    }
    if (g->N_global == 0) {
        set_error("sbmbp_plan_create needs a rank-local graph (sbmbp_graph_from_pairs_range)");
        return SBMBP_ERR_ARG;
    }
    if (Q < 1 || Q > SBMBP_MAX_Q) {
        set_error("Q must be in [1, " + std::to_string(SBMBP_MAX_Q) + "]");
        return SBMBP_ERR_UNSUPPORTED;
    }
    auto *p = new sbmbp_plan();
    p->g = g;
    p->rank = rank;
    p->world = world;
    p->Q = Q;
    p->prec = precision;
    p->qt = pick_qt(Q);
    p->starts.assign(range_starts, range_starts + world + 1);
    if (p->starts[rank] != g->node_lo || p->starts[rank + 1] != g->node_lo + g->N) {
        delete p;
        set_error("range_starts[rank] does not match the graph's node range");
        return SBMBP_ERR_ARG;
    }
    sbmbp_engine probe;
    probe.prec = precision;
    probe.qt = p->qt;
    dispatch(&probe, [&](auto t, auto qt) {
        tile_geometry<decltype(t), decltype(qt)::value>(p->te, p->tn);
        return SBMBP_OK;
    });
    p->tiles = make_tiles(*g, p->te, p->tn);
    const uint64_t M = g->M;
    const size_t elt = (precision == SBMBP_F64) ? 8 : 4;
    // multi-GPU: every (tile, bucket) group of out-messages is one NVLink write burst, so buckets are made coarser
    // with the number of ranks (measured at 8 GPUs: 16 MiB regions 7.0 ms/sweep, 48 MiB 5.2 ms, 128 MiB 4.0 ms)
    double region_mb = std::min(128.0, 16.0 * world);
    if (std::getenv("SBMBP_REGION_MB")) region_mb = region_mb_setting();
    const uint64_t region_slots = uint64_t(region_mb * 1048576.0 / double(Q * elt));
    // buckets over the local nodes
    std::vector<uint64_t> bucket_start;  // first in-slot of each bucket
    {
        uint64_t next = 0;
        for (uint32_t i = 0; i < g->N; ++i)
            if (g->row_ptr[i] >= next && (region_slots != 0 || bucket_start.empty())) {
                bucket_start.push_back(g->row_ptr[i]);
                next = g->row_ptr[i] + (region_slots ? region_slots : M + 1);
            }
        if (bucket_start.empty()) bucket_start.push_back(0);
    }
    p->nbuckets = unsigned(bucket_start.size());
    // in-slots in global source order: key = (source node j, slot e); for a fixed j, slot order is destination order
    std::vector<uint64_t> keys(M);
    for (uint64_t e = 0; e < M; ++e) keys[e] = (uint64_t(g->col[e]) << 32) | e;
    {
        // parallel sort: split by source-node chunks, sort each chunk in a thread
        const unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        const unsigned nchunks = (M > (1u << 20)) ? nthreads * 4 : 1;
        if (nchunks == 1) {
            std::sort(keys.begin(), keys.end());
        } else {
            const uint64_t span = (uint64_t(g->N_global) + nchunks - 1) / nchunks;
            std::vector<uint64_t> cnt(nchunks + 1, 0);
            for (uint64_t e = 0; e < M; ++e) cnt[(keys[e] >> 32) / span + 1]++;
            for (unsigned c = 0; c < nchunks; ++c) cnt[c + 1] += cnt[c];
            std::vector<uint64_t> sorted(M), cur(cnt.begin(), cnt.end() - 1);
            for (uint64_t e = 0; e < M; ++e) sorted[cur[(keys[e] >> 32) / span]++] = keys[e];
            keys.swap(sorted);
            std::vector<std::thread> pool;
            std::atomic<unsigned> nextc{0};
            for (unsigned t = 0; t < nthreads; ++t)
                pool.emplace_back([&]() {
                    for (unsigned c = nextc++; c < nchunks; c = nextc++)
                        std::sort(keys.begin() + cnt[c], keys.begin() + cnt[c + 1]);
                });
            for (auto &th : pool) th.join();
        }
    }
    p->gather.assign(M, 0);
    p->sendlist.assign(world, {});
    p->recvlist.assign(world, {});
    p->expect.assign(world, 0);
    {
        std::vector<uint64_t> cursor(bucket_start);
        for (uint64_t k = 0; k < M; ++k) {
            const uint32_t j = uint32_t(keys[k] >> 32);
            const uint64_t e = keys[k] & 0xffffffffull;
            const size_t b = size_t(std::upper_bound(bucket_start.begin(), bucket_start.end(), e) - bucket_start.begin()) - 1;
            const unsigned where = unsigned(cursor[b]++);
            p->gather[e] = where;
            p->sendlist[owner_of(p->starts, j)].push_back(where);
        }
    }
    for (uint64_t e = 0; e < M; ++e) p->expect[owner_of(p->starts, g->col[e])]++;
    *out = p;
    return SBMBP_OK;
}

int sbmbp_plan_destroy(sbmbp_plan *p) {
    delete p;
    return SBMBP_OK;
}

int sbmbp_plan_sendlist(sbmbp_plan *p, int peer, const uint32_t **data, uint64_t *n) {
    if (!p || peer < 0 || peer >= p->world || !data || !n) {
        set_error("bad argument");
        return SBMBP_ERR_ARG;
    }
    *data = p->sendlist[peer].data();
    *n = p->sendlist[peer].size();
    return SBMBP_OK;
}

int sbmbp_plan_expect(sbmbp_plan *p, int peer, uint64_t *n) {
    if (!p || peer < 0 || peer >= p->world || !n) {
        set_error("bad argument");
        return SBMBP_ERR_ARG;
    }
    *n = p->expect[peer];
    return SBMBP_OK;
}

int sbmbp_plan_recv(sbmbp_plan *p, int peer, const uint32_t *data, uint64_t n) {
    if (!p || peer < 0 || peer >= p->world || (n && !data)) {
        set_error("bad argument");
        return SBMBP_ERR_ARG;
    }
    if (n != p->expect[peer]) {
        set_error("peer " + std::to_string(peer) + " sent " + std::to_string(n) + " positions, expected " +
                  std::to_string(p->expect[peer]));
        return SBMBP_ERR_ARG;
    }
    p->recvlist[peer].assign(data, data + n);
    return SBMBP_OK;
}

int sbmbp_plan_finish(sbmbp_plan *p) {
    if (!p) {
        set_error("null plan");
        return SBMBP_ERR_ARG;
    }
    const sbmbp_graph &g = *p->g;
    for (int k = 0; k < p->world; ++k)
        if (p->recvlist[k].size() != p->expect[k]) {
            set_error("positions from peer " + std::to_string(k) + " are missing");
            return SBMBP_ERR_STATE;
        }
    // out-slot (j, i) of this rank <- position chosen by owner(i); both sides enumerate the edges between two
    // ranks in (source node, destination node) order
    p->pos.assign(g.M, 0);
    std::vector<uint64_t> cur(p->world, 0);
    for (uint64_t s = 0; s < g.M; ++s) {
        const int o = owner_of(p->starts, g.col[s]);
        const unsigned where = p->recvlist[o][cur[o]++];
        p->pos[s] = (unsigned(o) << 29) | where;
    }
    p->pos_slot = p->pos;
    // ---- tile order.  Within a tile the entries are sorted by (local first, then owner), then position: local writes
    // advance through this rank's regions, remote ones through the outbox.  Hub tiles keep slot order.
    const unsigned my = unsigned(p->rank);
    auto key_of = [&](unsigned ow) -> uint64_t {
        const unsigned o = ow >> 29, where = ow & ((1u << 29) - 1u);
        return (uint64_t(o == my ? 0u : o + 1u) << 32) | where;
    };
    p->info.assign(g.M, 0);
    p->rpos.assign(g.M, 0);
    const size_t ntiles = p->tiles.size();
    const unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    auto parallel_for = [&](size_t n, const std::function<void(size_t, size_t)> &fn) {
        std::vector<std::thread> pool;
        const size_t per = (n + nthreads - 1) / nthreads;
        for (unsigned i = 0; i < nthreads; ++i) {
            const size_t lo = std::min(n, size_t(i) * per), hi = std::min(n, lo + per);
            if (lo < hi) pool.emplace_back(fn, lo, hi);
        }
        for (auto &th : pool) th.join();
    };
    parallel_for(ntiles, [&](size_t lo, size_t hi) {
        std::vector<std::pair<uint64_t, std::pair<unsigned, unsigned>>> tmp;  // key, (owner|where, info)
        for (size_t b = lo; b < hi; ++b) {
            const Tile &t = p->tiles[b];
            if (t.ne > unsigned(p->te)) {  // hub: slot order
                for (unsigned k = 0; k < t.ne; ++k) p->rpos[t.e0 + k] = p->pos[t.e0 + k];
                continue;
            }
            tmp.resize(t.ne);
            for (unsigned n = 0; n < t.nn; ++n) {
                const uint32_t node = t.n0 + n;
                const unsigned flag = (g.deg[node] >= kLargeDegree) ? 0x80000000u : 0u;
                for (uint64_t sl = g.row_ptr[node]; sl < g.row_ptr[node + 1]; ++sl) {
                    const unsigned k = unsigned(sl - t.e0);
                    tmp[k] = {key_of(p->pos[sl]), {p->pos[sl], flag | (n << 16) | k}};
                }
            }
            std::sort(tmp.begin(), tmp.end());
            for (unsigned k = 0; k < t.ne; ++k) {
                p->rpos[t.e0 + k] = tmp[k].second.first;
                p->info[t.e0 + k] = tmp[k].second.second;
            }
        }
    });
    // ---- outbox order and shipping descriptors, per super-tile (a run of tps consecutive tiles): the remote entries of a
    // super-tile sorted by (owner, position) -- the same key as inside a tile, so a tile's run stays a run -- get
    // consecutive outbox indices; maximal runs that are also consecutive at the owner become one descriptor each
    if (const char *env = std::getenv("SBMBP_SUPERTILE")) {  // a power of two: the kernels shift and mask
        const unsigned want = unsigned(std::max(1, std::atoi(env)));
        p->tps = 1;
        while (p->tps * 2 <= want && p->tps < 1024) p->tps *= 2;
    }
    p->nsuper = unsigned((ntiles + p->tps - 1) / p->tps);
    p->out_start.assign(p->nsuper + 1, 0);
    p->ship_start.assign(p->nsuper + 1, 0);
    std::vector<uint64_t> remote_in(p->nsuper + 1, 0);
    for (unsigned sp = 0; sp < p->nsuper; ++sp) {
        uint64_t cnt = 0;
        const size_t t1 = std::min(ntiles, size_t(sp + 1) * p->tps);
        const uint64_t e_lo = p->tiles[size_t(sp) * p->tps].e0;
        const uint64_t e_hi = (t1 < ntiles) ? p->tiles[t1].e0 : g.M;
        for (uint64_t e = e_lo; e < e_hi; ++e) cnt += (p->rpos[e] >> 29) != my;
        remote_in[sp + 1] = remote_in[sp] + cnt;
    }
    p->n_remote = remote_in[p->nsuper];
    if (p->n_remote >= (1ull << 31)) {
        set_error("more than 2^31 remote out-messages on one rank");
        return SBMBP_ERR_UNSUPPORTED;
    }
    for (unsigned sp = 0; sp <= p->nsuper; ++sp) p->out_start[sp] = unsigned(remote_in[sp]);
    p->out_rpos.assign(size_t(p->n_remote), 0);
    std::vector<std::vector<ShipDesc>> per_super(p->nsuper);
    parallel_for(p->nsuper, [&](size_t lo, size_t hi) {
        std::vector<std::pair<uint64_t, uint64_t>> rem;  // key, entry
        for (size_t sp = lo; sp < hi; ++sp) {
            const size_t t1 = std::min(ntiles, (sp + 1) * size_t(p->tps));
            const uint64_t e_lo = p->tiles[sp * size_t(p->tps)].e0;
            const uint64_t e_hi = (t1 < ntiles) ? p->tiles[t1].e0 : g.M;
            rem.clear();
            for (uint64_t e = e_lo; e < e_hi; ++e) {
                const unsigned ow = p->rpos[e];
                if ((ow >> 29) == my) p->pos[e] = ow & ((1u << 29) - 1u);  // own buffer: the position itself
                else rem.push_back({key_of(ow), e});
            }
            std::sort(rem.begin(), rem.end());
            std::vector<ShipDesc> &out = per_super[sp];
            for (size_t k = 0; k < rem.size(); ++k) {
                const unsigned idx = p->out_start[sp] + unsigned(k);
                const unsigned ow = p->rpos[rem[k].second];
                const unsigned o = ow >> 29, where = ow & ((1u << 29) - 1u);
                p->pos[rem[k].second] = kRemoteBit | idx;
                p->out_rpos[idx] = ow;
                if (!out.empty() && out.back().rank == o && out.back().dst + out.back().len == where) {
                    out.back().len++;
                } else {
                    ShipDesc d;
                    d.src = idx;
                    d.dst = where;
                    d.len = 1;
                    d.rank = o;
                    out.push_back(d);
                }
            }
        }
    });
    p->ship.clear();
    for (unsigned sp = 0; sp < p->nsuper; ++sp) {
        p->ship_start[sp] = unsigned(p->ship.size());
        p->ship.insert(p->ship.end(), per_super[sp].begin(), per_super[sp].end());
    }
    p->ship_start[p->nsuper] = unsigned(p->ship.size());
    for (auto &v : p->recvlist) std::vector<unsigned>().swap(v);
    p->finished = true;
    return SBMBP_OK;
}

int sbmbp_plan_layout(sbmbp_plan *p, const uint32_t **gather, const uint32_t **pos, const uint32_t **info,
                      const uint32_t **pos_slot, uint64_t *M, uint32_t *ntiles) {
    if (!p || !p->finished) {
        set_error("plan not finished");
        return SBMBP_ERR_STATE;
    }
    if (gather) *gather = p->gather.data();
    if (pos) *pos = p->pos.data();
    if (info) *info = p->info.data();
    if (pos_slot) *pos_slot = p->pos_slot.data();
    if (M) *M = p->g->M;
    if (ntiles) *ntiles = unsigned(p->tiles.size());
    return SBMBP_OK;
}

int sbmbp_plan_exchange_tables(sbmbp_plan *p, const uint32_t **rpos, uint32_t *tiles_per_super, uint32_t *nsuper,
                               const uint32_t **out_start, const uint32_t **ship_start, const uint32_t **ship,
                               uint64_t *n_ship, uint64_t *n_remote) {
    if (!p || !p->finished) {
        set_error("plan not finished");
        return SBMBP_ERR_STATE;
    }
    static_assert(sizeof(ShipDesc) == 4 * sizeof(uint32_t), "ShipDesc is four words");
    if (rpos) *rpos = p->rpos.data();
    if (tiles_per_super) *tiles_per_super = p->tps;
    if (nsuper) *nsuper = p->nsuper;
    if (out_start) *out_start = p->out_start.data();
    if (ship_start) *ship_start = p->ship_start.data();
    if (ship) *ship = reinterpret_cast<const uint32_t *>(p->ship.data());
    if (n_ship) *n_ship = p->ship.size();
    if (n_remote) *n_remote = p->n_remote;
    return SBMBP_OK;
}

int sbmbp_create_dist(sbmbp_plan *p, uint32_t deg_corr_flag, int device, sbmbp_engine **out) {
    if (!p || !out || !p->finished) {
        set_error("plan missing or not finished");
        return SBMBP_ERR_STATE;
    }
    if (deg_corr_flag > 2) {
        set_error("bad deg_corr_flag");
        return SBMBP_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the engine has no CPU fallback");
        return SBMBP_ERR_NODEVICE;
    }
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= ndev || device >= kMaxDevices) {
        set_error("device index out of range");
        return SBMBP_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("device '") + prop.name + "' is not sm_100");
        return SBMBP_ERR_NODEVICE;
    }
    const sbmbp_graph *g = p->g;
    auto *e = new sbmbp_engine();
    e->g = g;
    e->N = g->N;
    e->M = g->M;
    e->N_global = g->N_global;
    e->Q = p->Q;
    e->dc = deg_corr_flag;
    e->prec = p->prec;
    e->qt = p->qt;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->exact_pairs_max_n = exact_pairs_default();
    e->dist = true;
    e->buf_slots = std::max<uint64_t>(e->M, 1);
    e->rank = p->rank;
    e->world = p->world;
    e->fast_path = true;
    if (const char *env = std::getenv("SBMBP_NO_PIPE")) e->pipe_path = std::atoi(env) == 0;
    e->nbuckets = p->nbuckets;
    e->ntiles = unsigned(p->tiles.size());
    const size_t elt = (p->prec == SBMBP_F64) ? 8 : 4;
    const size_t msg_bytes = std::max<size_t>(e->M * e->Q, 1) * elt;
    auto fail = [&](int rc) {
        sbmbp_destroy(e);
        return rc;
    };
#define CREATE_TRY(expr)                                                     \
    do {                                                                     \
        cudaError_t _err = (expr);                                           \
        if (_err != cudaSuccess) {                                           \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_err)); \
            return fail(SBMBP_ERR_CUDA);                                     \
        }                                                                    \
    } while (0)
    CREATE_TRY(cudaMalloc(&e->d_row_ptr, (size_t(e->N) + 1) * sizeof(unsigned long long) + kBulkPad));
    CREATE_TRY(cudaMalloc(&e->d_rev, std::max<size_t>(e->M, 1) * sizeof(unsigned) + kBulkPad));
    CREATE_TRY(cudaMalloc(&e->d_pos, std::max<size_t>(e->M, 1) * sizeof(unsigned) + kBulkPad));
    CREATE_TRY(cudaMalloc(&e->d_info, std::max<size_t>(e->M, 1) * sizeof(unsigned) + kBulkPad));
    CREATE_TRY(cudaMalloc(&e->d_S[0], msg_bytes));
    CREATE_TRY(cudaMalloc(&e->d_S[1], msg_bytes));
    // outbox / mirror: one entry per REMOTE out-message
    CREATE_TRY(cudaMalloc(&e->d_mirror, std::max<size_t>(size_t(p->n_remote) * e->Q, 1) * elt));
    CREATE_TRY(cudaMemset(e->d_mirror, 0, std::max<size_t>(size_t(p->n_remote) * e->Q, 1) * elt));
    CREATE_TRY(cudaMalloc(&e->d_rpos, std::max<size_t>(e->M, 1) * sizeof(unsigned)));
    CREATE_TRY(cudaMalloc(&e->d_out_start, (size_t(p->nsuper) + 1) * sizeof(unsigned)));
    CREATE_TRY(cudaMalloc(&e->d_ship, std::max<size_t>(p->ship.size(), 1) * sizeof(ShipDesc)));
    CREATE_TRY(cudaMalloc(&e->d_ship_start, (size_t(p->nsuper) + 1) * sizeof(unsigned)));
    if (!p->ship.empty())
        CREATE_TRY(cudaMemcpy(e->d_ship, p->ship.data(), p->ship.size() * sizeof(ShipDesc), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(e->d_ship_start, p->ship_start.data(), (size_t(p->nsuper) + 1) * sizeof(unsigned), cudaMemcpyHostToDevice));
    if (const char *env = std::getenv("SBMBP_SHIP_TMA")) e->ship_tma = std::atoi(env);
    CREATE_TRY(cudaMalloc(&e->d_out_rpos, std::max<size_t>(p->n_remote, 1) * sizeof(unsigned)));
    if (p->n_remote)
        CREATE_TRY(cudaMemcpy(e->d_out_rpos, p->out_rpos.data(), size_t(p->n_remote) * sizeof(unsigned), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&e->d_sync, sizeof(SyncBlock)));
    CREATE_TRY(cudaMemset(e->d_sync, 0, sizeof(SyncBlock)));
#ifdef SBMBP_TUNING
    CREATE_TRY(cudaMalloc(&e->d_trace, (256 + 4 * 1024) * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(e->d_trace, 0, (256 + 4 * 1024) * sizeof(unsigned long long)));
#endif
    e->tps = p->tps;
    e->nsuper = p->nsuper;
    e->n_remote = p->n_remote;
    e->sync_peer[p->rank] = e->d_sync;
    CREATE_TRY(cudaMalloc(&e->d_marg, std::max<size_t>(size_t(e->N) * e->Q, 1) * sizeof(double)));
    CREATE_TRY(cudaMalloc(&e->d_tiles, std::max<size_t>(e->ntiles, 1) * sizeof(Tile)));
    CREATE_TRY(cudaMalloc(&e->d_prm, sizeof(DevParams)));
    CREATE_TRY(cudaMalloc(&e->d_field[0], sizeof(Field)));
    CREATE_TRY(cudaMalloc(&e->d_field[1], sizeof(Field)));
    CREATE_TRY(cudaMalloc(&e->d_ctl, sizeof(Ctl)));
    CREATE_TRY(cudaMalloc(&e->d_partial, std::max<size_t>(size_t(e->ntiles) * (e->qt + 1), 1) * sizeof(double)));
    CREATE_TRY(cudaMalloc(&e->d_row, (kMaxQ + 1) * sizeof(double)));
    CREATE_TRY(cudaMalloc(&e->d_out, kOutDoubles * sizeof(double)));
    CREATE_TRY(cudaMallocHost(&e->h_ctl, sizeof(Ctl)));
    CREATE_TRY(cudaMallocHost(&e->h_out, kOutDoubles * sizeof(double)));
    CREATE_TRY(cudaEventCreate(&e->ev0));
    CREATE_TRY(cudaEventCreate(&e->ev1));
    CREATE_TRY(cudaMemcpy(e->d_row_ptr, g->row_ptr.data(), (size_t(e->N) + 1) * sizeof(unsigned long long),
                          cudaMemcpyHostToDevice));
    if (e->M) {
        CREATE_TRY(cudaMemcpy(e->d_rev, p->gather.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(e->d_pos, p->pos.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(e->d_info, p->info.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(e->d_rpos, p->rpos.data(), e->M * sizeof(unsigned), cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaMemcpy(e->d_out_start, p->out_start.data(), (size_t(p->nsuper) + 1) * sizeof(unsigned), cudaMemcpyHostToDevice));
    if (e->ntiles)
        CREATE_TRY(cudaMemcpy(e->d_tiles, p->tiles.data(), p->tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemset(e->d_ctl, 0, sizeof(Ctl)));
    CREATE_TRY(cudaMemset(e->d_field[0], 0, sizeof(Field)));
    CREATE_TRY(cudaMemset(e->d_field[1], 0, sizeof(Field)));
    CREATE_TRY(cudaMemset(e->d_row, 0, (kMaxQ + 1) * sizeof(double)));
#undef CREATE_TRY
    e->peer[0][e->rank] = e->d_S[0];
    e->peer[1][e->rank] = e->d_S[1];
    e->na.assign(e->Q, 0);
    e->cab.assign(size_t(e->Q) * e->Q, 0.0);
    e->eta.assign(e->Q, 0.0);
    *out = e;
    return SBMBP_OK;
}

// 3 x 64 bytes: the CUDA IPC handles of this rank's two message buffers and of its sync block (flags + rows)
int sbmbp_dist_ipc_export(sbmbp_engine *e, void *handles) {
    TRY(need(e, false, false));
    if (!e->dist || !handles) {
        set_error("not a multi-GPU engine");
        return SBMBP_ERR_STATE;
    }
    cudaIpcMemHandle_t h[3];
    CUDA_TRY(cudaIpcGetMemHandle(&h[0], e->d_S[0]));
    CUDA_TRY(cudaIpcGetMemHandle(&h[1], e->d_S[1]));
    CUDA_TRY(cudaIpcGetMemHandle(&h[2], e->d_sync));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handles, h, 192);
    return SBMBP_OK;
}

int sbmbp_dist_ipc_import(sbmbp_engine *e, int peer, const void *handles) {
    TRY(need(e, false, false));
    if (!e->dist || !handles || peer < 0 || peer >= e->world || peer == e->rank) {
        set_error("bad peer");
        return SBMBP_ERR_ARG;
    }
    cudaIpcMemHandle_t h[3];
    std::memcpy(h, handles, 192);
    for (int b = 0; b < 3; ++b) {
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h[b], cudaIpcMemLazyEnablePeerAccess));
        if (b < 2) e->peer[b][peer] = ptr;
        else e->sync_peer[peer] = ptr;
        e->ipc_opened.push_back(ptr);
    }
    return SBMBP_OK;
}

// after every rank has a message state (and a barrier): fetch this rank's out-messages from their owners
int sbmbp_dist_sync_mirror(sbmbp_engine *e) {
    TRY(need(e, false, true));
    if (!e->dist) {
        set_error("not a multi-GPU engine");
        return SBMBP_ERR_STATE;
    }
    for (int k = 0; k < e->world; ++k)
        if (!e->peer[0][k] || !e->peer[1][k]) {
            set_error("peer " + std::to_string(k) + " has not been imported");
            return SBMBP_ERR_STATE;
        }
    if (e->M) {
        PeerTable pt;
        const int b = int(e->sweeps_done & 1u);
        for (int k = 0; k < 8; ++k) pt.p[k] = e->peer[b][k];
        const size_t n = size_t(e->M) * e->Q;
        const unsigned blocks = unsigned(std::min<size_t>((n + 255) / 256, size_t(e->sm_count) * 16));
        if (e->prec == SBMBP_F64)
            mirror_pull_kernel<double><<<blocks, 256, 0, e->stream>>>(static_cast<double *>(e->d_mirror), e->d_pos, e->d_rpos, pt, e->M, e->Q);
        else
            mirror_pull_kernel<float><<<blocks, 256, 0, e->stream>>>(static_cast<float *>(e->d_mirror), e->d_pos, e->d_rpos, pt, e->M, e->Q);
        CUDA_TRY(cudaGetLastError());
        e->stat_launches += 1;
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SBMBP_OK;
}

// this rank's share of init_h: row = [sum_i w_i psi_i^t (t < Q), 0]; device pointer to Q+1 doubles (stride qt+1)
int sbmbp_dist_field_local(sbmbp_engine *e, void **row_dev, uint32_t *ncols) {
    TRY(need(e, true, true));
    const unsigned blocks = std::max(1u, std::min((e->N + kThreads - 1) / kThreads, unsigned(8 * e->sm_count)));
    TRY(ensure_scratch(e, size_t(blocks) * kMaxQ + kMaxQ));
    CUDA_TRY(cudaMemsetAsync(e->d_row, 0, (kMaxQ + 1) * sizeof(double), e->stream));
    field_partial_kernel<<<blocks, kThreads, 0, e->stream>>>(e->d_marg, e->d_row_ptr, e->N, e->Q, e->dc, e->d_scratch);
    TRY(reduce_columns(e, e->d_scratch, blocks, kMaxQ, e->d_scratch + size_t(blocks) * kMaxQ));
    CUDA_TRY(cudaMemcpyAsync(e->d_row, e->d_scratch + size_t(blocks) * kMaxQ, e->Q * sizeof(double),
                             cudaMemcpyDeviceToDevice, e->stream));
    e->stat_launches += 1;
    if (row_dev) *row_dev = e->d_row;
    if (ncols) *ncols = unsigned(e->qt + 1);
    return SBMBP_OK;
}

int sbmbp_dist_arm(sbmbp_engine *e, float crit, uint32_t max_sweeps) {
    TRY(need(e, true, true));
    return arm_ctl(e, crit, max_sweeps);
}

// n sweeps of this rank's nodes, launched back to back with no host in between: every kernel ships its remote
// out-messages to their owners, publishes the rank's row and flag in all ranks' sync blocks and the next one waits for all
// flags before it touches anything (dist_exchange.cuh).  The last sweep stays OPEN (its rows are in the sync block, the
// field / control block are one sweep behind) until sbmbp_dist_close or the next sbmbp_dist_sweeps.
int sbmbp_dist_sweeps(sbmbp_engine *e, uint32_t n, double damping) {
    TRY(need(e, true, true));
    if (!e->dist) {
        set_error("not a multi-GPU engine");
        return SBMBP_ERR_STATE;
    }
    for (int k = 0; k < e->world; ++k)
        if (!e->sync_peer[k] || !e->peer[0][k]) {
            set_error("peer " + std::to_string(k) + " has not been imported");
            return SBMBP_ERR_STATE;
        }
    for (uint32_t i = 0; i < n; ++i) {
        TRY(dispatch(e, [&](auto t, auto qt) { return launch_dist_sweep<decltype(t), decltype(qt)::value>(e, damping); }));
        e->dist_open = true;
        e->dist_seq += 1;
    }
    e->state_version++;
    return SBMBP_OK;  // the sweep statistics are settled at the close (kernels after a converged sweep are no-ops)
}

// Closes the open sweep (waits for every rank's flag on the device, reduces all ranks' rows -> field, control block,
// convergence decision).  With sync != 0 the control block is read back: maxdiff / converged / niter are returned and
// the host's sweep counter is corrected if the batch stopped early because it converged.
int sbmbp_dist_close(sbmbp_engine *e, int sync, double *maxdiff, int *converged, int *niter) {
    TRY(need(e, true, true));
    if (!e->dist) {
        set_error("not a multi-GPU engine");
        return SBMBP_ERR_STATE;
    }
    if (e->dist_open) {
        TRY(dispatch(e, [&](auto t, auto qt) { return launch_dist_close<decltype(qt)::value>(e); }));
        e->dist_open = false;
    }
#ifdef SBMBP_TUNING
    if (e->d_trace && std::getenv("SBMBP_DIST_TRACE")) {
        unsigned long long t[256];
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        CUDA_TRY(cudaMemcpy(t, e->d_trace, sizeof(t), cudaMemcpyDeviceToHost));
        for (unsigned q = 1; q < 64; ++q) {
            const unsigned long long *a = t + 4 * q, *b = t + 4 * (q - 1);
            if (a[0] && b[2] && a[0] > b[0])
                std::fprintf(stderr, "[rank %d] seq%%64=%u: entry +%.1f us after prev entry; waited %.1f us for flags; publish %.1f us after entry\n",
                             e->rank, q, (a[0] - b[0]) * 1e-3, (a[1] - a[0]) * 1e-3, (a[2] - a[0]) * 1e-3);
        }
        CUDA_TRY(cudaMemset(e->d_trace, 0, sizeof(t)));
    }
#endif
    const unsigned counted = e->sweeps_done;
    if (sync) {
        TRY(download_ctl(e));  // also refreshes e->sweeps_done
        e->dist_seq = e->sweeps_done;
        if (maxdiff) *maxdiff = e->h_ctl->last_maxdiff;
        if (converged) *converged = e->h_ctl->converged;
        if (niter) *niter = e->h_ctl->niter;
    } else {
        e->sweeps_done = e->dist_seq;  // valid while no convergence stop is armed (crit < 0)
    }
    e->stat_sweeps += e->sweeps_done - counted;  // sweeps actually executed since the last close
    e->stat_edge_updates += uint64_t(e->sweeps_done - counted) * e->M;
    return SBMBP_OK;
}

// gathered: device pointer to world x (qt+1) doubles, rank-major.  advance = 1 closes a sweep, 0 an init_h.
// With sync != 0 the control block is read back: maxdiff / converged / niter become available.
int sbmbp_dist_finalize(sbmbp_engine *e, const void *gathered_dev, int advance, int sync, double *maxdiff,
                        int *converged, int *niter) {
    TRY(need(e, true, true));
    if (advance) {
        set_error("sbmbp_dist_finalize closes init_h only (advance = 0); sweeps are closed on the device, see sbmbp_dist_close");
        return SBMBP_ERR_ARG;
    }
    bp_finalize_dist_kernel<<<1, 32, 0, e->stream>>>(static_cast<const double *>(gathered_dev), unsigned(e->world),
                                                     unsigned(e->qt + 1), e->Q, e->d_prm, e->d_field[0], e->d_field[1],
                                                     e->d_ctl, advance);
    CUDA_TRY(cudaGetLastError());
    e->stat_launches += 1;
    if (advance) {
        e->state_version++;
        e->stat_sweeps += 1;
        e->stat_edge_updates += e->M;
    } else {
        e->field_valid = true;
    }
    if (sync) {
        TRY(download_ctl(e));
        if (maxdiff) *maxdiff = e->h_ctl->last_maxdiff;
        if (converged) *converged = e->h_ctl->converged;
        if (niter) *niter = e->h_ctl->niter;
    } else if (advance) {
        e->sweeps_done += 1;  // valid while no convergence stop is armed (crit < 0)
    }
    return SBMBP_OK;
}

// local sums for the overlap / EM expectations: row[0..Q) = sum psi, row[kMaxQ..) = sum d psi, then the
// Q x Q confusion matrix at stride kMaxQ (see node_stats_kernel); the caller all-reduces them
// ---- multi-GPU reductions for the free energy and the EM statistics: every call returns this rank's share, the caller
// all-reduces (sum).  deg_corr_flag != 0: the dc terms need the degrees of remote neighbours -> sbmbp_dist_set_degrees.

// Degrees of ALL nodes (deg_global[N_global], the ranks' degree arrays concatenated by the caller): the
// degree-corrected free energy and EM statistics carry d_i d_l per edge (belief_propagation.cpp:464, :584), and l may
// live on another rank.  Builds the per-slot neighbour degree the edge pass reads.
int sbmbp_dist_set_degrees(sbmbp_engine *e, const uint32_t *deg_global) {
    TRY(need(e, false, false));
    if (!e->dist || !deg_global) {
        set_error("sbmbp_dist_set_degrees: multi-GPU engine and a degree array needed");
        return SBMBP_ERR_ARG;
    }
    const auto &g = *e->g;
    std::vector<unsigned> degn(std::max<uint64_t>(e->M, 1), 0u);
    for (uint64_t s = 0; s < e->M; ++s) {
        if (g.col[s] >= e->N_global) {
            set_error("neighbour id out of range");
            return SBMBP_ERR_RANGE;
        }
        degn[s] = deg_global[g.col[s]];
    }
    if (!e->d_degsrc) CUDA_TRY(cudaMalloc(&e->d_degsrc, degn.size() * sizeof(unsigned)));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaMemcpy(e->d_degsrc, degn.data(), degn.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    e->state_version++;
    return SBMBP_OK;
}

// edge pass over this rank's rows: row = [f_site, f_edge, entropy_site, entropy_edge, cab two-point sums (qt x qt)] as sums
int sbmbp_dist_energy_local(sbmbp_engine *e, int which, double *row, uint32_t cap, uint32_t *ncols) {
    TRY(need(e, true, true));
    if (!e->dist) {
        set_error("sbmbp_dist_energy_local: multi-GPU engine only");
        return SBMBP_ERR_UNSUPPORTED;
    }
    if (e->dc != 0 && !e->d_degsrc) {
        set_error("sbmbp_dist_energy_local: deg_corr_flag != 0 needs sbmbp_dist_set_degrees first");
        return SBMBP_ERR_STATE;
    }
    if (!e->field_valid) {
        set_error("the field is not current: run init_h / a sweep first");
        return SBMBP_ERR_STATE;
    }
    if (which == 1 && !(e->dc == 0 && e->beta != 1.0)) which = 0;
    std::vector<double> out;
    TRY(dispatch(e, [&](auto t, auto qt) { return launch_energy<decltype(t), decltype(qt)::value>(e, which, out); }));
    if (ncols) *ncols = uint32_t(out.size());
    if (row) std::copy(out.begin(), out.begin() + std::min<size_t>(out.size(), cap), row);
    return SBMBP_OK;
}

// T_k = sum over this rank's nodes of psi_i^{(x) k} (Q^order entries, first digit fastest)
int sbmbp_dist_moment_local(sbmbp_engine *e, uint32_t order, double *T, uint64_t cap) {
    TRY(need(e, false, true));
    std::vector<double> t;
    TRY(moment_tensor(e, order, t));
    if (T) std::copy(t.begin(), t.begin() + std::min<uint64_t>(t.size(), cap), T);
    return SBMBP_OK;
}

// sum over this rank's directed edges (i, l) of log1p(-psi_i^T W1 psi_l) (mode 1, series form of the non-edge term) or
// log(psi_i^T W psi_l) (mode 0); marg_global_dev: device pointer to the all-gathered marginals [N_global][Q]
int sbmbp_dist_edge_pairs_local(sbmbp_engine *e, const void *marg_global_dev, int mode, double *result) {
    TRY(need(e, true, true));
    if (!e->dist || !marg_global_dev || !result || (mode != 0 && mode != 1)) {
        set_error("bad argument");
        return SBMBP_ERR_ARG;
    }
    return edge_pairs_sum(e, mode == 0 ? e->d_prm->W : e->d_prm->W1, nullptr, mode, result,
                          static_cast<const double *>(marg_global_dev));
}

int sbmbp_dist_node_stats(sbmbp_engine *e, const uint32_t *true_conf_local, double *row, uint32_t *ncols) {
    TRY(need(e, false, true));
    std::vector<double> r;
    TRY(node_stats(e, true_conf_local, r));
    if (row) std::copy(r.begin(), r.end(), row);
    if (ncols) *ncols = kNodeCols;
    return SBMBP_OK;
}

}  // extern "C"
