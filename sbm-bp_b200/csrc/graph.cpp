// Host graph builder (C++17, no CUDA).  See graph.hpp and include/sbmbp.h.
//
// Reference behaviour reproduced bit-exactly on well-formed input:
//   * load_edge_list (graph_utilities.cpp:42-58): one "u v" pair per line, any blank-separated columns
//     after the second are ignored.  The reference re-pushes the previous pair on a blank line, which
//     the set-based adjacency then merges away; here blank lines are simply skipped.  A line that does
//     not start with two unsigned integers makes the reference invent an edge (0, stale v); here it is
//     SBMBP_ERR_PARSE.
//   * edge_to_adj (graph_utilities.cpp:60-77): symmetric std::set per vertex == neighbours ascending,
//     duplicates merged, a self-loop kept once.  Ids >= N make the reference index out of bounds later
//     (belief_propagation.cpp:125); here they are SBMBP_ERR_RANGE.
//   * bp_allocate (belief_propagation.cpp:246-266): graph_neis_inv_[i][l] = rank of i in the set of
//     j = graph_neis_[i][l].  Stored as rev[e] = row_ptr[j] + rank.
//   * blockmodel_t ctor (blockmodel.cpp:27-46): degrees, max degree, E = (sum of degrees) / 2.
#include "graph.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>

#include "../../include/sbmbp.h"

namespace sbmbp {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
const char *get_error() { return g_error.c_str(); }

static inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// parses one unsigned 32-bit integer at p (after blanks); returns false when none is there
static inline bool parse_u32(const char *&p, const char *end, uint32_t &out) {
    while (p < end && is_blank(*p)) ++p;
    const char *d = p;
    if (d < end && *d == '+') ++d;
    if (d >= end || *d < '0' || *d > '9') return false;
    uint64_t v = 0;
    while (d < end && *d >= '0' && *d <= '9') {
        v = v * 10 + uint64_t(*d - '0');
        if (v > 0xffffffffull) return false;
        ++d;
    }
    if (d < end && !is_blank(*d) && *d != '\n') return false;  // "12abc"
    out = uint32_t(v);
    p = d;
    return true;
}

int parse_edgelist(const char *path, std::vector<uint32_t> &u, std::vector<uint32_t> &v) {
    u.clear();
    v.clear();
    FILE *f = std::fopen(path, "rb");
    if (!f) {
        set_error(std::string("cannot open edge list: ") + path);
        return SBMBP_ERR_IO;
    }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf(size_t(sz) + 1);
    size_t got = sz > 0 ? std::fread(buf.data(), 1, size_t(sz), f) : 0;
    std::fclose(f);
    buf[got] = '\n';
    const char *p = buf.data(), *end = buf.data() + got;
    u.reserve(got / 8);
    v.reserve(got / 8);
    uint64_t line_no = 0;
    while (p < end) {
        ++line_no;
        const char *eol = static_cast<const char *>(std::memchr(p, '\n', size_t(end - p)));
        if (!eol) eol = end;
        const char *q = p;
        while (q < eol && is_blank(*q)) ++q;
        if (q < eol) {
            uint32_t a, b;
            if (!parse_u32(q, eol, a) || !parse_u32(q, eol, b)) {
                set_error(std::string(path) + ": line " + std::to_string(line_no) +
                          " is not 'u v' (the reference would silently invent an edge here)");
                return SBMBP_ERR_PARSE;
            }
            u.push_back(a);
            v.push_back(b);
        }
        p = eol + 1;
    }
    return SBMBP_OK;
}

int build_graph(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N, sbmbp_graph &g) {
    g = sbmbp_graph();
    g.N = N;
    // 1. raw directed counts; a self-loop contributes one entry (std::set semantics)
    std::vector<uint64_t> off(size_t(N) + 1, 0);
    for (uint64_t k = 0; k < n_pairs; ++k) {
        if (u[k] >= N || v[k] >= N) {
            set_error("vertex id " + std::to_string(std::max(u[k], v[k])) + " >= N = " + std::to_string(N) +
                      " at pair " + std::to_string(k));
            return SBMBP_ERR_RANGE;
        }
        off[size_t(u[k]) + 1]++;
        if (u[k] != v[k]) off[size_t(v[k]) + 1]++;
    }
    for (uint32_t i = 0; i < N; ++i) off[i + 1] += off[i];
    const uint64_t total = off[N];
    if (total >= 0xffffffffull) {
        set_error("more than 2^32-1 directed edges: rev is a 32-bit slot index");
        return SBMBP_ERR_UNSUPPORTED;
    }
    // 2. scatter
    std::vector<uint32_t> raw(total);
    {
        std::vector<uint64_t> cur(off.begin(), off.end() - 1);
        for (uint64_t k = 0; k < n_pairs; ++k) {
            raw[cur[u[k]]++] = v[k];
            if (u[k] != v[k]) raw[cur[v[k]]++] = u[k];
        }
    }
    // 3. sort + unique each row, in parallel over node ranges of equal raw size
    g.deg.assign(N, 0);
    unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (total < (1u << 20)) nthreads = 1;
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            uint32_t *b = raw.data() + off[i], *e = raw.data() + off[i + 1];
            std::sort(b, e);
            g.deg[i] = uint32_t(std::unique(b, e) - b);
        }
    };
    if (nthreads == 1) {
        work(0, N);
    } else {
        std::vector<std::thread> pool;
        uint32_t lo = 0;
        for (unsigned t = 0; t < nthreads; ++t) {
            uint64_t target = total / nthreads * (t + 1);
            uint32_t hi = (t + 1 == nthreads)
                              ? N
                              : uint32_t(std::upper_bound(off.begin(), off.end(), target) - off.begin() - 1);
            if (hi < lo) hi = lo;
            if (hi > N) hi = N;
            pool.emplace_back(work, lo, hi);
            lo = hi;
        }
        for (auto &th : pool) th.join();
    }
    // 4. compact
    g.row_ptr.assign(size_t(N) + 1, 0);
    for (uint32_t i = 0; i < N; ++i) {
        g.row_ptr[i + 1] = g.row_ptr[i] + g.deg[i];
        if (g.deg[i] > g.max_degree) g.max_degree = g.deg[i];
    }
    g.M = g.row_ptr[N];
    g.E = g.M / 2;
    g.col.resize(g.M);
    for (uint32_t i = 0; i < N; ++i)
        std::copy(raw.data() + off[i], raw.data() + off[i] + g.deg[i], g.col.data() + g.row_ptr[i]);
    raw.clear();
    raw.shrink_to_fit();
    // 5. reverse index.  Rows are ascending and i runs ascending, so the rank of i inside row j equals the
    //    number of times j has been met as a neighbour so far.
    g.rev.resize(g.M);
    {
        std::vector<uint32_t> seen(N, 0);
        for (uint32_t i = 0; i < N; ++i)
            for (uint64_t e = g.row_ptr[i]; e < g.row_ptr[i + 1]; ++e) {
                uint32_t j = g.col[e];
                g.rev[e] = uint32_t(g.row_ptr[j] + seen[j]++);
            }
    }
    return SBMBP_OK;
}

int build_graph_range(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N_global, uint32_t lo,
                      uint32_t hi, sbmbp_graph &g) {
    g = sbmbp_graph();
    if (lo > hi || hi > N_global) {
        set_error("bad node range");
        return SBMBP_ERR_ARG;
    }
    const uint32_t N = hi - lo;
    g.N = N;
    g.node_lo = lo;
    g.N_global = N_global;
    std::vector<uint64_t> off(size_t(N) + 1, 0);
    auto in = [&](uint32_t x) { return x >= lo && x < hi; };
    for (uint64_t k = 0; k < n_pairs; ++k) {
        if (u[k] >= N_global || v[k] >= N_global) {
            set_error("vertex id >= N at pair " + std::to_string(k));
            return SBMBP_ERR_RANGE;
        }
        if (in(u[k])) off[size_t(u[k] - lo) + 1]++;
        if (u[k] != v[k] && in(v[k])) off[size_t(v[k] - lo) + 1]++;
    }
    for (uint32_t i = 0; i < N; ++i) off[i + 1] += off[i];
    const uint64_t total = off[N];
    if (total >= (1ull << 29)) {
        set_error("more than 2^29 in-edges on one rank: positions carry the owner rank in their top 3 bits");
        return SBMBP_ERR_UNSUPPORTED;
    }
    std::vector<uint32_t> raw(total);
    {
        std::vector<uint64_t> cur(off.begin(), off.end() - 1);
        for (uint64_t k = 0; k < n_pairs; ++k) {
            if (in(u[k])) raw[cur[u[k] - lo]++] = v[k];
            if (u[k] != v[k] && in(v[k])) raw[cur[v[k] - lo]++] = u[k];
        }
    }
    g.deg.assign(N, 0);
    unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (total < (1u << 20)) nthreads = 1;
    auto work = [&](uint32_t a, uint32_t b) {
        for (uint32_t i = a; i < b; ++i) {
            uint32_t *x = raw.data() + off[i], *y = raw.data() + off[i + 1];
            std::sort(x, y);
            g.deg[i] = uint32_t(std::unique(x, y) - x);
        }
    };
    {
        std::vector<std::thread> pool;
        const uint32_t per = (N + nthreads - 1) / nthreads;
        for (unsigned t = 0; t < nthreads; ++t) {
            const uint32_t a = std::min(N, t * per), b = std::min(N, a + per);
            if (a < b) pool.emplace_back(work, a, b);
        }
        for (auto &th : pool) th.join();
    }
    g.row_ptr.assign(size_t(N) + 1, 0);
    for (uint32_t i = 0; i < N; ++i) {
        g.row_ptr[i + 1] = g.row_ptr[i] + g.deg[i];
        if (g.deg[i] > g.max_degree) g.max_degree = g.deg[i];
    }
    g.M = g.row_ptr[N];
    g.E = g.M / 2;
    g.col.resize(g.M);
    for (uint32_t i = 0; i < N; ++i)
        std::copy(raw.data() + off[i], raw.data() + off[i] + g.deg[i], g.col.data() + g.row_ptr[i]);
    return SBMBP_OK;
}

}  // namespace sbmbp
