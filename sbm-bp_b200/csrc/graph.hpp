// Host-side graph of the BP engine: destination-sorted CSR with a reverse-edge index.
// Replaces edge_list_t / adj_list_t (reference types.h:12-15) and the flattened copies
// graph_neis_ / graph_neis_inv_ (belief_propagation.cpp:246-266).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct sbmbp_graph {
    uint32_t N = 0;
    uint64_t M = 0;            // directed edges
    uint64_t E = 0;            // blockmodel_t::get_E(): M / 2
    uint32_t max_degree = 0;   // blockmodel_t::get_graph_max_degree()
    std::vector<uint64_t> row_ptr;  // N + 1
    std::vector<uint32_t> col;      // M: neighbours of i ascending == graph_neis_[i]
    std::vector<uint32_t> rev;      // M: slot of the reverse edge == row_ptr[col[e]] + graph_neis_inv_[i][l]
    std::vector<uint32_t> deg;      // N
    // rank-local slice of a larger graph (multi-GPU): rows of the global nodes [node_lo, node_lo + N); col holds
    // GLOBAL neighbour ids and rev is empty.  N_global == 0 marks a complete graph.
    uint32_t node_lo = 0, N_global = 0;
};

namespace sbmbp {

void set_error(const std::string &msg);
const char *get_error();

// returns SBMBP_* status
int parse_edgelist(const char *path, std::vector<uint32_t> &u, std::vector<uint32_t> &v);
int build_graph(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N, sbmbp_graph &g);
// rows of the nodes [lo, hi) only; pairs without an endpoint in the range are ignored
int build_graph_range(const uint32_t *u, const uint32_t *v, uint64_t n_pairs, uint32_t N_global, uint32_t lo,
                      uint32_t hi, sbmbp_graph &g);

}  // namespace sbmbp
