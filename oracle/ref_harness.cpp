// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI harness around the UNMODIFIED reference sources.  oracle/Makefile compiles
// /root/reference/src/{belief_propagation,blockmodel,graph_utilities,output_functions}.cpp
// in place (nothing is copied) and links them with this file into
// oracle/_ref/libsbmbp_ref.so.  The harness replays the object wiring of
// reference src/main.cpp:236-365 without Boost, and reaches the protected engine state
// (belief_propagation.h:19-87) through a subclass.  Only tests/, smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load the result.
//
// What it offers on top of the reference's public methods:
//   * CSR views of graph_neis_ / graph_neis_inv_ (belief_propagation.cpp:246-266)
//   * get/set of mmap_, real_psi_, h_ in the reference's (i, l, q) order
//   * a "Jacobi sweep by the reference's own arithmetic": every node is updated with the
//     reference routine the converge() loop would pick (belief_propagation.cpp:397-401)
//     from one frozen snapshot, and only the slots that call touched are restored
//     (SURVEY.md section 8c).  This is the level-1 parity target for the synchronous GPU sweep.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <streambuf>
#include <string>
#include <vector>
#include <set>
#include <utility>
#include <algorithm>
#include <fstream>
#include <chrono>

// belief_propagation.h:148-165 keeps the free-energy pieces private; the harness needs them
// one by one.  All std headers the reference includes are already included above, so the
// macro only affects the reference's own class definitions.
#define private protected
#include "belief_propagation.h"
#include "graph_utilities.h"
#undef private

namespace {

struct null_buf : std::streambuf {
    int overflow(int c) override { return c; }
};

template <class Base>
struct engine_access : Base {
    using Base::N_;
    using Base::Q_;
    using Base::mmap_;
    using Base::real_psi_;
    using Base::h_;
    using Base::exph_;
    using Base::graph_neis_;
    using Base::graph_neis_inv_;
    using Base::adj_list_ptr_;
    using Base::cab_;
    using Base::na_;
    using Base::eta_;
    using Base::na_expect_;
    using Base::nna_expect_;
    using Base::cab_expect_;
    using Base::conf_planted_;
    using Base::LARGE_DEGREE;
    using Base::init_h;
    using Base::learning_step;
    using Base::compute_na_expect;
    using Base::compute_cab_expect;
    using Base::compute_f_site;
    using Base::compute_f_edge;
    using Base::compute_f_non_edge;
    using Base::compute_entropy_site;
    using Base::compute_entropy_edge;
    using Base::compute_entropy_non_edge;
};

struct ref_handle {
    adj_list_t adj;
    uint_vec_t memberships;
    std::unique_ptr<blockmodel_t> bm;
    bool learn_mode = false;
    std::unique_ptr<belief_propagation> algo;  // bp_basic (learn) or bp_conditional (infer), main.cpp:318-323
    bp_blockmodel_state state;
    std::mt19937 engine;
    std::vector<uint64_t> row_ptr;
    null_buf nb;
    std::streambuf *old_clog = nullptr;

    // both subclasses only add a virtual override, so the protected layout is the base's
    engine_access<belief_propagation> &acc() {
        return *static_cast<engine_access<belief_propagation> *>(algo.get());
    }
};

void build_common(ref_handle *h, const edge_list_t &el, const uint32_t *block_sizes, uint32_t Q,
                  uint32_t dc_flag, int learn_mode) {
    // main.cpp:239-252 (memberships from -n), :271-281 (N, adjacency), :296 (blockmodel)
    uint32_t N = 0;
    for (uint32_t r = 0; r < Q; ++r) N += block_sizes[r];
    h->memberships.assign(N, 0);
    uint32_t shift = 0;
    for (uint32_t r = 0; r < Q; ++r) {
        for (uint32_t i = 0; i < block_sizes[r]; ++i) h->memberships[shift + i] = r;
        shift += block_sizes[r];
    }
    h->adj = edge_to_adj(el, N);
    h->bm.reset(new blockmodel_t(h->memberships, Q, (unsigned)h->adj.size(), dc_flag, &h->adj));
    h->learn_mode = learn_mode != 0;
    if (learn_mode) h->algo.reset(new bp_basic());
    else h->algo.reset(new bp_conditional());
}

void fill_row_ptr(ref_handle *h) {
    auto &a = h->acc();
    h->row_ptr.assign((size_t)a.N_ + 1, 0);
    for (uint32_t i = 0; i < a.N_; ++i) h->row_ptr[i + 1] = h->row_ptr[i] + a.graph_neis_[i].size();
}

}  // namespace

extern "C" {

void *ref_create_from_pairs(const uint32_t *u, const uint32_t *v, uint64_t n_pairs,
                            const uint32_t *block_sizes, uint32_t Q, uint32_t dc_flag, int learn_mode) {
    auto *h = new ref_handle();
    h->old_clog = std::clog.rdbuf(&h->nb);  // the >=50-degree path logs one line per call (:816)
    edge_list_t el;
    el.reserve(n_pairs);
    for (uint64_t k = 0; k < n_pairs; ++k) el.push_back(std::make_pair(u[k], v[k]));
    build_common(h, el, block_sizes, Q, dc_flag, learn_mode);
    return h;
}

void *ref_create_from_file(const char *path, const uint32_t *block_sizes, uint32_t Q, uint32_t dc_flag,
                           int learn_mode) {
    auto *h = new ref_handle();
    h->old_clog = std::clog.rdbuf(&h->nb);
    edge_list_t el;
    load_edge_list(el, path);  // return value ignored, as main.cpp:278 does
    build_common(h, el, block_sizes, Q, dc_flag, learn_mode);
    return h;
}

void ref_destroy(void *hp) {
    auto *h = static_cast<ref_handle *>(hp);
    if (h->old_clog) std::clog.rdbuf(h->old_clog);
    delete h;
}

// raw edge list as load_edge_list sees it (quirks of graph_utilities.cpp:42-58 included)
uint64_t ref_load_edge_list(const char *path, uint32_t *u, uint32_t *v, uint64_t cap) {
    edge_list_t el;
    load_edge_list(el, path);
    for (uint64_t k = 0; k < el.size() && k < cap; ++k) {
        u[k] = el[k].first;
        v[k] = el[k].second;
    }
    return el.size();
}

// main.cpp:338-340.  flag 0 only here; conf is empty, true_conf = block ordering (main.cpp:284-286).
void ref_init_messages(void *hp, uint32_t seed, double beta) {
    auto *h = static_cast<ref_handle *>(hp);
    h->engine.seed(seed);
    int_vec_t beliefs;
    h->algo->init_messages(*h->bm, 0, beliefs, h->memberships, h->engine);
    h->algo->init_special_needs(false);
    h->algo->set_beta(beta);
    fill_row_ptr(h);
}

// --mb_rand (main.cpp:185-187, :299-301): blockmodel_t::shuffle permutes the memberships with the run's engine BEFORE
// init_messages draws from it; true_conf was taken from the ordered memberships earlier (:284-286).
void ref_init_messages_mb_rand(void *hp, uint32_t seed, double beta) {
    auto *h = static_cast<ref_handle *>(hp);
    h->engine.seed(seed);
    h->bm->shuffle(h->engine);
    int_vec_t beliefs;
    h->algo->init_messages(*h->bm, 0, beliefs, h->memberships, h->engine);
    h->algo->init_special_needs(false);
    h->algo->set_beta(beta);
    fill_row_ptr(h);
}

// main.cpp:325-340 with an explicit beliefs vector (what --beliefs_path / -f produce); flags 0-3.
// Flags 2 and 3 assert(conf_planted_[i] != 1) (belief_propagation.cpp:179,:197): refuse instead of aborting the test run.
int ref_init_messages_flag(void *hp, uint32_t flag, const int32_t *conf, uint32_t seed, double beta) {
    auto *h = static_cast<ref_handle *>(hp);
    const uint32_t N = h->bm->get_N();
    if (flag > 3) return -2;
    if (flag >= 2)
        for (uint32_t i = 0; i < N; ++i)
            if (conf[i] == 1) return -1;
    h->engine.seed(seed);
    int_vec_t beliefs;
    if (flag != 0) beliefs.assign(conf, conf + N);
    // a fresh engine object, as in the binary (one init_messages per run): bp_allocate's resize() keeps the values of an
    // earlier initialisation, and flag 3 leaves the odd-ranked slots untouched
    if (h->learn_mode) h->algo.reset(new bp_basic());
    else h->algo.reset(new bp_conditional());
    h->algo->init_messages(*h->bm, flag, beliefs, h->memberships, h->engine);
    h->algo->init_special_needs(false);
    h->algo->set_beta(beta);
    fill_row_ptr(h);
    return 0;
}

void ref_get_conf_planted(void *hp, int32_t *conf) {
    auto &a = static_cast<ref_handle *>(hp)->acc();
    for (uint32_t i = 0; i < a.N_; ++i) conf[i] = a.conf_planted_[i];
}

uint32_t ref_N(void *hp) { return static_cast<ref_handle *>(hp)->acc().N_; }
uint32_t ref_Q(void *hp) { return static_cast<ref_handle *>(hp)->acc().Q_; }
uint64_t ref_M(void *hp) { return static_cast<ref_handle *>(hp)->row_ptr.back(); }
uint32_t ref_E(void *hp) { return static_cast<ref_handle *>(hp)->bm->get_E(); }
uint32_t ref_max_degree(void *hp) { return static_cast<ref_handle *>(hp)->bm->get_graph_max_degree(); }

// CSR view: col = graph_neis_, rev = row_ptr[j] + graph_neis_inv_ (global slot of the reverse edge)
void ref_get_csr(void *hp, uint64_t *row_ptr, uint32_t *col, uint32_t *rev_local, uint64_t *rev_global) {
    auto *h = static_cast<ref_handle *>(hp);
    auto &a = h->acc();
    for (uint32_t i = 0; i <= a.N_; ++i) row_ptr[i] = h->row_ptr[i];
    for (uint32_t i = 0; i < a.N_; ++i) {
        for (size_t l = 0; l < a.graph_neis_[i].size(); ++l) {
            uint64_t e = h->row_ptr[i] + l;
            uint32_t j = a.graph_neis_[i][l];
            col[e] = j;
            if (rev_local) rev_local[e] = a.graph_neis_inv_[i][l];
            if (rev_global) rev_global[e] = h->row_ptr[j] + a.graph_neis_inv_[i][l];
        }
    }
}

// main.cpp:345-353 + belief_propagation.cpp:290-317
void ref_set_params_direct(void *hp, const double *pa, const double *cab_upper) {
    auto *h = static_cast<ref_handle *>(hp);
    uint32_t Q = h->bm->get_Q();
    double_vec_t pav(pa, pa + Q), cabv(cab_upper, cab_upper + (size_t)Q * (Q + 1) / 2);
    h->state = bp_param_from_direct(*h->bm, pav, cabv);
    h->algo->expand_bp_params(h->state);
}

void ref_set_params_epsilon_c(void *hp, double eps, double c) {
    auto *h = static_cast<ref_handle *>(hp);
    h->state = bp_param_from_epsilon_c(*h->bm, eps, c);
    h->algo->expand_bp_params(h->state);
}

void ref_set_params_raw(void *hp, const uint32_t *na, const double *cab) {
    auto *h = static_cast<ref_handle *>(hp);
    uint32_t Q = h->bm->get_Q();
    h->state.na.assign(na, na + Q);
    h->state.cab.assign(Q, std::vector<double>(Q));
    for (uint32_t a = 0; a < Q; ++a)
        for (uint32_t b = 0; b < Q; ++b) h->state.cab[a][b] = cab[a * Q + b];
    h->algo->expand_bp_params(h->state);
}

void ref_get_params(void *hp, uint32_t *na, double *cab, double *eta) {
    auto &a = static_cast<ref_handle *>(hp)->acc();
    for (uint32_t q = 0; q < a.Q_; ++q) {
        if (na) na[q] = a.na_[q];
        if (eta) eta[q] = a.eta_[q];
        for (uint32_t t = 0; t < a.Q_; ++t)
            if (cab) cab[q * a.Q_ + t] = a.cab_[q][t];
    }
}

void ref_set_beta(void *hp, double beta) { static_cast<ref_handle *>(hp)->algo->set_beta(beta); }

// state in the reference's own order: msg[(row_ptr[i]+l)*Q+q] = mmap_[i][l][q], marg[i*Q+q] = real_psi_[i][q]
void ref_get_state(void *hp, double *msg, double *marg, double *hq) {
    auto *h = static_cast<ref_handle *>(hp);
    auto &a = h->acc();
    uint32_t Q = a.Q_;
    for (uint32_t i = 0; i < a.N_; ++i) {
        if (marg)
            for (uint32_t q = 0; q < Q; ++q) marg[(size_t)i * Q + q] = a.real_psi_[i][q];
        if (msg)
            for (size_t l = 0; l < a.mmap_[i].size(); ++l)
                for (uint32_t q = 0; q < Q; ++q) msg[(h->row_ptr[i] + l) * Q + q] = a.mmap_[i][l][q];
    }
    if (hq)
        for (uint32_t q = 0; q < Q; ++q) hq[q] = a.h_[q];
}

void ref_set_state(void *hp, const double *msg, const double *marg) {
    auto *h = static_cast<ref_handle *>(hp);
    auto &a = h->acc();
    uint32_t Q = a.Q_;
    for (uint32_t i = 0; i < a.N_; ++i) {
        if (marg)
            for (uint32_t q = 0; q < Q; ++q) a.real_psi_[i][q] = marg[(size_t)i * Q + q];
        if (msg)
            for (size_t l = 0; l < a.mmap_[i].size(); ++l)
                for (uint32_t q = 0; q < Q; ++q) a.mmap_[i][l][q] = msg[(h->row_ptr[i] + l) * Q + q];
    }
    a.init_h();
}

void ref_init_h(void *hp) { static_cast<ref_handle *>(hp)->acc().init_h(); }

// One synchronous sweep computed by the reference's own node-update routines from a frozen state.
// new_msg is indexed like the state (slot row_ptr[i2]+l2 that the update of i wrote,
// belief_propagation.cpp:1057-1066); node_diff[i] is the value the routine returned.
// The handle's state is left untouched.  Returns the sweep max-diff.
double ref_jacobi_sweep(void *hp, double damping, double *new_msg, double *new_marg, double *node_diff) {
    auto *h = static_cast<ref_handle *>(hp);
    auto &a = h->acc();
    const uint32_t Q = a.Q_;
    a.init_h();
    const double_vec_t h0 = a.h_, exph0 = a.exph_;
    double maxdiff = -100.0;
    std::vector<double> saved;
    for (uint32_t i = 0; i < a.N_; ++i) {
        const size_t d = a.graph_neis_[i].size();
        saved.resize((d + 1) * Q);
        for (size_t l = 0; l < d; ++l) {
            const auto &slot = a.mmap_[a.graph_neis_[i][l]][a.graph_neis_inv_[i][l]];
            for (uint32_t q = 0; q < Q; ++q) saved[l * Q + q] = slot[q];
        }
        for (uint32_t q = 0; q < Q; ++q) saved[d * Q + q] = a.real_psi_[i][q];

        double diff;  // dispatch of belief_propagation.cpp:397-401
        if (d >= a.LARGE_DEGREE) diff = h->algo->bp_iter_update_psi_large_degree(i, damping);
        else diff = h->algo->bp_iter_update_psi(i, damping);
        if (diff > maxdiff) maxdiff = diff;
        if (node_diff) node_diff[i] = diff;

        for (size_t l = 0; l < d; ++l) {
            auto &slot = a.mmap_[a.graph_neis_[i][l]][a.graph_neis_inv_[i][l]];
            const uint64_t g = h->row_ptr[a.graph_neis_[i][l]] + a.graph_neis_inv_[i][l];
            for (uint32_t q = 0; q < Q; ++q) {
                if (new_msg) new_msg[g * Q + q] = slot[q];
                slot[q] = saved[l * Q + q];
            }
        }
        for (uint32_t q = 0; q < Q; ++q) {
            if (new_marg) new_marg[(size_t)i * Q + q] = a.real_psi_[i][q];
            a.real_psi_[i][q] = saved[d * Q + q];
        }
        a.h_ = h0;
        a.exph_ = exph0;
    }
    return maxdiff;
}

// the reference's own random-sequential converge(); returns niter (belief_propagation.cpp:386-415)
int ref_converge(void *hp, float crit, uint32_t max_iter, float damping) {
    auto *h = static_cast<ref_handle *>(hp);
    return h->algo->converge(crit, max_iter, damping, h->engine);
}

// same, with the wall time of converge() alone (the CPU baseline of SURVEY.md section 8d)
int ref_converge_timed(void *hp, float crit, uint32_t max_iter, float damping, double *seconds) {
    auto *h = static_cast<ref_handle *>(hp);
    auto t0 = std::chrono::steady_clock::now();
    int it = h->algo->converge(crit, max_iter, damping, h->engine);
    auto t1 = std::chrono::steady_clock::now();
    *seconds = std::chrono::duration<double>(t1 - t0).count();
    return it;
}

double ref_free_energy(void *hp) { return static_cast<ref_handle *>(hp)->algo->compute_free_energy(); }
double ref_f_site(void *hp) { return static_cast<ref_handle *>(hp)->acc().compute_f_site(); }
double ref_f_edge(void *hp) { return static_cast<ref_handle *>(hp)->acc().compute_f_edge(); }
double ref_f_non_edge(void *hp) { return static_cast<ref_handle *>(hp)->acc().compute_f_non_edge(); }
double ref_entropy(void *hp) { return static_cast<ref_handle *>(hp)->algo->compute_entropy(); }
double ref_entropy_site(void *hp) { return static_cast<ref_handle *>(hp)->acc().compute_entropy_site(); }
double ref_entropy_edge(void *hp) { return static_cast<ref_handle *>(hp)->acc().compute_entropy_edge(); }
double ref_entropy_non_edge(void *hp) { return static_cast<ref_handle *>(hp)->acc().compute_entropy_non_edge(); }
double ref_overlap(void *hp) { return static_cast<ref_handle *>(hp)->algo->compute_overlap(); }

void ref_em_stats(void *hp, double *na_exp, double *nna_exp, double *cab_exp) {
    auto &a = static_cast<ref_handle *>(hp)->acc();
    a.compute_na_expect();
    a.compute_cab_expect();
    for (uint32_t q = 0; q < a.Q_; ++q) {
        na_exp[q] = a.na_expect_[q];
        nna_exp[q] = a.nna_expect_[q];
        for (uint32_t t = 0; t < a.Q_; ++t) cab_exp[q * a.Q_ + t] = a.cab_expect_[q][t];
    }
}

void ref_learning_step(void *hp, float lr) { static_cast<ref_handle *>(hp)->acc().learning_step(lr); }

// the reference's learning() (belief_propagation.cpp:14-51); its stdout lines are swallowed and the
// learned parameters are read back from the engine instead.
void ref_learning(void *hp, float crit, uint32_t max_time, float lr, float damping, uint32_t *na, double *cab,
                  double *eta) {
    auto *h = static_cast<ref_handle *>(hp);
    null_buf sink;
    std::streambuf *old = std::cout.rdbuf(&sink);
    h->algo->learning(*h->bm, h->state, crit, max_time, lr, damping, h->engine);
    std::cout.rdbuf(old);
    ref_get_params(hp, na, cab, eta);
}

// the reference's inference() (belief_propagation.cpp:77-99) with its stdout line captured verbatim
int ref_inference(void *hp, float crit, uint32_t max_time, float damping, char *out, int out_cap) {
    auto *h = static_cast<ref_handle *>(hp);
    std::ostringstream oss;
    std::streambuf *old = std::cout.rdbuf(oss.rdbuf());
    h->algo->inference(*h->bm, h->state, crit, max_time, damping, h->engine);
    std::cout.rdbuf(old);
    std::string s = oss.str();
    int n = (int)std::min<size_t>(s.size(), (size_t)out_cap - 1);
    std::memcpy(out, s.data(), n);
    out[n] = 0;
    return (int)s.size();
}

}  // extern "C"
