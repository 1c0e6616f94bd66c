// C++ host mirror of the reference's object model over the C ABI of libsbmbp.so.
// Same names, argument meaning and call order as the reference classes that src/main.cpp:277-365 wires
// together (blockmodel.h:10-108, belief_propagation.h:18-178); all numerics live behind sbmbp.h.
#pragma once
#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sbmbp.h"

namespace sbmbp_host {

using uint_vec_t = std::vector<unsigned int>;
using double_vec_t = std::vector<double>;

struct error : std::runtime_error {
    int code;
    error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline void check(int rc) {
    if (rc != SBMBP_OK) throw error(rc, sbmbp_last_error());
}

// types.h:23-26
struct bp_blockmodel_state {
    double_vec_t cab;  // Q*Q row-major
    uint_vec_t na;
};

// Graph and block sizes: the part of blockmodel_t the BP path reads (get_N/Q/E/graph_max_degree/deg_corr_flag)
class blockmodel_t {
public:
    blockmodel_t(const uint_vec_t &n, const std::string &edge_list_path, unsigned int deg_corr_flag)
        : n_(n), deg_corr_flag_(deg_corr_flag) {
        N_ = 0;
        for (auto s : n) N_ += s;
        check(sbmbp_graph_from_edgelist(edge_list_path.c_str(), N_, &g_));
        check(sbmbp_graph_info(g_, nullptr, &M_, &E_, &max_degree_));
    }
    ~blockmodel_t() { sbmbp_graph_destroy(g_); }
    blockmodel_t(const blockmodel_t &) = delete;
    blockmodel_t &operator=(const blockmodel_t &) = delete;

    unsigned int get_N() const { return N_; }
    unsigned int get_Q() const { return unsigned(n_.size()); }
    uint64_t get_E() const { return E_; }
    uint64_t get_M() const { return M_; }
    unsigned int get_graph_max_degree() const { return max_degree_; }
    unsigned int get_deg_corr_flag() const { return deg_corr_flag_; }
    const sbmbp_graph *graph() const { return g_; }
    // main.cpp:239-252: memberships from the block sizes
    uint_vec_t get_memberships() const {
        uint_vec_t mb;
        mb.reserve(N_);
        for (unsigned r = 0; r < n_.size(); ++r) mb.insert(mb.end(), n_[r], r);
        return mb;
    }

private:
    uint_vec_t n_;
    unsigned int deg_corr_flag_;
    uint32_t N_ = 0, max_degree_ = 0;
    uint64_t M_ = 0, E_ = 0;
    sbmbp_graph *g_ = nullptr;
};

inline bp_blockmodel_state bp_param_from_direct(const blockmodel_t &bm, const double_vec_t &pa, const double_vec_t &cab) {
    const unsigned Q = bm.get_Q();
    if (pa.size() < Q || cab.size() < size_t(Q) * (Q + 1) / 2)
        throw error(SBMBP_ERR_ARG, "--pa needs Q values and --cab Q(Q+1)/2 values");
    bp_blockmodel_state st;
    st.na.resize(Q);
    st.cab.resize(size_t(Q) * Q);
    check(sbmbp_params_from_direct(bm.get_N(), Q, pa.data(), cab.data(), st.na.data(), st.cab.data()));
    return st;
}

inline bp_blockmodel_state bp_param_from_epsilon_c(const blockmodel_t &bm, double epsilon, double c) {
    const unsigned Q = bm.get_Q();
    bp_blockmodel_state st;
    st.na.resize(Q);
    st.cab.resize(size_t(Q) * Q);
    check(sbmbp_params_from_epsilon_c(bm.get_N(), Q, epsilon, c, st.na.data(), st.cab.data()));
    return st;
}

class belief_propagation {
public:
    belief_propagation(const blockmodel_t &bm, int precision = SBMBP_F64, int device = -1) : bm_(bm) {
        check(sbmbp_create(bm.graph(), bm.get_Q(), bm.get_deg_corr_flag(), precision, device, &e_));
    }
    ~belief_propagation() { sbmbp_destroy(e_); }
    belief_propagation(const belief_propagation &) = delete;
    belief_propagation &operator=(const belief_propagation &) = delete;

    // belief_propagation.cpp:101-215: conf is the beliefs vector (-1 = unknown), used by flags 1-3
    // The run's generator (main.cpp:236 `std::mt19937 engine(seed)`), owned by the engine: seed it, let --mb_rand's
    // shuffle draw from it (main.cpp:299-301), then init_messages and converge() draw from it in turn.
    void seed(unsigned int seed) { check(sbmbp_seed_schedule(e_, seed)); }
    void shuffle_memberships() { check(sbmbp_rng_shuffle(e_, bm_.get_N())); }
    void init_messages(unsigned int bp_messages_init_flag, const std::vector<int> &conf, const uint_vec_t &true_conf) {
        conf_true_ = true_conf;
        if (bp_messages_init_flag != 0 && conf.size() < bm_.get_N())
            throw error(SBMBP_ERR_ARG, "the beliefs vector is shorter than the number of nodes (the reference reads out of bounds here)");
        check(sbmbp_init_messages_continue(e_, bp_messages_init_flag, bp_messages_init_flag ? conf.data() : nullptr));
    }
    void init_messages(unsigned int bp_messages_init_flag, const std::vector<int> &conf, const uint_vec_t &true_conf,
                       unsigned int seed_value) {
        seed(seed_value);
        init_messages(bp_messages_init_flag, conf, true_conf);
    }
    // main.cpp:318-323: bp_conditional for -m infer (planted nodes frozen), bp_basic for -m learn
    void set_conditional(bool on) { check(sbmbp_set_conditional(e_, on ? 1 : 0)); }
    // addition: graph-coloured asynchronous sweeps instead of synchronous ones (same fixed point, fewer sweeps)
    // "replay": the reference's own random-sequential schedule, draw for draw (one warp; small graphs)
    void set_schedule(const std::string &name) {
        check(sbmbp_set_schedule(e_, name == "colored" ? SBMBP_SCHED_COLORED : name == "replay" ? SBMBP_SCHED_REPLAY : SBMBP_SCHED_SYNC));
    }
    void init_special_needs(bool if_output_marginals) { if_output_marginals_ = if_output_marginals; }
    void set_beta(double beta) { beta_ = beta; }
    void expand_bp_params(const bp_blockmodel_state &st) { check(sbmbp_set_params(e_, st.na.data(), st.cab.data(), beta_)); }

    int converge(float conv_crit, unsigned int time_conv, float dumping_rate) {
        int niter = -1;
        check(sbmbp_converge(e_, conv_crit, time_conv, dumping_rate, &niter));
        return niter;
    }
    double compute_free_energy() {
        double f = 0;
        check(sbmbp_free_energy(e_, &f, nullptr, nullptr, nullptr));
        return f;
    }
    double compute_entropy() {
        double s = 0;
        check(sbmbp_entropy(e_, &s));
        return s;
    }
    double compute_overlap() {
        double ov = 0;
        check(sbmbp_overlap(e_, conf_true_.data(), &ov));
        return ov;
    }

    // belief_propagation.cpp:77-99: "<entropy> <free_energy> <overlap> <niter> \n" [+ N lines of Q marginals]
    void inference(const bp_blockmodel_state &st, float conv_crit, unsigned int time_conv, float dumping_rate,
                   std::ostream &out = std::cout) {
        expand_bp_params(st);
        int niter = converge(conv_crit, time_conv, dumping_rate);
        double f = compute_free_energy();
        double e = compute_entropy();
        out << e << " " << f << " " << compute_overlap() << " " << niter << " \n";
        if (if_output_marginals_) {
            const unsigned Q = bm_.get_Q(), N = bm_.get_N();
            std::vector<double> marg(size_t(N) * Q);
            check(sbmbp_get_marginals(e_, marg.data()));
            for (unsigned i = 0; i < N; ++i) {
                for (unsigned q = 0; q < Q; ++q) out << marg[size_t(i) * Q + q] << " ";
                out << "\n";
            }
        }
    }

    // belief_propagation.cpp:14-51: eta line, Q lines of cab on stdout; "overlap:<x>" on clog
    void learning(const bp_blockmodel_state &st, float learning_conv_crit, unsigned int learning_max_time,
                  float learning_rate, float dumping_rate, std::ostream &out = std::cout) {
        expand_bp_params(st);
        const unsigned Q = bm_.get_Q();
        std::vector<uint32_t> na(Q);
        std::vector<double> cab(size_t(Q) * Q), eta(Q);
        check(sbmbp_learn(e_, learning_conv_crit, learning_max_time, learning_rate, dumping_rate, na.data(),
                          cab.data(), eta.data(), nullptr));
        for (unsigned q = 0; q < Q; ++q) out << eta[q] << " ";
        out << "\n";
        for (unsigned r = 0; r < Q; ++r) {
            for (unsigned s = 0; s < Q; ++s) out << cab[r * Q + s] << " ";
            out << "\n";
        }
        std::clog << "overlap:" << compute_overlap() << "\n";
    }

    // updates that met a b_l[q] < 1e-50: the reference's result there depends on stale scratch (see sbmbp_tiny_events)
    uint64_t tiny_events() {
        uint64_t n = 0;
        check(sbmbp_tiny_events(e_, &n));
        return n;
    }

    sbmbp_engine *engine() { return e_; }

private:
    const blockmodel_t &bm_;
    sbmbp_engine *e_ = nullptr;
    double beta_ = 1.0;
    bool if_output_marginals_ = false;
    uint_vec_t conf_true_;
};

}  // namespace sbmbp_host
