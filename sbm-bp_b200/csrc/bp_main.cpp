// bin/bp -- drop-in command line of the reference's `bp` binary (src/main.cpp:33-366) on the B200 engine.
//
// Same options (long and short names, multitoken vectors, defaults), same validation messages and exit
// codes, same stdout: infer -> "<entropy> <free_energy> <overlap> <niter> \n" [+ marginals]; learn -> the eta
// line and Q lines of c_ab, "overlap:<x>" on stderr.  Boost.program_options is replaced by a small table-driven
// parser.  Additions (not in the reference): --precision f64|f32, --device <k>.
// Memberships (main.cpp:176-193, :239-269, :299-301): -n ordering by default, --mb the vector itself, --mb_rand the
// shuffle's draws on the run's generator; --mb_path is a TODO stub in the reference (memberships stay empty there and the
// overlap reads out of bounds) and is refused here with a message.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "bp_host.hpp"

using namespace sbmbp_host;

namespace {

struct opt_spec {
    const char *name;
    char short_name;  // 0 = none
    int kind;         // 0 flag, 1 single value, 2 multitoken
    const char *help;
};

const opt_spec kOptions[] = {
    {"edge_list_path", 'l', 1, "Path to the input edgelist file."},
    {"n", 'n', 2, "Block sizes vector."},
    {"beta", 'b', 1, "beta, the inverse temperature (default 1)"},
    {"mb_rand", 0, 0, "Randomize initial block memberships."},
    {"mb_n", 0, 0, "Initialize membership from n [DEFAULT]."},
    {"mb", 0, 2, "Directly initialize membership from input vector."},
    {"mb_path", 0, 1, "use an external file to define the memberships."},
    {"epsilon_c", 0, 2, "Assign epsilon and c to define cab and pa [DEFAULT]."},
    {"bp_messages_init_flag", 'i', 1, "flag to initialize BP. Valid values are 0, 1, 2, and 3. Default 0."},
    {"beliefs_path", 0, 1, "Path to planted membership."},
    {"true_conf_path", 0, 1, "Path to true membership."},
    {"deg_corr_flag", 0, 1, "0 no degree correction, 1 degree correction, 2 another version of it. Default 0."},
    {"learning_rate", 'r', 1, "learning_rate, 0.0 to 1.0. Default 0.2."},
    {"dumping_rate", 'R', 1, "dumping_rate, 0.0 to 1.0. Default 1.0 (no dumping)."},
    {"bp_conv_crit", 'e', 1, "convergence criterium of BP. Default 5.0e-6."},
    {"learning_conv_crit", 'E', 1, "convergence criterium of learning. Default 1.0e-6."},
    {"time_conv", 't', 1, "maximum time for BP to converge, default 100."},
    {"probabilities", 'P', 2, "(accepted, unused -- as in the reference)"},
    {"fixed_nodes", 'f', 2, "Fixed nodes with known labels."},
    {"cab_rand", 0, 0, "(accepted, unused)"},
    {"cab_ppm", 0, 0, "(accepted, unused)"},
    {"cab_ec", 0, 0, "(accepted, unused)"},
    {"cab_file", 0, 1, "(accepted, unused)"},
    {"pa", 0, 2, "pa vector."},
    {"cab", 0, 2, "cab vector (upper triangle, row-major)."},
    {"if_output_marginals", 0, 0, "whether output marginals in the infer mode"},
    {"mode", 'm', 1, "Mode for the algorithm; valid values: infer | learn."},
    {"seed", 'd', 1, "Seed of the pseudo random number generator (Mersenne-twister 19937)."},
    {"precision", 0, 1, "[B200 engine] message storage: f64 (default) or f32."},
    {"schedule", 0, 1, "Update schedule: sync (default), colored (graph-coloured asynchronous sweeps) or replay (the reference's own random-sequential schedule, draw for draw; small graphs)."},
    {"device", 0, 1, "[B200 engine] CUDA device index (default: current)."},
    {"help", 'h', 0, "Produce this help message."},
};

const opt_spec *find_long(const std::string &n) {
    for (const auto &o : kOptions)
        if (n == o.name) return &o;
    return nullptr;
}
const opt_spec *find_short(char c) {
    for (const auto &o : kOptions)
        if (o.short_name && o.short_name == c) return &o;
    return nullptr;
}

bool looks_like_option(const char *s) {
    if (s[0] != '-' || s[1] == 0) return false;
    // "-1", "-.5", "-1e-3" are values
    if ((s[1] >= '0' && s[1] <= '9') || s[1] == '.') return false;
    return true;
}

using var_map_t = std::map<std::string, std::vector<std::string>>;

bool parse_args(int argc, char const *argv[], var_map_t &vm) {
    for (int i = 1; i < argc; ++i) {
        std::string tok = argv[i];
        const opt_spec *spec = nullptr;
        std::string inline_value;
        bool has_inline = false;
        if (tok.rfind("--", 0) == 0) {
            std::string name = tok.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                inline_value = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_inline = true;
            }
            spec = find_long(name);
        } else if (tok.size() >= 2 && tok[0] == '-') {
            spec = find_short(tok[1]);
            if (tok.size() > 2) {
                inline_value = tok.substr(2);
                has_inline = true;
            }
        }
        if (!spec) {
            std::clog << "unrecognised option '" << tok << "'\n";
            return false;
        }
        auto &vals = vm[spec->name];
        if (spec->kind == 0) {
            vals.push_back("");
            continue;
        }
        if (has_inline) vals.push_back(inline_value);
        if (spec->kind == 1) {
            if (!has_inline) {
                if (i + 1 >= argc) {
                    std::clog << "the required argument for option '--" << spec->name << "' is missing\n";
                    return false;
                }
                vals.push_back(argv[++i]);
            }
        } else {
            while (i + 1 < argc && !looks_like_option(argv[i + 1])) vals.push_back(argv[++i]);
            if (vals.empty()) {
                std::clog << "the required argument for option '--" << spec->name << "' is missing\n";
                return false;
            }
        }
    }
    return true;
}

template <typename T>
bool to_value(const std::string &s, T &out) {
    std::istringstream is(s);
    is >> out;
    return !is.fail() && is.eof();
}

template <typename T>
bool get_vec(const var_map_t &vm, const char *name, std::vector<T> &out) {
    auto it = vm.find(name);
    if (it == vm.end()) return true;
    for (const auto &s : it->second) {
        T v;
        if (!to_value(s, v)) {
            std::clog << "the argument ('" << s << "') for option '--" << name << "' is invalid\n";
            return false;
        }
        out.push_back(v);
    }
    return true;
}

template <typename T>
bool get_one(const var_map_t &vm, const char *name, T &out) {
    auto it = vm.find(name);
    if (it == vm.end() || it->second.empty()) return true;
    if (!to_value(it->second.back(), out)) {
        std::clog << "the argument ('" << it->second.back() << "') for option '--" << name << "' is invalid\n";
        return false;
    }
    return true;
}

size_t count(const var_map_t &vm, const char *name) { return vm.count(name); }

// graph_utilities.cpp:25-40
bool load_confs(uint_vec_t &conf, const std::string &path) {
    conf.clear();
    std::ifstream f(path.c_str());
    if (!f.is_open()) return false;
    std::string line;
    unsigned int membership = 0;
    while (getline(f, line)) {
        std::stringstream ls(line);
        ls >> membership;
        conf.push_back(membership);
    }
    return true;
}

// graph_utilities.cpp:9-23: one int per line; a line that does not parse pushes 0 (operator>> zeroes the target)
bool load_beliefs(std::vector<int> &beliefs, const std::string &path) {
    beliefs.clear();
    std::ifstream f(path.c_str());
    if (!f.is_open()) return false;
    std::string line;
    int membership = 0;
    while (getline(f, line)) {
        std::stringstream ls(line);
        ls >> membership;
        beliefs.push_back(membership);
    }
    return true;
}

void print_help(const char *argv0) {
    std::clog << "BP algorithms for the SBM (final output only)\n";
    std::clog << "Usage:\n  " << argv0 << " [--option_1=value] [--option_s2=value] ...\n";
    std::clog << "Options:\n";
    for (const auto &o : kOptions) {
        std::string left = "  ";
        if (o.short_name) left += std::string("-") + o.short_name + " [ --" + o.name + " ]";
        else left += std::string("--") + o.name;
        if (o.kind) left += " arg";
        std::clog << left << (left.size() < 36 ? std::string(36 - left.size(), ' ') : " ") << o.help << "\n";
    }
}

}  // namespace

int main(int argc, char const *argv[]) {
    var_map_t var_map;
    if (!parse_args(argc, argv, var_map)) return 1;

    if (count(var_map, "help") > 0 || argc == 1) {
        print_help(argv[0]);
        return 0;
    }
    // ---- validation, message for message as main.cpp:162-233
    if (count(var_map, "edge_list_path") == 0) {
        std::clog << "edge_list_path is required (-e flag)\n";
        return 1;
    }
    if (count(var_map, "mode") == 0) {
        std::clog << "mode is required (-m flag)\n";
        return 1;
    }
    if (count(var_map, "n") == 0) {
        std::clog << "n is required (-n flag)\n";
        return 1;
    }
    if (count(var_map, "mb_n") + count(var_map, "mb") + count(var_map, "mb_path") > 1) {
        std::clog << "Error! Please just select one option to assign the membership vector.\n";
        return 1;
    }
    if (count(var_map, "epsilon_c") + (count(var_map, "pa") * count(var_map, "cab")) > 1) {
        std::clog << "Error! Please just choose one way to initialize the pa/cab parameter.\n";
        return 1;
    } else if (count(var_map, "epsilon_c") == 0 && (count(var_map, "pa") + count(var_map, "cab")) < 2) {
        std::clog << "Error! Please just input both pa/cab parameters.\n";
        return 1;
    }
    const bool cab_ec = count(var_map, "epsilon_c") > 0;

    std::string edge_list_path, mode, true_conf_path, precision = "f64";
    uint_vec_t n;
    double_vec_t epsilon_c, pa, cab;
    double beta = 1.;
    unsigned int bp_messages_init_flag = 0, deg_corr_flag = 0, time_conv = 100, seed = 0;
    float learning_rate = 0.2f, dumping_rate = 1.0f, bp_conv_crit = 5.0e-6f, learning_conv_crit = 1.0e-6f;
    int device = -1;
    if (!get_one(var_map, "edge_list_path", edge_list_path) || !get_one(var_map, "mode", mode) ||
        !get_vec(var_map, "n", n) || !get_vec(var_map, "epsilon_c", epsilon_c) || !get_vec(var_map, "pa", pa) ||
        !get_vec(var_map, "cab", cab) || !get_one(var_map, "beta", beta) ||
        !get_one(var_map, "bp_messages_init_flag", bp_messages_init_flag) ||
        !get_one(var_map, "deg_corr_flag", deg_corr_flag) || !get_one(var_map, "time_conv", time_conv) ||
        !get_one(var_map, "learning_rate", learning_rate) || !get_one(var_map, "dumping_rate", dumping_rate) ||
        !get_one(var_map, "bp_conv_crit", bp_conv_crit) || !get_one(var_map, "learning_conv_crit", learning_conv_crit) ||
        !get_one(var_map, "seed", seed) || !get_one(var_map, "true_conf_path", true_conf_path) ||
        !get_one(var_map, "precision", precision) || !get_one(var_map, "device", device))
        return 1;

    if (bp_messages_init_flag != 0 && count(var_map, "fixed_nodes") == 0) {
        if (count(var_map, "beliefs_path") == 0) {
            std::clog << "Error! Please assign the file path of the initial belief of node membership.\n";
            return 1;
        }
    } else if (count(var_map, "fixed_nodes") > 0) {
        std::clog << "Randomly assign initial messages, except certain fixed nodes.\n";
    } else {
        std::clog << "Randomly assign initial messages!\n";
    }
    if (bp_messages_init_flag > 3) {
        std::clog << "Error! bp_messages_init_flag must be 0, 1, 2 or 3.\n";  // the reference asserts (:106)
        return 1;
    }
    std::string beliefs_path, schedule = "sync";
    if (!get_one(var_map, "schedule", schedule)) return 1;
    if (schedule != "sync" && schedule != "colored" && schedule != "replay") {
        std::clog << "Error! --schedule must be sync, colored or replay.\n";
        return 1;
    }
    uint_vec_t fixed_nodes, mb;
    if (!get_one(var_map, "beliefs_path", beliefs_path) || !get_vec(var_map, "fixed_nodes", fixed_nodes) ||
        !get_vec(var_map, "mb", mb))
        return 1;
    if (cab_ec && epsilon_c.size() < 2) {
        std::clog << "Error! epsilon_c needs two values: epsilon and c.\n";
        return 1;
    }
    if (count(var_map, "seed") == 0)
        seed = (unsigned int)std::chrono::high_resolution_clock::now().time_since_epoch().count();
    if (precision != "f64" && precision != "f32") {
        std::clog << "Error! --precision must be f64 or f32.\n";
        return 1;
    }

    try {
        // ---- objects, in the order of main.cpp:236-353
        blockmodel_t blockmodel(n, edge_list_path, deg_corr_flag);
        // main.cpp:176-193 / :239-269: where the memberships come from
        std::string memberships_status;
        if (count(var_map, "mb_n") + count(var_map, "mb") + count(var_map, "mb_path") == 0) memberships_status = "from_n";
        const bool memberships_randomize = count(var_map, "mb_rand") > 0;
        if (memberships_randomize) {
        } else if (count(var_map, "mb_n") > 0) {
            memberships_status = "from_n";
        } else if (count(var_map, "mb") > 0) {
            memberships_status = "direct";
        } else if (count(var_map, "mb_path") > 0) {
            memberships_status = "from_file";
        }
        uint_vec_t memberships_init;
        if (memberships_status == "from_n") {
            memberships_init = blockmodel.get_memberships();
        } else if (memberships_status == "direct") {
            if (mb.size() != blockmodel.get_N()) {
                std::clog << "Error! Size of assigned membership vector does not fit the number of nodes assigned by n.\n";
                return 1;
            }
            memberships_init = mb;
        } else {
            std::clog << "Error! --mb_path (and --mb_rand combined with --mb / --mb_path) leaves the memberships empty "
                         "in the reference (main.cpp:267-269 is a TODO); not supported.\n";
            return 1;
        }
        uint_vec_t true_conf;
        if (count(var_map, "true_conf_path") == 0) {
            std::clog << "Warning! Assign true conf using ordered node membership.\n";
            true_conf = memberships_init;
        } else if (!load_confs(true_conf, true_conf_path) || true_conf.size() < blockmodel.get_N()) {
            std::clog << "Warning! Reading true_conf_path error. Assign true conf using ordered node membership.\n";
            true_conf = memberships_init;
        }
        belief_propagation algorithm(blockmodel, precision == "f64" ? SBMBP_F64 : SBMBP_F32, device);
        algorithm.set_conditional(mode != "learn");  // main.cpp:318-323
        algorithm.set_schedule(schedule);
        // main.cpp:325-336: the beliefs file is read whether or not it exists; -f overrides entries with the true labels
        std::vector<int> beliefs;
        load_beliefs(beliefs, beliefs_path);
        if (count(var_map, "fixed_nodes") > 0) {
            beliefs.resize(true_conf.size(), -1);
            for (auto vtx : fixed_nodes) {
                if (vtx >= beliefs.size()) {
                    std::clog << "Error! fixed node id out of range.\n";
                    return 1;
                }
                beliefs[vtx] = int(true_conf[vtx]);
            }
        }
        // one generator for the whole run (main.cpp:236): --mb_rand's shuffle draws from it first (:299-301)
        algorithm.seed(seed);
        if (memberships_randomize) algorithm.shuffle_memberships();
        algorithm.init_messages(bp_messages_init_flag, beliefs, true_conf);
        algorithm.init_special_needs(count(var_map, "if_output_marginals") > 0);
        algorithm.set_beta(beta);

        bp_blockmodel_state state;
        if (cab_ec) state = bp_param_from_epsilon_c(blockmodel, epsilon_c[0], epsilon_c[1]);
        else state = bp_param_from_direct(blockmodel, pa, cab);

        if (mode == "infer") {
            algorithm.inference(state, bp_conv_crit, time_conv, dumping_rate);
        } else if (mode == "learn") {
            algorithm.learning(state, learning_conv_crit, time_conv, learning_rate, dumping_rate);
        }  // any other mode: silently nothing, exit 0 (main.cpp:361-366)
        if (const uint64_t tiny = algorithm.tiny_events())
            std::clog << "warning: " << tiny << " message updates met a term below 1e-50; the reference's result for such "
                         "states depends on stale scratch memory (belief_propagation.cpp:1013-1042): parity is not claimed\n";
    } catch (const error &err) {
        std::clog << "Error! " << err.what() << "\n";
        return 1;
    }
    return 0;
}
