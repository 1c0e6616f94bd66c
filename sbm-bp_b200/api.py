"""ctypes binding of libsbmbp.so (include/sbmbp.h) with the reference's vocabulary.

Mirrors the objects src/main.cpp:277-365 of the reference wires together:

    edge_list = load_edge_list(path)                 graph_utilities.cpp:42-58
    bm  = blockmodel_t(n, edge_list, deg_corr_flag)  blockmodel.cpp:7-49 (+ edge_to_adj, graph_utilities.cpp:60-77)
    bp  = belief_propagation(bm, precision="f64")    belief_propagation.h:18-178
    bp.init_messages(seed); bp.set_beta(b)
    bp.expand_bp_params(bp_param_from_direct(bm, pa, cab))
    bp.inference(conv_crit, time_conv, dumping_rate) / bp.learning(...)

Everything numerical happens in the CUDA library; this module only marshals numpy arrays.  It raises
``SbmbpError`` when the library is missing or a call fails -- there is no Python or CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBMBP_LIB", os.path.join(_HERE, "libsbmbp.so"))  # override only while tuning build variants

F64, F32 = 0, 1
_PREC = {"f64": F64, "fp64": F64, "double": F64, F64: F64, "f32": F32, "fp32": F32, "float": F32, F32: F32}

ERR_NAMES = {1: "ARG", 2: "IO", 3: "PARSE", 4: "RANGE", 5: "CUDA", 6: "NODEVICE", 7: "STATE", 8: "UNSUPPORTED"}


class SbmbpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sbmbp error %s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


_lib = None


def lib():
    """The loaded libsbmbp.so.  Fails loudly if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SbmbpError(-1, "libsbmbp.so is not built at %s: run __graft_entry__.build(); there is no fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.sbmbp_last_error.restype = C.c_char_p
        _lib.sbmbp_version.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise SbmbpError(rc, lib().sbmbp_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# every symbol include/sbmbp.h declares (tests assert they are all exported)
SYMBOLS = [
    "sbmbp_version", "sbmbp_last_error", "sbmbp_graph_from_edgelist", "sbmbp_graph_from_pairs",
    "sbmbp_graph_destroy", "sbmbp_graph_info", "sbmbp_graph_csr", "sbmbp_parse_edgelist", "sbmbp_ell_layout", "sbmbp_debug_trace", "sbmbp_sweep_kernel_name",
    "sbmbp_params_from_direct", "sbmbp_params_from_epsilon_c", "sbmbp_create", "sbmbp_destroy",
    "sbmbp_set_stream", "sbmbp_set_params", "sbmbp_get_params", "sbmbp_init_random",
    "sbmbp_init_random_device", "sbmbp_init_messages", "sbmbp_set_conditional", "sbmbp_set_schedule", "sbmbp_seed_schedule", "sbmbp_rng_shuffle", "sbmbp_init_messages_continue",
    "sbmbp_graph_coloring", "sbmbp_set_state", "sbmbp_get_state", "sbmbp_get_marginals", "sbmbp_sweep",
    "sbmbp_sweeps_async", "sbmbp_sync", "sbmbp_time_sweep_kernel", "sbmbp_converge", "sbmbp_free_energy", "sbmbp_entropy",
    "sbmbp_overlap", "sbmbp_em_stats", "sbmbp_learn", "sbmbp_stats",
    "sbmbp_graph_from_pairs_range", "sbmbp_plan_create", "sbmbp_plan_sendlist", "sbmbp_plan_expect", "sbmbp_plan_recv",
    "sbmbp_plan_finish", "sbmbp_plan_layout", "sbmbp_plan_destroy", "sbmbp_create_dist", "sbmbp_dist_ipc_export",
    "sbmbp_dist_ipc_import", "sbmbp_dist_sync_mirror", "sbmbp_dist_field_local", "sbmbp_dist_arm",
    "sbmbp_dist_sweeps", "sbmbp_dist_close", "sbmbp_plan_exchange_tables", "sbmbp_dist_finalize", "sbmbp_dist_node_stats", "sbmbp_dist_energy_local",
    "sbmbp_dist_moment_local", "sbmbp_dist_edge_pairs_local", "sbmbp_dist_set_degrees",
    "sbmbp_tiny_events", "sbmbp_set_exact_pairs_max_n", "sbmbp_non_edge_series_order", "sbmbp_non_edge_series_term",
]


def non_edge_series_order(Q, N, beta, cab):
    """Number of terms of the non-edge moment series the engine uses for an N-node, Q-group model (host only)."""
    cab = np.ascontiguousarray(cab, np.float64)
    K = C.c_uint32()
    _check(lib().sbmbp_non_edge_series_order(C.c_uint32(Q), C.c_double(N), C.c_double(beta), _p(cab), C.byref(K)))
    return K.value


def non_edge_series_term(Q, N, beta, cab, k, T):
    """-<W1^(x)k, T (x) T> / k for the order-k moment tensor T (flat, first digit fastest); host only."""
    cab = np.ascontiguousarray(cab, np.float64)
    T = np.ascontiguousarray(T, np.float64)
    assert T.size == Q ** k
    out = C.c_double()
    _check(lib().sbmbp_non_edge_series_term(C.c_uint32(Q), C.c_double(N), C.c_double(beta), _p(cab), C.c_uint32(k), _p(T),
                                            C.byref(out)))
    return out.value


def load_edge_list(path):
    """graph_utilities.cpp:42-58.  Returns (u, v) uint32 arrays of the pairs in file order."""
    n = C.c_uint64(0)
    _check(lib().sbmbp_parse_edgelist(os.fsencode(path), None, None, C.c_uint64(0), C.byref(n)))
    u = np.zeros(max(n.value, 1), np.uint32)
    v = np.zeros(max(n.value, 1), np.uint32)
    _check(lib().sbmbp_parse_edgelist(os.fsencode(path), _p(u), _p(v), C.c_uint64(n.value), C.byref(n)))
    return u[: n.value], v[: n.value]


class bp_blockmodel_state:
    """types.h:23-26: the (na, cab) pair handed to expand_bp_params."""

    def __init__(self, na, cab):
        self.na = np.ascontiguousarray(na, np.uint32)
        self.cab = np.ascontiguousarray(cab, np.float64)


class blockmodel_t:
    """Graph + block sizes (blockmodel.h:10-108 as far as the BP path uses it).

    ``n`` is the -n block-size vector (Q = len(n), N = sum(n), main.cpp:271-276); the edge list is either a
    path or a (u, v) pair of arrays.  Builds the destination-sorted CSR with the reverse-edge index on the host.
    """

    def __init__(self, n, edge_list, deg_corr_flag=0):
        self.n = np.ascontiguousarray(n, np.uint32)
        self._Q = int(len(self.n))
        self._N = int(self.n.sum())
        self._dc = int(deg_corr_flag)
        g = C.c_void_p()
        if isinstance(edge_list, (str, bytes, os.PathLike)):
            _check(lib().sbmbp_graph_from_edgelist(os.fsencode(edge_list), C.c_uint32(self._N), C.byref(g)))
        else:
            u = np.ascontiguousarray(edge_list[0], np.uint32)
            v = np.ascontiguousarray(edge_list[1], np.uint32)
            if len(u) != len(v):
                raise ValueError("u and v differ in length")
            _check(lib().sbmbp_graph_from_pairs(_p(u), _p(v), C.c_uint64(len(u)), C.c_uint32(self._N), C.byref(g)))
        self._g = g
        N, M, E, md = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        _check(lib().sbmbp_graph_info(self._g, C.byref(N), C.byref(M), C.byref(E), C.byref(md)))
        self._M, self._E, self._maxdeg = M.value, E.value, md.value
        # main.cpp:239-252: memberships from the block sizes; also the default true configuration (:284-286)
        self.memberships = np.repeat(np.arange(self._Q, dtype=np.uint32), self.n)

    def close(self):
        if getattr(self, "_g", None):
            lib().sbmbp_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_N(self):
        return self._N

    def get_Q(self):
        return self._Q

    def get_E(self):
        return self._E

    def get_M(self):
        """directed edges (2E): one synchronous sweep performs exactly this many message updates"""
        return self._M

    def get_graph_max_degree(self):
        return self._maxdeg

    def get_deg_corr_flag(self):
        return self._dc

    def csr(self):
        """(row_ptr u64[N+1], col u32[M] == graph_neis_, rev u32[M] == row_ptr[col] + graph_neis_inv_, deg u32[N])"""
        rp, col, rev, deg = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        _check(lib().sbmbp_graph_csr(self._g, C.byref(rp), C.byref(col), C.byref(rev), C.byref(deg)))
        row_ptr = np.ctypeslib.as_array(rp, shape=(self._N + 1,)).copy()
        if self._M:
            c = np.ctypeslib.as_array(col, shape=(self._M,)).copy()
            r = np.ctypeslib.as_array(rev, shape=(self._M,)).copy()
        else:
            c = np.zeros(0, np.uint32)
            r = np.zeros(0, np.uint32)
        d = np.ctypeslib.as_array(deg, shape=(self._N,)).copy() if self._N else np.zeros(0, np.uint32)
        return row_ptr, c, r, d


def graph_coloring(blockmodel):
    """The greedy colouring behind the coloured schedule (host only): (color u8[N], number of colours)."""
    color = np.zeros(max(blockmodel._N, 1), np.uint8)
    nc = C.c_uint32(0)
    _check(lib().sbmbp_graph_coloring(blockmodel._g, _p(color), C.byref(nc)))
    return color[: blockmodel._N], nc.value


def ell_layout(blockmodel, region_slots=0):
    """Host-side view of the degree-class message layout (sbmbp_ell_layout): dict of numpy arrays.  No GPU needed."""
    g, M, N = blockmodel._g, blockmodel._M, blockmodel._N
    n_cls, n_node, n_chunks, n_buckets = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    n_idx = C.c_uint64(0)
    _check(lib().sbmbp_ell_layout(g, C.c_uint64(region_slots), None, None, None, C.c_uint32(0), C.byref(n_cls), None,
                                  C.byref(n_node), None, None, C.c_uint64(0), C.byref(n_idx), C.byref(n_chunks),
                                  C.byref(n_buckets)))
    pos, gather = np.zeros(max(M, 1), np.uint32), np.zeros(max(M, 1), np.uint32)
    classes = np.zeros((max(n_cls.value, 1), 5), np.uint32)
    node = np.zeros(max(n_node.value, 1), np.uint32)
    rev_idx, pos_idx = np.zeros(max(n_idx.value, 1), np.uint32), np.zeros(max(n_idx.value, 1), np.uint32)
    _check(lib().sbmbp_ell_layout(g, C.c_uint64(region_slots), _p(pos), _p(gather), _p(classes), C.c_uint32(n_cls.value),
                                  C.byref(n_cls), _p(node), C.byref(n_node), _p(rev_idx), _p(pos_idx),
                                  C.c_uint64(n_idx.value), C.byref(n_idx), C.byref(n_chunks), C.byref(n_buckets)))
    return {"pos": pos[:M], "gather": gather[:M], "classes": classes[: n_cls.value], "node": node[: n_node.value],
            "rev_idx": rev_idx[: n_idx.value], "pos_idx": pos_idx[: n_idx.value], "n_chunks": n_chunks.value,
            "n_buckets": n_buckets.value}


def bp_param_from_direct(blockmodel, pa, cab):
    """blockmodel.cpp:274-302: cab is the upper triangle in row-major order."""
    Q, N = blockmodel.get_Q(), blockmodel.get_N()
    pa = np.ascontiguousarray(pa, np.float64)
    cu = np.ascontiguousarray(cab, np.float64)
    if len(pa) != Q or len(cu) != Q * (Q + 1) // 2:
        raise ValueError("pa needs Q entries and cab Q(Q+1)/2")
    na = np.zeros(Q, np.uint32)
    full = np.zeros((Q, Q), np.float64)
    _check(lib().sbmbp_params_from_direct(C.c_uint32(N), C.c_uint32(Q), _p(pa), _p(cu), _p(na), _p(full)))
    return bp_blockmodel_state(na, full)


def bp_param_from_epsilon_c(blockmodel, epsilon, c):
    """blockmodel.cpp:229-272."""
    Q, N = blockmodel.get_Q(), blockmodel.get_N()
    na = np.zeros(Q, np.uint32)
    full = np.zeros((Q, Q), np.float64)
    _check(lib().sbmbp_params_from_epsilon_c(C.c_uint32(N), C.c_uint32(Q), C.c_double(epsilon), C.c_double(c), _p(na), _p(full)))
    return bp_blockmodel_state(na, full)


class belief_propagation:
    """The engine: one graph on one B200 (belief_propagation.h:18-178)."""

    def __init__(self, blockmodel, precision="f64", device=-1):
        self.bm = blockmodel
        self.N, self.Q, self.M = blockmodel.get_N(), blockmodel.get_Q(), blockmodel.get_M()
        self.precision = _PREC[precision]
        self._beta = 1.0
        self._state = None
        self.conf_true = blockmodel.memberships.copy()
        e = C.c_void_p()
        _check(lib().sbmbp_create(blockmodel._g, C.c_uint32(self.Q), C.c_uint32(blockmodel.get_deg_corr_flag()),
                                  C.c_int(self.precision), C.c_int(device), C.byref(e)))
        self._e = e

    def close(self):
        if getattr(self, "_e", None):
            lib().sbmbp_destroy(self._e)
            self._e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        """Run on this cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        _check(lib().sbmbp_set_stream(self._e, C.c_void_p(int(cuda_stream))))

    # ---- reference-named surface
    def init_messages(self, seed, bp_messages_init_flag=0, true_conf=None, conf=None):
        """belief_propagation.cpp:101-215: the same draws as std::mt19937(seed); flags 1-3 take the beliefs vector
        `conf` (-1 = unknown) that main.cpp builds from --beliefs_path / -f."""
        if true_conf is not None:
            self.conf_true = np.ascontiguousarray(true_conf, np.uint32)
        if bp_messages_init_flag != 0:
            if conf is None or len(conf) != self.N:
                raise SbmbpError(2, "bp_messages_init_flag 1-3 need a beliefs vector with one entry per node")
            cf = np.ascontiguousarray(conf, np.int32)
            _check(lib().sbmbp_init_messages(self._e, C.c_uint32(bp_messages_init_flag), _p(cf), C.c_uint32(seed)))
            return
        _check(lib().sbmbp_init_random(self._e, C.c_uint32(seed)))

    def set_schedule(self, schedule="sync"):
        """"sync" (default), "colored": graph-coloured asynchronous sweeps (SBMBP_SCHED_COLORED), or "replay": the
        reference's own random-sequential schedule draw for draw (SBMBP_SCHED_REPLAY; belief_propagation.cpp:392-405)."""
        _check(lib().sbmbp_set_schedule(self._e, C.c_int({"sync": 0, "colored": 1, "replay": 2}[schedule])))

    def shuffle_memberships(self):
        """--mb_rand (main.cpp:299-301): the draws of std::shuffle over the N memberships, on the engine's generator."""
        _check(lib().sbmbp_rng_shuffle(self._e, C.c_uint32(self.N)))

    def init_messages_continue(self, bp_messages_init_flag=0, conf=None):
        """init_messages drawing from the engine's generator where it stands (after seed_schedule / shuffle_memberships)."""
        cf = None
        if bp_messages_init_flag != 0:
            if conf is None or len(conf) != self.N:
                raise SbmbpError(2, "bp_messages_init_flag 1-3 need a beliefs vector with one entry per node")
            cf = np.ascontiguousarray(conf, np.int32)
        _check(lib().sbmbp_init_messages_continue(self._e, C.c_uint32(bp_messages_init_flag), _p(cf)))

    def seed_schedule(self, seed):
        """std::mt19937(seed) as the generator the replay schedule draws from (init_messages leaves its own behind)."""
        _check(lib().sbmbp_seed_schedule(self._e, C.c_uint32(seed)))

    def set_conditional(self, on=True):
        """bp_conditional (-m infer: planted nodes frozen, belief_propagation.cpp:1100-1126) vs bp_basic (-m learn)."""
        _check(lib().sbmbp_set_conditional(self._e, C.c_int(1 if on else 0)))

    def init_messages_device(self, seed):
        """Same distribution from a counter-based generator on the GPU (large graphs)."""
        _check(lib().sbmbp_init_random_device(self._e, C.c_uint64(seed)))

    def set_beta(self, beta):
        self._beta = float(beta)
        if self._state is not None:
            self.expand_bp_params(self._state)

    def expand_bp_params(self, state):
        """belief_propagation.cpp:290-317."""
        self._state = state
        na = np.ascontiguousarray(state.na, np.uint32)
        cab = np.ascontiguousarray(state.cab, np.float64).reshape(-1)
        _check(lib().sbmbp_set_params(self._e, _p(na), _p(cab), C.c_double(self._beta)))

    def get_params(self):
        na = np.zeros(self.Q, np.uint32)
        cab = np.zeros((self.Q, self.Q), np.float64)
        eta = np.zeros(self.Q, np.float64)
        _check(lib().sbmbp_get_params(self._e, _p(na), _p(cab), _p(eta)))
        return na, cab, eta

    def set_state(self, msg=None, marg=None):
        """msg[M,Q] in the reference order (mmap_[i][l][q]), marg[N,Q] (real_psi_)."""
        m = np.ascontiguousarray(msg, np.float64) if msg is not None else None
        g = np.ascontiguousarray(marg, np.float64) if marg is not None else None
        _check(lib().sbmbp_set_state(self._e, _p(m), _p(g)))

    def get_state(self):
        msg = np.zeros((max(self.M, 1), self.Q), np.float64)
        marg = np.zeros((max(self.N, 1), self.Q), np.float64)
        h = np.zeros(self.Q, np.float64)
        _check(lib().sbmbp_get_state(self._e, _p(msg), _p(marg), _p(h)))
        return msg[: self.M], marg[: self.N], h

    def get_marginals(self):
        marg = np.zeros((max(self.N, 1), self.Q), np.float64)
        _check(lib().sbmbp_get_marginals(self._e, _p(marg)))
        return marg[: self.N]

    def sweep(self, dumping_rate=1.0):
        """One synchronous sweep over all M directed edges; returns the max-diff of norm_m_at_i (:1059-1063)."""
        md = C.c_double(0)
        _check(lib().sbmbp_sweep(self._e, C.c_double(dumping_rate), C.byref(md)))
        return md.value

    def sweeps_async(self, n, dumping_rate=1.0):
        _check(lib().sbmbp_sweeps_async(self._e, C.c_uint32(n), C.c_double(dumping_rate)))

    def sync(self):
        _check(lib().sbmbp_sync(self._e))

    def sweep_kernel_name(self):
        """Which sweep kernel the engine launches for this graph / these parameters (reporting only)."""
        buf = C.create_string_buffer(256)
        _check(lib().sbmbp_sweep_kernel_name(self._e, buf, C.c_uint32(256)))
        return buf.value.decode()

    def time_sweep_kernel(self, dumping_rate=1.0):
        """One sweep; returns the device milliseconds of the sweep kernel alone."""
        ms = C.c_float(0)
        _check(lib().sbmbp_time_sweep_kernel(self._e, C.c_double(dumping_rate), C.byref(ms)))
        return ms.value

    def converge(self, conv_crit=5e-6, time_conv=100, dumping_rate=1.0):
        """belief_propagation.cpp:386-415: returns the sweep index at convergence or -1."""
        it = C.c_int(0)
        _check(lib().sbmbp_converge(self._e, C.c_float(conv_crit), C.c_uint32(time_conv), C.c_float(dumping_rate), C.byref(it)))
        return it.value

    def compute_free_energy(self, parts=False):
        f, fs, fe, fn = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        _check(lib().sbmbp_free_energy(self._e, C.byref(f), C.byref(fs), C.byref(fe), C.byref(fn)))
        return (f.value, fs.value, fe.value, fn.value) if parts else f.value

    def tiny_events(self):
        """Edge updates that met a b_l[q] < 1e-50, where the reference's own result is not a function of its inputs."""
        n = C.c_uint64()
        _check(lib().sbmbp_tiny_events(self._e, C.byref(n)))
        return n.value

    def set_exact_pairs_max_n(self, n):
        """Largest N whose O(N^2) non-edge terms are summed pair by pair; beyond it the moment series (0 = always)."""
        _check(lib().sbmbp_set_exact_pairs_max_n(self._e, C.c_uint32(int(n))))

    def compute_entropy(self):
        s = C.c_double()
        _check(lib().sbmbp_entropy(self._e, C.byref(s)))
        return s.value

    def compute_overlap(self):
        ov = C.c_double()
        conf = np.ascontiguousarray(self.conf_true, np.uint32)
        _check(lib().sbmbp_overlap(self._e, _p(conf), C.byref(ov)))
        return ov.value

    def em_stats(self):
        """(na_expect, nna_expect, cab_expect) of compute_na_expect / compute_cab_expect (:428-440, :892-989)."""
        na = np.zeros(self.Q, np.float64)
        nna = np.zeros(self.Q, np.float64)
        cab = np.zeros((self.Q, self.Q), np.float64)
        _check(lib().sbmbp_em_stats(self._e, _p(na), _p(nna), _p(cab)))
        return na, nna, cab

    def inference(self, state, conv_crit=5e-6, time_conv=100, dumping_rate=1.0):
        """belief_propagation.cpp:77-99.  Returns (entropy, free_energy, overlap, niter) -- the stdout line."""
        self.expand_bp_params(state)
        niter = self.converge(conv_crit, time_conv, dumping_rate)
        f = self.compute_free_energy()
        e = self.compute_entropy()
        return e, f, self.compute_overlap(), niter

    def learning(self, state, learning_conv_crit=1e-6, learning_max_time=100, learning_rate=0.2, dumping_rate=1.0):
        """belief_propagation.cpp:14-51.  Returns (eta, cab, na, em_iterations); overlap via compute_overlap()."""
        self.expand_bp_params(state)
        na = np.zeros(self.Q, np.uint32)
        cab = np.zeros((self.Q, self.Q), np.float64)
        eta = np.zeros(self.Q, np.float64)
        it = C.c_int(0)
        _check(lib().sbmbp_learn(self._e, C.c_float(learning_conv_crit), C.c_uint32(learning_max_time),
                                 C.c_float(learning_rate), C.c_float(dumping_rate), _p(na), _p(cab), _p(eta), C.byref(it)))
        self._state = bp_blockmodel_state(na, cab)
        return eta, cab, na, it.value

    def stats(self):
        eu, sw, la = C.c_uint64(), C.c_uint64(), C.c_uint64()
        bpe, sec = C.c_double(), C.c_double()
        _check(lib().sbmbp_stats(self._e, C.byref(eu), C.byref(sw), C.byref(la), C.byref(bpe), C.byref(sec)))
        return {"edge_updates": eu.value, "sweeps": sw.value, "launches": la.value, "bytes_per_edge": bpe.value,
                "sweep_seconds": sec.value}


# the reference's two concrete classes differ only for clamped nodes (bp_messages_init_flag != 0), which this
# build does not offer yet; both names resolve to the same engine
bp_basic = belief_propagation
bp_conditional = belief_propagation
