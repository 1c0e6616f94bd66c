"""TEST INFRASTRUCTURE ONLY -- ctypes front ends of the two CPU checkers.

* ``Oracle``    : oracle/libbp_oracle.so, the plain-C restatement (bp_oracle.c), kind "port".
* ``Reference`` : oracle/_ref/libsbmbp_ref.so, the UNMODIFIED reference sources behind
                  oracle/ref_harness.cpp, kind "reference".  Built here from /root/reference; on the
                  GPU box only the prebuilt file is used.

Both expose the same methods, named after the reference functions they stand for
(belief_propagation.cpp / blockmodel.cpp / graph_utilities.cpp).  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libbp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsbmbp_ref.so")
REFERENCE_ROOT = "/root/reference"

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the checkers (make -C oracle).  The reference build is attempted only where
    /root/reference exists; elsewhere a prebuilt oracle/_ref travels with the snapshot."""
    need = force or not os.path.exists(ORACLE_SO)
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")) and not os.path.exists(REF_SO):
        need = True
    if need:
        subprocess.check_call(["make", "-s", "-C", HERE, "all"], stdout=sys.stderr)  # stdout belongs to the caller (bench.py's JSON line)


def have_reference():
    return os.path.exists(REF_SO)


def _opt(a, dtype):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype).ctypes.data_as(C.c_void_p)


class _Base:
    """Shared surface; subclasses bind the symbol prefix ('orc_' or 'ref_')."""

    _lib = None
    _p = ""
    kind = ""

    def _f(self, name, restype=None):
        fn = getattr(self._lib, self._p + name)
        fn.restype = restype
        return fn

    # -- construction: main.cpp:239-296
    def __init__(self, u, v, block_sizes, dc_flag=0, learn_mode=False):
        u = np.ascontiguousarray(u, dtype=np.uint32)
        v = np.ascontiguousarray(v, dtype=np.uint32)
        bs = np.ascontiguousarray(block_sizes, dtype=np.uint32)
        self.Q = int(len(bs))
        self._h = C.c_void_p(self._create(u, v, bs, int(dc_flag), bool(learn_mode)))
        self._post_create()
        self.N = int(self._f("N", C.c_uint32)(self._h))
        self.M = int(self._f("M", C.c_uint64)(self._h))
        self.dc_flag = int(dc_flag)

    def _post_create(self):
        pass

    def close(self):
        if getattr(self, "_h", None):
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def E(self):
        return int(self._f("E", C.c_uint32)(self._h))

    @property
    def max_degree(self):
        return int(self._f("max_degree", C.c_uint32)(self._h))

    def csr(self):
        """(row_ptr u64[N+1], col u32[M], rev_local u32[M], rev_global u64[M]) == graph_neis_ / graph_neis_inv_"""
        row_ptr = np.zeros(self.N + 1, np.uint64)
        col = np.zeros(max(self.M, 1), np.uint32)
        rl = np.zeros(max(self.M, 1), np.uint32)
        rg = np.zeros(max(self.M, 1), np.uint64)
        self._f("get_csr")(self._h, row_ptr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p),
                           rl.ctypes.data_as(C.c_void_p), rg.ctypes.data_as(C.c_void_p))
        return row_ptr, col[: self.M], rl[: self.M], rg[: self.M]

    # -- parameters
    def set_params_direct(self, pa, cab_upper):
        pa = np.ascontiguousarray(pa, np.float64)
        cu = np.ascontiguousarray(cab_upper, np.float64)
        assert len(pa) == self.Q and len(cu) == self.Q * (self.Q + 1) // 2
        self._f("set_params_direct")(self._h, pa.ctypes.data_as(C.c_void_p), cu.ctypes.data_as(C.c_void_p))

    def set_params_epsilon_c(self, eps, c):
        self._f("set_params_epsilon_c")(self._h, C.c_double(eps), C.c_double(c))

    def set_params_raw(self, na, cab):
        na = np.ascontiguousarray(na, np.uint32)
        cab = np.ascontiguousarray(cab, np.float64).reshape(-1)
        self._f("set_params_raw")(self._h, na.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p))

    def get_params(self):
        na = np.zeros(self.Q, np.uint32)
        cab = np.zeros((self.Q, self.Q), np.float64)
        eta = np.zeros(self.Q, np.float64)
        self._f("get_params")(self._h, na.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p),
                              eta.ctypes.data_as(C.c_void_p))
        return na, cab, eta

    def set_beta(self, beta):
        self._f("set_beta")(self._h, C.c_double(beta))

    # -- state
    def get_state(self):
        msg = np.zeros((max(self.M, 1), self.Q), np.float64)
        marg = np.zeros((max(self.N, 1), self.Q), np.float64)
        h = np.zeros(self.Q, np.float64)
        self._f("get_state")(self._h, msg.ctypes.data_as(C.c_void_p), marg.ctypes.data_as(C.c_void_p),
                             h.ctypes.data_as(C.c_void_p))
        return msg[: self.M], marg[: self.N], h

    def set_state(self, msg, marg):
        self._f("set_state")(self._h, _opt(msg, np.float64), _opt(marg, np.float64))

    def init_h(self):
        self._f("init_h")(self._h)

    def get_conf_planted(self):
        conf = np.zeros(max(self.N, 1), np.int32)
        self._f("get_conf_planted")(self._h, conf.ctypes.data_as(C.c_void_p))
        return conf[: self.N]

    # -- sweeps
    def jacobi_sweep(self, damping=1.0):
        """One synchronous sweep by the reference arithmetic from the frozen current state.
        Returns (new_msg[M,Q] in reference slot order, new_marg[N,Q], node_diff[N], maxdiff)."""
        nm = np.zeros((max(self.M, 1), self.Q), np.float64)
        ng = np.zeros((max(self.N, 1), self.Q), np.float64)
        nd = np.zeros(max(self.N, 1), np.float64)
        md = self._f("jacobi_sweep", C.c_double)(self._h, C.c_double(damping), nm.ctypes.data_as(C.c_void_p),
                                                 ng.ctypes.data_as(C.c_void_p), nd.ctypes.data_as(C.c_void_p))
        return nm[: self.M], ng[: self.N], nd[: self.N], float(md)

    def converge(self, crit=5e-6, max_iter=100, damping=1.0):
        """The reference's random-sequential converge() (belief_propagation.cpp:386-415)."""
        return int(self._f("converge", C.c_int)(self._h, C.c_float(crit), C.c_uint32(max_iter), C.c_float(damping)))

    # -- reductions
    def free_energy(self):
        return float(self._f("free_energy", C.c_double)(self._h))

    def f_site(self):
        return float(self._f("f_site", C.c_double)(self._h))

    def f_edge(self):
        return float(self._f("f_edge", C.c_double)(self._h))

    def f_non_edge(self):
        return float(self._f("f_non_edge", C.c_double)(self._h))

    def entropy(self):
        return float(self._f("entropy", C.c_double)(self._h))

    def entropy_site(self):
        return float(self._f("entropy_site", C.c_double)(self._h))

    def entropy_edge(self):
        return float(self._f("entropy_edge", C.c_double)(self._h))

    def entropy_non_edge(self):
        return float(self._f("entropy_non_edge", C.c_double)(self._h))

    def overlap(self):
        return float(self._f("overlap", C.c_double)(self._h))

    def em_stats(self):
        na = np.zeros(self.Q, np.float64)
        nna = np.zeros(self.Q, np.float64)
        cab = np.zeros((self.Q, self.Q), np.float64)
        self._f("em_stats")(self._h, na.ctypes.data_as(C.c_void_p), nna.ctypes.data_as(C.c_void_p),
                            cab.ctypes.data_as(C.c_void_p))
        return na, nna, cab

    def learning_step(self, lr):
        self._f("learning_step")(self._h, C.c_float(lr))


class Oracle(_Base):
    """Plain-C restatement (oracle/bp_oracle.c)."""

    _p = "orc_"
    kind = "port"

    def _create(self, u, v, bs, dc_flag, learn_mode):
        if Oracle._lib is None:
            build()
            Oracle._lib = C.CDLL(ORACLE_SO)
        fn = self._f("create", C.c_void_p)
        return fn(u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_uint64(len(u)),
                  bs.ctypes.data_as(C.c_void_p), C.c_uint32(len(bs)), C.c_uint32(dc_flag))

    def init_messages(self, seed, beta=1.0):
        self._f("init_messages")(self._h, C.c_uint32(seed))
        self.set_beta(beta)

    def init_messages_mb_rand(self, seed, beta=1.0):
        """--mb_rand: libstdc++'s std::shuffle draws (restated) before init_messages."""
        self._f("init_messages_mb_rand")(self._h, C.c_uint32(seed))
        self.set_beta(beta)

    def shuffle(self, n):
        """The permutation std::shuffle applies to 0..n-1 with the generator where it stands."""
        perm = np.arange(n, dtype=np.uint32)
        self._f("shuffle")(self._h, perm.ctypes.data_as(C.c_void_p), C.c_uint64(n))
        return perm

    def init_messages_flag(self, flag, conf, seed, beta=1.0):
        """belief_propagation.cpp:101-215 with an explicit beliefs vector; returns 0, or -1 where the reference asserts."""
        conf = np.ascontiguousarray(conf, np.int32)
        assert len(conf) == self.N
        rc = int(self._f("init_messages_flag", C.c_int)(self._h, C.c_uint32(flag), conf.ctypes.data_as(C.c_void_p),
                                                        C.c_uint32(seed)))
        self.set_beta(beta)
        return rc

    def set_conditional(self, on):
        """bp_conditional (-m infer) vs bp_basic (-m learn): whether planted nodes are frozen."""
        self._f("set_conditional")(self._h, C.c_int(1 if on else 0))

    def sync_sweep(self, damping=1.0):
        return float(self._f("sync_sweep", C.c_double)(self._h, C.c_double(damping)))

    def referee_sweep(self, damping=1.0, h=None):
        """The same Jacobi sweep evaluated in long double (log domain): the yardstick for hub-node parity.
        h: the field to evaluate with (default: init_h of the current state, as the reference holds it).
        Returns (new_msg, new_marg, skipped[N]); skipped marks b < 1e-50 nodes (not modelled, left unchanged)."""
        hp = None if h is None else np.ascontiguousarray(h, np.float64)
        nm = np.zeros((max(self.M, 1), self.Q), np.float64)
        ng = np.zeros((max(self.N, 1), self.Q), np.float64)
        sk = np.zeros(max(self.N, 1), np.uint8)
        self._f("referee_sweep", C.c_int)(self._h, C.c_double(damping), None if hp is None else hp.ctypes.data_as(C.c_void_p),
                                          nm.ctypes.data_as(C.c_void_p),
                                          ng.ctypes.data_as(C.c_void_p), sk.ctypes.data_as(C.c_void_p))
        return nm[: self.M], ng[: self.N], sk[: self.N].astype(bool)

    def sync_converge(self, crit=5e-6, max_iter=100, damping=1.0):
        return int(self._f("sync_converge", C.c_int)(self._h, C.c_float(crit), C.c_uint32(max_iter),
                                                     C.c_float(damping)))

    def update_node(self, i, damping=1.0):
        return float(self._f("update_node", C.c_double)(self._h, C.c_uint32(i), C.c_double(damping)))

    def learning(self, crit=1e-6, max_time=100, lr=0.2, damping=1.0, sync=False):
        it = int(self._f("learning", C.c_int)(self._h, C.c_float(crit), C.c_uint32(max_time), C.c_float(lr),
                                              C.c_float(damping), C.c_int(1 if sync else 0)))
        na, cab, eta = self.get_params()
        return na, cab, eta, it

    def set_true_conf(self, conf):
        conf = np.ascontiguousarray(conf, np.uint32)
        self._f("set_true_conf")(self._h, conf.ctypes.data_as(C.c_void_p))

    def seed(self, seed):
        self._f("seed")(self._h, C.c_uint32(seed))

    def uniform(self):
        return float(self._f("uniform", C.c_double)(self._h))

    @staticmethod
    def load_edge_list(path, cap=1 << 24):
        if Oracle._lib is None:
            build()
            Oracle._lib = C.CDLL(ORACLE_SO)
        u = np.zeros(cap, np.uint32)
        v = np.zeros(cap, np.uint32)
        fn = Oracle._lib.orc_load_edge_list
        fn.restype = C.c_uint64
        n = int(fn(os.fsencode(path), u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_uint64(cap)))
        return u[:n].copy(), v[:n].copy()


class Reference(_Base):
    """The unmodified reference engine behind oracle/ref_harness.cpp."""

    _p = "ref_"
    kind = "reference"

    def _create(self, u, v, bs, dc_flag, learn_mode):
        if Reference._lib is None:
            build()
            Reference._lib = C.CDLL(REF_SO)
        fn = self._f("create_from_pairs", C.c_void_p)
        return fn(u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_uint64(len(u)),
                  bs.ctypes.data_as(C.c_void_p), C.c_uint32(len(bs)), C.c_uint32(dc_flag),
                  C.c_int(1 if learn_mode else 0))

    def _post_create(self):
        # bp_allocate runs inside init_messages (belief_propagation.cpp:107); sizes exist only after it
        self._f("init_messages")(self._h, C.c_uint32(0), C.c_double(1.0))

    def init_messages(self, seed, beta=1.0):
        self._f("init_messages")(self._h, C.c_uint32(seed), C.c_double(beta))

    def init_messages_flag(self, flag, conf, seed, beta=1.0):
        conf = np.ascontiguousarray(conf, np.int32)
        assert len(conf) == self.N
        return int(self._f("init_messages_flag", C.c_int)(self._h, C.c_uint32(flag), conf.ctypes.data_as(C.c_void_p),
                                                          C.c_uint32(seed), C.c_double(beta)))

    def init_messages_mb_rand(self, seed, beta=1.0):
        """--mb_rand: the memberships are shuffled with the run's engine before init_messages draws from it."""
        self._f("init_messages_mb_rand")(self._h, C.c_uint32(seed), C.c_double(beta))

    def converge_timed(self, crit=5e-6, max_iter=100, damping=1.0):
        sec = C.c_double(0)
        it = int(self._f("converge_timed", C.c_int)(self._h, C.c_float(crit), C.c_uint32(max_iter),
                                                    C.c_float(damping), C.byref(sec)))
        return it, sec.value

    def learning(self, crit=1e-6, max_time=100, lr=0.2, damping=1.0):
        na = np.zeros(self.Q, np.uint32)
        cab = np.zeros((self.Q, self.Q), np.float64)
        eta = np.zeros(self.Q, np.float64)
        self._f("learning")(self._h, C.c_float(crit), C.c_uint32(max_time), C.c_float(lr), C.c_float(damping),
                            na.ctypes.data_as(C.c_void_p), cab.ctypes.data_as(C.c_void_p),
                            eta.ctypes.data_as(C.c_void_p))
        return na, cab, eta

    def inference(self, crit=5e-6, max_time=100, damping=1.0):
        """stdout line of inference() (belief_propagation.cpp:88): 'e f overlap niter \\n'"""
        buf = C.create_string_buffer(1 << 16)
        self._f("inference", C.c_int)(self._h, C.c_float(crit), C.c_uint32(max_time), C.c_float(damping), buf,
                                      C.c_int(len(buf)))
        return buf.value.decode()

    @staticmethod
    def load_edge_list(path, cap=1 << 24):
        if Reference._lib is None:
            build()
            Reference._lib = C.CDLL(REF_SO)
        u = np.zeros(cap, np.uint32)
        v = np.zeros(cap, np.uint32)
        fn = Reference._lib.ref_load_edge_list
        fn.restype = C.c_uint64
        n = int(fn(os.fsencode(path), u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_uint64(cap)))
        return u[:n].copy(), v[:n].copy()
