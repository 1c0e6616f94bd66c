"""Multi-GPU belief propagation: one process per GPU over torch.distributed.

Rank p owns the nodes [range_starts[p], range_starts[p+1]) -- their rows, marginals and the buffers holding every
message INTO them.  libsbmbp's DIST sweep kernel collects its remote out-messages in an outbox and ships every
completed super-tile to the owners' buffers with TMA bulk copies over CUDA-IPC mappings (NVLink) while the other CTAs
keep computing; between sweeps the ranks meet on device-side flags and exchange their rows of Q+1 doubles (field
partials, max-diff) through IPC-mapped sync blocks (csrc/dist_exchange.cuh).  torch.distributed only carries the
plumbing: the one-off exchange of buffer positions and IPC handles, init_h, and the reductions of the free energy and
the EM statistics.

The host-side plan (``DistPlan``) needs no GPU and is what the world_size-2 gloo tests exercise.
"""
import ctypes as C
import itertools

import numpy as np

from . import api
from .api import _check, _p, lib


class DistPlan:
    """Rank-local graph + buffer layout + the position lists the producers need (host only)."""

    def __init__(self, u, v, N_global, range_starts, rank, world, Q, precision="f64"):
        self.rank, self.world, self.Q, self.N_global = int(rank), int(world), int(Q), int(N_global)
        self.starts = np.ascontiguousarray(range_starts, np.uint32)
        self.precision = api._PREC[precision]
        u = np.ascontiguousarray(u, np.uint32)
        v = np.ascontiguousarray(v, np.uint32)
        g = C.c_void_p()
        _check(lib().sbmbp_graph_from_pairs_range(_p(u), _p(v), C.c_uint64(len(u)), C.c_uint32(N_global),
                                                  C.c_uint32(int(self.starts[rank])), C.c_uint32(int(self.starts[rank + 1])),
                                                  C.byref(g)))
        self._g = g
        N, M = C.c_uint32(), C.c_uint64()
        _check(lib().sbmbp_graph_info(self._g, C.byref(N), C.byref(M), None, None))
        self.N_local, self.M_local = N.value, M.value
        p = C.c_void_p()
        _check(lib().sbmbp_plan_create(self._g, C.c_uint32(Q), C.c_int(self.precision), C.c_int(rank), C.c_int(world),
                                       _p(self.starts), C.byref(p)))
        self._p = p

    def sendlist(self, peer):
        data, n = C.POINTER(C.c_uint32)(), C.c_uint64()
        _check(lib().sbmbp_plan_sendlist(self._p, C.c_int(peer), C.byref(data), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, np.uint32)
        return np.ctypeslib.as_array(data, shape=(n.value,)).copy()

    def expect(self, peer):
        n = C.c_uint64()
        _check(lib().sbmbp_plan_expect(self._p, C.c_int(peer), C.byref(n)))
        return n.value

    def recv(self, peer, values):
        values = np.ascontiguousarray(values, np.uint32)
        _check(lib().sbmbp_plan_recv(self._p, C.c_int(peer), _p(values), C.c_uint64(len(values))))

    def finish(self):
        _check(lib().sbmbp_plan_finish(self._p))

    def layout(self):
        """(gather[M], pos[M] tile-sorted, info[M], pos_slot[M]) as numpy copies."""
        ga, po, inf, ps = (C.POINTER(C.c_uint32)() for _ in range(4))
        M, nt = C.c_uint64(), C.c_uint32()
        _check(lib().sbmbp_plan_layout(self._p, C.byref(ga), C.byref(po), C.byref(inf), C.byref(ps), C.byref(M), C.byref(nt)))
        m = M.value
        if m == 0:
            z = np.zeros(0, np.uint32)
            return z, z, z, z
        return tuple(np.ctypeslib.as_array(x, shape=(m,)).copy() for x in (ga, po, inf, ps))

    def exchange_tables(self):
        """Halo-exchange tables of a finished plan: dict(rpos[M], tps, nsuper, out_start, ship_start, ship[n,4], n_remote)."""
        rp, os_, ss, sh = (C.POINTER(C.c_uint32)() for _ in range(4))
        tps, ns, nsh, nrem = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(lib().sbmbp_plan_exchange_tables(self._p, C.byref(rp), C.byref(tps), C.byref(ns), C.byref(os_), C.byref(ss),
                                                C.byref(sh), C.byref(nsh), C.byref(nrem)))
        arr = lambda ptr, n, shape=None: (np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, np.uint32))
        ship = arr(sh, 4 * nsh.value).reshape(-1, 4)
        return dict(rpos=arr(rp, self.M_local), tps=tps.value, nsuper=ns.value, out_start=arr(os_, ns.value + 1),
                    ship_start=arr(ss, ns.value + 1), ship=ship, n_remote=nrem.value)

    def csr(self):
        rp, col = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)()
        _check(lib().sbmbp_graph_csr(self._g, C.byref(rp), C.byref(col), None, None))
        row_ptr = np.ctypeslib.as_array(rp, shape=(self.N_local + 1,)).copy()
        c = np.ctypeslib.as_array(col, shape=(self.M_local,)).copy() if self.M_local else np.zeros(0, np.uint32)
        return row_ptr, c

    def exchange(self, group=None):
        """Send every producer rank the positions of its messages (torch.distributed), then finish the plan."""
        import torch
        import torch.distributed as dist

        world, rank = self.world, self.rank
        send = [self.sendlist(k) for k in range(world)]
        if world == 1:
            self.recv(0, send[0])
        elif dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            ins = [torch.from_numpy(s.astype(np.int32)).to(dev) for s in send]
            outs = [torch.empty(self.expect(k), dtype=torch.int32, device=dev) for k in range(world)]
            dist.all_to_all(outs, ins, group=group)
            for k in range(world):
                self.recv(k, outs[k].cpu().numpy().astype(np.uint32))
        else:  # gloo: small test graphs, pickled lists are fine
            gathered = [None] * world
            dist.all_gather_object(gathered, send, group=group)
            for k in range(world):
                self.recv(k, gathered[k][rank])
        self.finish()

    def close(self):
        if getattr(self, "_p", None):
            lib().sbmbp_plan_destroy(self._p)
            self._p = None
        if getattr(self, "_g", None):
            lib().sbmbp_graph_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DevRow:
    """numpy-style view of a device row for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class distributed_belief_propagation:
    """belief_propagation across the ranks of a torch.distributed NCCL group (one GPU each)."""

    def __init__(self, plan, deg_corr_flag=0, device=None, group=None):
        import torch
        import torch.distributed as dist

        self.plan, self.group = plan, group
        self.rank, self.world, self.Q = plan.rank, plan.world, plan.Q
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._torch, self._dist = torch, dist
        if not plan_finished(plan):
            plan.exchange(group)
        e = C.c_void_p()
        _check(lib().sbmbp_create_dist(plan._p, C.c_uint32(deg_corr_flag), C.c_int(self.device), C.byref(e)))
        self._e = e
        _check(lib().sbmbp_set_stream(self._e, C.c_void_p(int(torch.cuda.current_stream().cuda_stream))))
        # CUDA IPC: everybody maps everybody's two message buffers
        mine = (C.c_ubyte * 192)()  # two message buffers + the sync block (flags, rows)
        _check(lib().sbmbp_dist_ipc_export(self._e, mine))
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(mine), group=group)
        for k in range(self.world):
            if k != self.rank:
                buf = (C.c_ubyte * 192).from_buffer_copy(handles[k])
                _check(lib().sbmbp_dist_ipc_import(self._e, C.c_int(k), buf))
        self._ncols = self.Q + 1
        dev = torch.device("cuda", self.device)
        self._gathered = torch.zeros(self.world * self._ncols, dtype=torch.float64, device=dev)
        self._row = None
        self.M_local, self.N_local = plan.M_local, plan.N_local
        self._beta = 1.0
        self.dc = int(deg_corr_flag)
        if self.dc != 0:
            self._share_degrees()

    def _share_degrees(self):
        """deg_corr_flag != 0: every rank learns the degrees of all nodes (one all-gather of N/P u32 per rank, once), so
        that the free-energy / EM edge pass has d_l of remote neighbours (belief_propagation.cpp:464, :584)."""
        full = gather_degrees(self.plan, self.group, self.device)
        _check(lib().sbmbp_dist_set_degrees(self._e, _p(full)))

    # ---- plumbing
    def _barrier(self):
        self._torch.cuda.synchronize()
        if self.world > 1:
            self._dist.barrier(group=self.group)

    def _row_tensor(self, ptr, ncols):
        if self._row is None or self._row[0] != ptr:
            t = self._torch.as_tensor(_DevRow(ptr, ncols), device=self._torch.device("cuda", self.device))
            self._row = (ptr, t)
        return self._row[1]

    def _allgather(self, ptr, ncols):
        if self._gathered.numel() != self.world * ncols:  # padded Q: the row is as wide as the compiled width + 1
            self._gathered = self._torch.zeros(self.world * ncols, dtype=self._torch.float64, device=self._gathered.device)
        row = self._row_tensor(ptr, ncols)
        if self.world > 1:
            self._dist.all_gather_into_tensor(self._gathered, row, group=self.group)
        else:
            self._gathered.copy_(row)
        return self._gathered

    # ---- reference-named surface
    def expand_bp_params(self, state):
        na = np.ascontiguousarray(state.na, np.uint32)
        cab = np.ascontiguousarray(state.cab, np.float64).reshape(-1)
        self._cab = cab.copy()
        _check(lib().sbmbp_set_params(self._e, _p(na), _p(cab), C.c_double(self._beta)))

    def set_beta(self, beta):
        """set_beta (belief_propagation.cpp:417-419); takes effect at the next expand_bp_params."""
        self._beta = float(beta)

    def init_messages_device(self, seed):
        _check(lib().sbmbp_init_random_device(self._e, C.c_uint64(seed * 1000003 + self.rank)))
        self._barrier()
        _check(lib().sbmbp_dist_sync_mirror(self._e))
        self._barrier()

    def set_state(self, msg=None, marg=None):
        """Rank-local state in the reference order of this rank's rows."""
        m = np.ascontiguousarray(msg, np.float64) if msg is not None else None
        g = np.ascontiguousarray(marg, np.float64) if marg is not None else None
        _check(lib().sbmbp_set_state(self._e, _p(m), _p(g)))
        self._barrier()
        _check(lib().sbmbp_dist_sync_mirror(self._e))
        self._barrier()

    def get_state(self):
        msg = np.zeros((max(self.M_local, 1), self.Q), np.float64)
        marg = np.zeros((max(self.N_local, 1), self.Q), np.float64)
        _check(lib().sbmbp_get_state(self._e, _p(msg), _p(marg), None))
        return msg[: self.M_local], marg[: self.N_local]

    def get_marginals(self):
        marg = np.zeros((max(self.N_local, 1), self.Q), np.float64)
        _check(lib().sbmbp_get_marginals(self._e, _p(marg)))
        return marg[: self.N_local]

    def init_h(self):
        """init_h (belief_propagation.cpp:320-332) over all ranks."""
        ptr, nc = C.c_void_p(), C.c_uint32()
        _check(lib().sbmbp_dist_field_local(self._e, C.byref(ptr), C.byref(nc)))
        g = self._allgather(ptr.value, nc.value)
        _check(lib().sbmbp_dist_finalize(self._e, C.c_void_p(g.data_ptr()), C.c_int(0), C.c_int(0), None, None, None))

    def _close(self, sync=True):
        md, conv, it = C.c_double(0), C.c_int(0), C.c_int(-1)
        _check(lib().sbmbp_dist_close(self._e, C.c_int(1 if sync else 0), C.byref(md), C.byref(conv), C.byref(it)))
        return md.value, conv.value, it.value

    def sweep(self, dumping_rate=1.0):
        """One synchronous sweep over the edges of all ranks; returns the global max-diff."""
        _check(lib().sbmbp_dist_arm(self._e, C.c_float(-1.0), C.c_uint32(1)))
        _check(lib().sbmbp_dist_sweeps(self._e, C.c_uint32(1), C.c_double(dumping_rate)))
        return self._close(True)[0]

    def sweeps_async(self, n, dumping_rate=1.0):
        """n sweeps back to back: no host synchronisation, no collective -- the ranks meet on device-side flags."""
        _check(lib().sbmbp_dist_arm(self._e, C.c_float(-1.0), C.c_uint32(n)))
        _check(lib().sbmbp_dist_sweeps(self._e, C.c_uint32(n), C.c_double(dumping_rate)))
        self._close(False)

    def converge(self, conv_crit=5e-6, time_conv=100, dumping_rate=1.0, check_every=None):
        """converge() (belief_propagation.cpp:386-415) over all ranks; every rank takes the same decision.  Sweeps are
        launched in batches (4, 8, ... 32); every kernel closes its predecessor on the device and turns into a no-op
        once a sweep has converged, so the host synchronises once per batch."""
        self.init_h()
        _check(lib().sbmbp_dist_arm(self._e, C.c_float(conv_crit), C.c_uint32(time_conv)))
        launched, batch = 0, 4
        while launched < time_conv:
            cur = min(batch, time_conv - launched)
            _check(lib().sbmbp_dist_sweeps(self._e, C.c_uint32(cur), C.c_double(dumping_rate)))
            launched += cur
            md, conv, it = self._close(True)
            if conv:
                return it
            batch = min(batch * 2, 32)
        return -1

    def compute_overlap(self, true_conf_local):
        """compute_overlap (belief_propagation.cpp:775-811) over all ranks."""
        conf = np.ascontiguousarray(true_conf_local, np.uint32)
        ncols = 2 * 32 + 32 * 32
        row = np.zeros(ncols, np.float64)
        nc = C.c_uint32()
        _check(lib().sbmbp_dist_node_stats(self._e, _p(conf), _p(row), C.byref(nc)))
        t = self._torch.from_numpy(row).to(self._torch.device("cuda", self.device))
        if self.world > 1:
            self._dist.all_reduce(t, group=self.group)
        row = t.cpu().numpy()
        Q = self.Q
        confm = row[64:].reshape(32, 32)[:Q, :Q]
        best = -1.0
        perms = itertools.permutations(range(Q)) if Q <= 8 else [tuple(range(Q))]
        for p in perms:
            best = max(best, sum(confm[t_, p[t_]] for t_ in range(Q)) / self.plan.N_global)
        return best

    # ---- free energy and EM over all ranks: local shares from the C ABI, all-reduced here
    def _allreduce(self, arr):
        t = self._torch.from_numpy(np.ascontiguousarray(arr, np.float64)).to(self._torch.device("cuda", self.device))
        if self.world > 1:
            self._dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _energy(self, which):
        qt = self.plan.qt if hasattr(self.plan, "qt") else self.Q
        cap = 4 + 32 * 32
        row = np.zeros(cap, np.float64)
        nc = C.c_uint32()
        _check(lib().sbmbp_dist_energy_local(self._e, C.c_int(which), _p(row), C.c_uint32(cap), C.byref(nc)))
        return self._allreduce(row[: nc.value])

    def _gather_marginals(self):
        """All ranks' marginals as one device tensor [N_global, Q] (node ranges are contiguous and ordered by rank)."""
        torch = self._torch
        dev = torch.device("cuda", self.device)
        local = torch.from_numpy(self.get_marginals()).to(dev)
        if self.world == 1:
            return local.contiguous()
        starts = [int(x) for x in self.plan.starts]
        sizes = [starts[r + 1] - starts[r] for r in range(self.world)]
        pad = max(sizes)
        buf = torch.zeros((pad, self.Q), dtype=torch.float64, device=dev)
        buf[: sizes[self.rank]] = local
        out = torch.empty((self.world, pad, self.Q), dtype=torch.float64, device=dev)
        self._dist.all_gather_into_tensor(out, buf, group=self.group)
        return torch.cat([out[r, : sizes[r]] for r in range(self.world)], dim=0).contiguous()

    def compute_free_energy(self, parts=False):
        """compute_free_energy (belief_propagation.cpp:744-750) over all ranks.  The O(N^2) non-edge term is the moment
        series of the single-GPU engine: sum_k <W1^(x)k, T_k (x) T_k> / k with the moment tensors
        all-reduced, minus the exact sum over the edges (each rank its rows, marginals all-gathered)."""
        N = float(self.plan.N_global)
        r = self._energy(0)
        fs, fe = r[0] / N, r[1] / (2.0 * N)
        if self.dc != 0:  # the dc branches of compute_f_non_edge add nothing (:692-697)
            f = -fs + fe
            return (f, fs, fe, 0.0) if parts else f
        Q = self.Q
        cab = np.asarray(self._cab, np.float64).reshape(Q, Q)
        from .api import non_edge_series_order, non_edge_series_term

        total = 0.0
        for k in range(0, non_edge_series_order(Q, N, self._beta, cab) + 1):  # the engine's own series arithmetic (k = 0: normalisation defect)
            T = np.zeros(Q ** k, np.float64)
            _check(lib().sbmbp_dist_moment_local(self._e, C.c_uint32(k), _p(T), C.c_uint64(T.size)))
            total += non_edge_series_term(Q, N, self._beta, cab, k, self._allreduce(T))
        g = self._gather_marginals()
        edges = C.c_double(0)
        _check(lib().sbmbp_dist_edge_pairs_local(self._e, C.c_void_p(g.data_ptr()), C.c_int(1), C.byref(edges)))
        self._torch.cuda.synchronize()
        edges = float(self._allreduce(np.array([edges.value]))[0])
        fn = (total - edges) / (2.0 * N)
        f = -fs + fe + fn
        return (f, fs, fe, fn) if parts else f

    def em_stats(self):
        """compute_na_expect + compute_cab_expect (belief_propagation.cpp:428-440, :892-989) over all ranks."""
        Q = self.Q
        row = np.zeros(2 * 32 + 32 * 32, np.float64)
        nc = C.c_uint32()
        _check(lib().sbmbp_dist_node_stats(self._e, None, _p(row), C.byref(nc)))
        row = self._allreduce(row)
        na, nna = row[:Q].copy(), row[32:32 + Q].copy()
        r = self._energy(1)
        qt = int(round((len(r) - 4) ** 0.5))
        N = float(self.plan.N_global)
        cab = np.zeros((Q, Q), np.float64)
        for q1 in range(Q):
            for q2 in range(q1, Q):
                v = r[4 + q1 * qt + q2]
                if na[q1] > 1e-50 and na[q2] > 1e-50:  # :970
                    w = na if self.dc == 0 else nna  # :972-985
                    v *= (2.0 * N if q1 == q2 else N) / (w[q1] * w[q2])
                cab[q1, q2] = cab[q2, q1] = v
        return na, nna, cab

    def learning(self, state, learning_conv_crit=1e-6, learning_max_time=100, learning_rate=0.2, dumping_rate=1.0):
        """learning() (belief_propagation.cpp:14-51) with learning_step (:53-75) over all ranks: every rank runs the same
        host loop on all-reduced statistics, so the parameters stay identical everywhere.  Returns (eta, cab, na, iters)
        like api.belief_propagation.learning."""
        Q, N = self.Q, int(self.plan.N_global)
        na = np.ascontiguousarray(state.na, np.uint32).copy()
        cab = np.ascontiguousarray(state.cab, np.float64).reshape(Q, Q).copy()
        lr = float(np.float32(learning_rate))
        crit = np.float32(learning_conv_crit)
        fold, fdiff = 0.0, 1.0
        it = 0
        from .api import bp_blockmodel_state

        self.expand_bp_params(bp_blockmodel_state(na, cab))
        for it in range(int(learning_max_time)):
            if fdiff < crit:
                crit = np.float32(crit * np.float32(0.1))
            self.converge(float(crit), int(learning_max_time), dumping_rate)
            na_e, _, cab_e = self.em_stats()
            fnew = self.compute_free_energy()
            fdiff = abs(fnew - fold)
            fold = fnew
            if not np.isfinite(fold) or fdiff < crit:
                break
            rest = N
            for i in range(Q - 1):  # n_a is an unsigned int, truncated every step (:60)
                na[i] = np.uint32(int(lr * na_e[i] + (1.0 - lr) * float(na[i])))
                rest -= int(na[i])
            na[Q - 1] = np.uint32(rest)
            cab = lr * cab_e + (1.0 - lr) * cab
            self.expand_bp_params(bp_blockmodel_state(na, cab))
        return na.astype(np.float64) / N, cab, na, it

    def stats(self):
        eu, sw, la = C.c_uint64(), C.c_uint64(), C.c_uint64()
        bpe, sec = C.c_double(), C.c_double()
        _check(lib().sbmbp_stats(self._e, C.byref(eu), C.byref(sw), C.byref(la), C.byref(bpe), C.byref(sec)))
        return {"edge_updates": eu.value, "sweeps": sw.value, "launches": la.value, "bytes_per_edge": bpe.value}

    def close(self):
        if getattr(self, "_e", None):
            self._torch.cuda.synchronize()
            lib().sbmbp_destroy(self._e)
            self._e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_degrees(plan, group=None, device=None):
    """The degrees of ALL nodes, u32[N_global]: the ranks' degree arrays (node ranges are contiguous and ordered by rank)
    all-gathered.  device: CUDA device index for an NCCL group, None for host tensors (gloo; the CPU tests)."""
    import torch
    import torch.distributed as dist

    rp, _ = plan.csr()
    deg = np.diff(rp.astype(np.int64)).astype(np.int32)
    if plan.world > 1:
        dev = torch.device("cpu") if device is None else torch.device("cuda", int(device))
        starts = [int(x) for x in plan.starts]
        sizes = [starts[r + 1] - starts[r] for r in range(plan.world)]
        pad = max(sizes)
        buf = torch.zeros(pad, dtype=torch.int32, device=dev)
        buf[: sizes[plan.rank]] = torch.from_numpy(deg).to(dev)
        if device is None:  # gloo
            rows = [torch.empty(pad, dtype=torch.int32, device=dev) for _ in range(plan.world)]
            dist.all_gather(rows, buf, group=group)
        else:  # NCCL: one flat collective
            out = torch.empty((plan.world, pad), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(out, buf, group=group)
            rows = [out[r] for r in range(plan.world)]
        deg = torch.cat([rows[r][: sizes[r]] for r in range(plan.world)]).cpu().numpy()
    full = np.ascontiguousarray(deg.astype(np.uint32))
    assert full.size == plan.N_global
    return full


def plan_finished(plan):
    ga = C.POINTER(C.c_uint32)()
    return lib().sbmbp_plan_layout(plan._p, C.byref(ga), None, None, None, None, None) == 0
