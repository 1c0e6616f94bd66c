"""Worker for the multi-rank tests (launched by torch.distributed.run from test_dist.py).

mode "plan": host-side plan on CPU ranks over gloo -- every producer's position for a message must be where the
             owner's gather index looks for it; each rank's gather index must be a permutation of its buffer.
mode "gpu" : one GPU per rank over NCCL -- the multi-GPU engine against the single-GPU engine on the same graph
             from the same state: sweep by sweep, then converge.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_graph(N, Q, seed):
    from sbm_bp_b200 import generators

    return generators.planted_sbm_epsilon_c(N, Q, 0.15, 5.0, seed=seed)


def check_plan(rank, world):
    import torch.distributed as dist

    from sbm_bp_b200 import generators
    from sbm_bp_b200.dist import DistPlan

    for (N, Q, prec, region, tps) in ((3000, 2, "f64", "0.002", "2"), (2500, 4, "f32", "0", "64")):
        os.environ["SBMBP_REGION_MB"] = region
        os.environ["SBMBP_SUPERTILE"] = tps
        u, v, sizes, upper = make_graph(N, Q, 7)
        starts = generators.rank_ranges(N, world)
        if N == 2500:  # uneven ranges
            starts = np.array([0] + [int(N * (k + 0.37) / world) for k in range(1, world)] + [N], np.uint32)
        plan = DistPlan(u, v, N, starts, rank, world, Q, prec)
        plan.exchange()
        gather, pos, info, pos_slot = plan.layout()
        row_ptr, col = plan.csr()
        assert sorted(gather.tolist()) == list(range(plan.M_local)), "gather is not a permutation"
        # halo exchange tables: the kernels' words, the outbox order and the shipping descriptors
        xt = plan.exchange_tables()
        rpos = xt["rpos"]
        assert sorted(rpos.tolist()) == sorted(pos_slot.tolist()), "tile order is a permutation of slot order"
        remote = (pos >> 31) == 1
        assert ((rpos >> 29) != rank).tolist() == remote.tolist(), "bit 31 marks exactly the remote entries"
        assert np.array_equal(pos[~remote], rpos[~remote] & ((1 << 29) - 1)), "local word = position in the own buffer"
        out_idx = pos[remote] & 0x7fffffff
        assert sorted(out_idx.tolist()) == list(range(xt["n_remote"])), "outbox indices are a permutation"
        # every outbox entry is covered by exactly one descriptor, which sends it to the owner / position of its message
        dest = np.full(xt["n_remote"], -1, np.int64)
        for (src, dst, ln, rk) in xt["ship"].tolist():
            assert ln > 0 and rk != rank and (dest[src:src + ln] == -1).all()
            dest[src:src + ln] = (rk << 29) + dst + np.arange(ln)
        assert (dest >= 0).all() and np.array_equal(dest[out_idx], rpos[remote].astype(np.int64)), "shipping descriptors"
        # descriptors are grouped by super-tile, sorted by source, and tile each super-tile's outbox range without gaps
        for sp in range(xt["nsuper"]):
            d = xt["ship"][xt["ship_start"][sp]:xt["ship_start"][sp + 1]]
            at = int(xt["out_start"][sp])
            for (src, dst, ln, rk) in d.tolist():
                assert src == at
                at += ln
            assert at == int(xt["out_start"][sp + 1])
        everyone = [None] * world
        dist.all_gather_object(everyone, (starts, row_ptr, col, gather))
        lo = int(starts[rank])
        node_of_slot = np.repeat(np.arange(plan.N_local), np.diff(row_ptr.astype(np.int64)))
        bad = 0
        for s in range(plan.M_local):
            j, i = lo + int(node_of_slot[s]), int(col[s])
            o = int(np.searchsorted(starts, i, side="right") - 1)
            assert o == (int(pos_slot[s]) >> 29), "owner bits"
            _, rp_o, col_o, gather_o = everyone[o]
            il = i - int(starts[o])
            a, b = int(rp_o[il]), int(rp_o[il + 1])
            e = a + int(np.searchsorted(col_o[a:b], j))
            assert col_o[e] == j
            bad += int(gather_o[e]) != (int(pos_slot[s]) & ((1 << 29) - 1))
        assert bad == 0, "%d out-messages would land in the wrong slot" % bad
        # the degree exchange of the degree-corrected free energy / EM (dist.gather_degrees): every rank ends up with the
        # degrees of ALL nodes, as the single-process graph builder counts them
        from sbm_bp_b200 import api
        from sbm_bp_b200.dist import gather_degrees

        want_deg = api.blockmodel_t(sizes, (u, v)).csr()[3]
        got_deg = gather_degrees(plan)
        assert got_deg.dtype == np.uint32 and np.array_equal(got_deg, want_deg), "degree all-gather"
        # per-rank generator: the union over ranks equals what every rank would need from the global graph
        gu, gv, _, _, gst = generators.planted_sbm_rank(N, Q, 0.15, 5.0, rank, world, seed=3)
        pairs = [None] * world
        dist.all_gather_object(pairs, (gu, gv))
        mine = set()
        for (pu, pv) in pairs:
            for a, b in zip(pu.tolist(), pv.tolist()):
                if gst[rank] <= a < gst[rank + 1] or gst[rank] <= b < gst[rank + 1]:
                    mine.add((min(a, b), max(a, b)))
        own = set((min(a, b), max(a, b)) for a, b in zip(gu.tolist(), gv.tolist()))
        assert mine == own, "rank-local generation is not consistent across ranks"
    print("rank %d plan ok" % rank)


def check_gpu(rank, world):
    import torch
    import torch.distributed as dist

    from sbm_bp_b200 import api, generators
    from sbm_bp_b200.dist import DistPlan, distributed_belief_propagation

    torch.cuda.set_device(rank)
    # the last three go through the GENERAL kernel (padded Q, deg_corr_flag 2, beta != 1) and its ship-and-publish kernel
    for (N, Q, prec, region, dc, tps, beta) in ((6000, 2, "f64", "0.01", 0, "3", 1.0), (5000, 4, "f32", "0", 1, "64", 1.0),
                                                (4000, 2, "f64", "16", 0, "1", 1.0), (4000, 2, "f64", "0", 1, "64", 1.0),
                                                (60000, 2, "f32", "0.25", 0, "8", 1.0), (4000, 3, "f64", "0.01", 0, "4", 1.0),
                                                (4000, 2, "f64", "0.01", 2, "8", 1.0), (4000, 2, "f32", "0", 0, "8", 1.3)):
        os.environ["SBMBP_REGION_MB"] = region
        os.environ["SBMBP_SUPERTILE"] = tps  # tiles per super-tile of the halo exchange
        general = Q == 3 or dc == 2 or beta != 1.0
        u, v, sizes, upper = make_graph(N, Q, 11)
        if dc:
            upper = [x / 25.0 for x in upper]  # the dc models' c_ab lives on the scale c / <d>^2
        starts = generators.rank_ranges(N, world)
        plan = DistPlan(u, v, N, starts, rank, world, Q, prec)
        bp = distributed_belief_propagation(plan, dc)
        bm = api.blockmodel_t(sizes, (u, v), dc)
        state = api.bp_param_from_direct(bm, [1.0 / Q] * Q, upper)
        bp.set_beta(beta)
        bp.expand_bp_params(state)
        rp, col, rev, deg = bm.csr()
        rng = np.random.default_rng(5)
        msg = rng.random((bm.get_M(), Q)) + 0.05
        msg /= msg.sum(1, keepdims=True)
        marg = rng.random((N, Q)) + 0.05
        marg /= marg.sum(1, keepdims=True)
        lo, hi = int(starts[rank]), int(starts[rank + 1])
        a, b = int(rp[lo]), int(rp[hi])
        bp.set_state(msg[a:b], marg[lo:hi])
        bp.init_h()
        single = api.belief_propagation(bm, prec, device=rank)
        single.set_beta(beta)
        single.expand_bp_params(state)
        single.set_state(msg, marg)
        assert ("bp_sweep_kernel" in single.sweep_kernel_name()) == general
        tol = 1e-12 if prec == "f64" else 1e-5
        for sweep in range(3):
            md_d = bp.sweep(1.0 if sweep != 1 else 0.7)
            md_s = single.sweep(1.0 if sweep != 1 else 0.7)
            assert abs(md_d - md_s) < tol, (sweep, md_d, md_s)
            m_d, g_d = bp.get_state()
            m_s, g_s, _ = single.get_state()
            err = max(np.max(np.abs(m_d - m_s[a:b]) / np.abs(m_s[a:b])), np.max(np.abs(g_d - g_s[lo:hi]) / np.abs(g_s[lo:hi])))
            assert err < tol, (sweep, err)
        # a batch of sweeps with no host in between (device-side flags between the ranks) == the same sweeps on one GPU
        bp.sweeps_async(5, 1.0)
        single.sweeps_async(5, 1.0)
        single.sync()
        m_d, g_d = bp.get_state()
        m_s, g_s, _ = single.get_state()
        err = max(np.max(np.abs(m_d - m_s[a:b]) / np.abs(m_s[a:b])), np.max(np.abs(g_d - g_s[lo:hi]) / np.abs(g_s[lo:hi])))
        assert err < 10 * tol, ("batch", err)
        it_d = bp.converge(5e-6, 300, 1.0)
        it_s = single.converge(5e-6, 300, 1.0)
        assert it_d == it_s and it_d >= 0, (it_d, it_s)
        g_d = bp.get_marginals()
        g_s = single.get_marginals()
        assert np.max(np.abs(g_d - g_s[lo:hi])) < (1e-10 if prec == "f64" else 1e-4)
        conf = bm.memberships
        ov_d = bp.compute_overlap(conf[lo:hi])
        ov_s = single.compute_overlap()
        assert abs(ov_d - ov_s) < 1e-9, (ov_d, ov_s)
        if True:
            # free energy and EM statistics over the ranks (dc != 0: with the all-gathered degrees of remote neighbours) == the single-GPU engine at the same (converged) state
            f_d = np.array(bp.compute_free_energy(parts=True))
            f_s = np.array(single.compute_free_energy(parts=True))
            ftol = 1e-10 if prec == "f64" else 1e-5
            assert np.max(np.abs(f_d - f_s) / (np.abs(f_s) + 1e-300)) < ftol, (f_d, f_s)
            na_d, nna_d, cab_d = bp.em_stats()
            na_s, nna_s, cab_s = single.em_stats()
            assert np.max(np.abs(na_d - na_s) / na_s) < ftol and np.max(np.abs(nna_d - nna_s) / nna_s) < ftol
            assert np.max(np.abs(cab_d - cab_s) / cab_s) < ftol, (cab_d, cab_s)
        if prec == "f64" and not general:
            # learning() over the ranks: same EM trajectory as the single-GPU driver (same synchronous schedule)
            start = api.bp_param_from_direct(bm, [0.45, 0.55] if Q == 2 else [1.0 / Q] * Q, [u_ * 1.3 for u_ in upper])
            bp.set_state(msg[a:b], marg[lo:hi])
            single.set_state(msg, marg)
            eta_d, cab_d, na_d, it_d2 = bp.learning(start, 1e-6, 60, 0.2, 1.0)
            eta_s, cab_s, na_s, it_s2 = single.learning(start, 1e-6, 60, 0.2, 1.0)
            assert it_d2 == it_s2, (it_d2, it_s2)
            assert (np.asarray(na_d) == np.asarray(na_s)).all(), (na_d, na_s)
            assert np.max(np.abs(np.asarray(cab_d) - np.asarray(cab_s)) / np.asarray(cab_s)) < 1e-8, (cab_d, cab_s)
        dist.barrier()
        if rank == 0:
            print("dist ok: N=%d Q=%d %s dc=%d niter=%d overlap=%.4f" % (N, Q, prec, dc, it_d, ov_d))
        bp.close()
    check_relabelled_partition(rank, world)
    dist.barrier()


def check_relabelled_partition(rank, world):
    """The partition as BASELINE.json states it -- random relabelling, ranges balanced on sum (d_i + const) -- on a
    power-law DC-SBM (hubs up to a few hundred edges, deg_corr_flag 1): the multi-GPU engine on the SHUFFLED labels against
    the single-GPU engine on the ORIGINAL labels.  Per node quantities must agree after mapping back (marginals, overlap,
    max-diff per sweep, sweep count); relabelling changes neighbour order, hence products only to rounding."""
    import torch.distributed as dist

    from sbm_bp_b200 import api, generators
    from sbm_bp_b200.dist import DistPlan, distributed_belief_propagation

    os.environ["SBMBP_REGION_MB"] = "0.05"
    os.environ["SBMBP_SUPERTILE"] = "4"
    N, Q, dc = 20000, 4, 1
    u, v, sizes, _ = generators.dc_sbm_powerlaw(N, Q, gamma=2.5, k_min=2.0, ratio=10.0, seed=5)
    P = generators.Partition(N, world, u, v, relabel_seed=11)
    assert not np.array_equal(P.starts, generators.rank_ranges(N, world))  # degree balance moved the cut
    bm = api.blockmodel_t(sizes, (u, v), dc)  # original labels, one GPU
    deg = bm.csr()[3]
    grp = bm.memberships
    D = np.bincount(grp, weights=deg, minlength=Q)
    m = np.zeros((Q, Q))
    np.add.at(m, (grp[u], grp[v]), 1.0)
    m = m + m.T
    cab = N * m / np.outer(D, D)
    state = api.bp_blockmodel_state(np.asarray(sizes, np.uint32), cab)
    single = api.belief_propagation(bm, "f64", device=rank)
    single.expand_bp_params(state)
    rng = np.random.default_rng(9)
    marg = rng.random((N, Q)) + 0.05
    marg /= marg.sum(1, keepdims=True)
    # messages: the same value on an edge whatever the labelling -> drawn per (source, destination) pair
    rp, col, _, _ = bm.csr()
    dst = np.repeat(np.arange(N), np.diff(rp.astype(np.int64)))
    key = (col.astype(np.uint64) * np.uint64(N) + dst.astype(np.uint64))
    msg = np.stack([np.sin(key.astype(np.float64) * (0.37 + q)) ** 2 + 0.05 for q in range(Q)], axis=1)
    msg /= msg.sum(1, keepdims=True)
    single.set_state(msg, marg)
    # shuffled labels, world ranks
    uu, vv = P.relabel(u, v)
    plan = DistPlan(uu, vv, N, P.starts, rank, world, Q, "f64")
    bp = distributed_belief_propagation(plan, dc)
    bp.expand_bp_params(state)
    rp_l, col_l = plan.csr()  # rows of the owned nodes (new ids), neighbours ascending in NEW ids
    lo = int(P.starts[rank])
    dst_l = np.repeat(np.arange(lo, lo + plan.N_local), np.diff(rp_l.astype(np.int64)))
    key_l = P.old_id[col_l].astype(np.uint64) * np.uint64(N) + P.old_id[dst_l].astype(np.uint64)
    msg_l = np.stack([np.sin(key_l.astype(np.float64) * (0.37 + q)) ** 2 + 0.05 for q in range(Q)], axis=1)
    msg_l /= msg_l.sum(1, keepdims=True)
    owned = P.owned(rank)
    bp.set_state(msg_l, marg[owned])
    bp.init_h()
    for sweep in range(3):
        md_d, md_s = bp.sweep(1.0), single.sweep(1.0)
        assert abs(md_d - md_s) < 1e-11, (sweep, md_d, md_s)
        g_d, g_s = bp.get_marginals(), single.get_marginals()
        assert np.max(np.abs(g_d - g_s[owned]) / g_s[owned]) < 1e-10, sweep
    it_d, it_s = bp.converge(5e-6, 400, 1.0), single.converge(5e-6, 400, 1.0)
    assert it_d == it_s and it_d >= 0, (it_d, it_s)
    assert np.max(np.abs(bp.get_marginals() - single.get_marginals()[owned])) < 1e-9
    ov_d, ov_s = bp.compute_overlap(grp[owned]), single.compute_overlap()
    assert abs(ov_d - ov_s) < 1e-9, (ov_d, ov_s)
    f_d, f_s = bp.compute_free_energy(), single.compute_free_energy()
    assert abs(f_d - f_s) <= 1e-9 * abs(f_s), (f_d, f_s)
    dist.barrier()
    if rank == 0:
        print("dist ok: relabelled + degree-balanced partition, DC-SBM N=%d max degree %d, ranges %s, niter=%d overlap=%.4f"
              % (N, int(deg.max()), P.starts.tolist(), it_d, ov_d))
    bp.close()


def main():
    import torch.distributed as dist

    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "plan":
        dist.init_process_group("gloo")
        check_plan(rank, world)
    else:
        import torch

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
        check_gpu(rank, world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
