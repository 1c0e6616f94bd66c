"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors (and against the compiled
reference itself where oracle/_ref is present), the host graph builder / loader / parameter constructors,
and the C-ABI surface of libsbmbp.so.  No device work.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, rel_err, upper_from_full


# --------------------------------------------------------------------------- oracle pinned by the goldens

@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_oracle_reproduces_reference_sweep_golden(built, name):
    """bp_oracle.c == the reference's own routines, bit for bit, on every golden sweep case."""
    from oracle.oracle import Oracle

    g = load_golden(name)
    O = Oracle(g["u"], g["v"], g["sizes"], int(g["dc"]))
    rp, col, rl, rg = O.csr()
    assert (rp == g["row_ptr"]).all() and (col == g["col"]).all()
    assert (rl == g["rev_local"]).all() and (rg == g["rev_global"]).all()
    assert O.M == int(g["M"]) and O.E == int(g["E"]) and O.max_degree == int(g["max_degree"])
    O.init_messages(int(g["seed"]), float(g["beta"]))
    O.set_params_raw(g["na"], g["cab"])
    msg0, marg0, h0 = O.get_state()
    assert (msg0 == g["msg0"]).all() and (marg0 == g["marg0"]).all()  # mt19937 + uniform_real_distribution restated
    O.init_h()
    assert (O.get_state()[2] == g["h0"]).all()
    nm, ng, nd, md = O.jacobi_sweep(float(g["damping"]))
    assert (nm == g["new_msg"]).all() and (ng == g["new_marg"]).all() and (nd == g["node_diff"]).all()
    assert md == float(g["maxdiff"])
    assert O.f_site() == float(g["f_site"]) and O.f_edge() == float(g["f_edge"])
    if O.N <= 1500:
        assert O.f_non_edge() == float(g["f_non_edge"])
    assert O.overlap() == float(g["overlap"])
    na, nna, cab = O.em_stats()
    assert (na == g["na_expect"]).all() and (nna == g["nna_expect"]).all() and (cab == g["cab_expect"]).all()
    if "entropy" in g:
        assert O.entropy_site() == float(g["entropy_site"]) and O.entropy_edge() == float(g["entropy_edge"])
        assert O.entropy_non_edge() == float(g["entropy_non_edge"])


@pytest.mark.parametrize("name", ["converge_cfg1_readme", "converge_cfg1_eps01", "converge_cfg1_eps01_dc1"])
def test_oracle_reproduces_reference_converge_golden(built, name):
    """Same seed -> same random-sequential trajectory as the reference's converge(): niter and state identical."""
    from oracle.oracle import Oracle

    g = load_golden(name)
    O = Oracle(g["u"], g["v"], g["sizes"], int(g["dc"]))
    O.init_messages(int(g["seed"]))
    O.set_params_raw(g["na"], g["cab"])
    assert O.converge(float(g["crit"]), int(g["tmax"]), 1.0) == int(g["niter"])
    assert (O.get_state()[1] == g["marg"]).all()
    assert O.free_energy() == float(g["f"]) and O.overlap() == float(g["overlap"])
    if "entropy" in g:
        assert O.entropy() == float(g["entropy"])


def test_oracle_reproduces_reference_learning_golden(built):
    from oracle.oracle import Oracle

    g = load_golden("learn_cfg1_515")
    O = Oracle(g["u"], g["v"], g["sizes"], int(g["dc"]))
    O.init_messages(int(g["seed"]))
    O.set_params_raw(g["na0"], g["cab0"])
    na, cab, eta, it = O.learning(float(g["crit"]), int(g["tmax"]), float(g["lr"]), 1.0, sync=False)
    assert (na == g["na"]).all() and (cab == g["cab"]).all() and (eta == g["eta"]).all()


def test_sync_schedule_reaches_reference_fixed_point(built):
    """the synchronous schedule (what the GPU runs) reaches the reference's fixed point."""
    from oracle.oracle import Oracle

    g = load_golden("converge_cfg1_eps01")
    O = Oracle(g["u"], g["v"], g["sizes"], 0)
    O.init_messages(3)
    O.set_params_raw(g["na"], g["cab"])
    assert O.sync_converge(1e-9, 2000, 1.0) >= 0
    marg = O.get_state()[1]
    err = min(np.max(np.abs(marg - g["marg"])), np.max(np.abs(marg[:, ::-1] - g["marg"])))
    assert err < 1e-4
    assert abs(O.free_energy() - float(g["f"])) <= 1e-6 * abs(float(g["f"]))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsbmbp_ref.so")),
                    reason="compiled reference not present")
@pytest.mark.parametrize("dc,Q", [(0, 2), (1, 3), (2, 2), (0, 5)])
def test_oracle_matches_compiled_reference_live(built, dc, Q):
    """Fresh seeded inputs through both checkers: identical outputs (the pin of the restatement)."""
    from oracle.oracle import Oracle, Reference
    from sbm_bp_b200 import generators

    N = 600
    sizes = [N // Q] * Q
    sizes[-1] += N - sum(sizes)
    rng = np.random.default_rng(Q * 10 + dc)
    cab = rng.uniform(0.5, 2.0, (Q, Q))
    cab = (cab + cab.T) / 2 + np.diag(rng.uniform(3, 6, Q))
    u, v = generators.planted_sbm(sizes, cab, seed=dc + 1)
    if dc:
        cab = cab / 20.0
    pa = np.asarray(sizes) / N
    R = Reference(u, v, sizes, dc)
    O = Oracle(u, v, sizes, dc)
    for a, b in zip(R.csr(), O.csr()):
        assert (a == b).all()
    R.init_messages(9, 0.9)
    O.init_messages(9, 0.9)
    R.set_params_direct(pa, upper_from_full(cab))
    O.set_params_direct(pa, upper_from_full(cab))
    for a, b in zip(R.get_params(), O.get_params()):
        assert (a == b).all()
    ra, oa = R.jacobi_sweep(0.8), O.jacobi_sweep(0.8)
    for a, b in zip(ra[:3], oa[:3]):
        assert (a == b).all()
    assert ra[3] == oa[3]
    assert R.converge(5e-6, 200, 1.0) == O.converge(5e-6, 200, 1.0)
    for a, b in zip(R.get_state(), O.get_state()):
        assert (a == b).all()
    assert R.free_energy() == O.free_energy()
    for a, b in zip(R.em_stats(), O.em_stats()):
        assert (a == b).all()
    R.learning_step(0.2)
    O.learning_step(0.2)
    for a, b in zip(R.get_params(), O.get_params()):
        assert (a == b).all()


def test_oracle_loader_quirks(built, tmp_path):
    """load_edge_list quirks: blank line repeats the previous pair, a non-numeric line pushes
    (0, stale v), a one-number line pushes (n, stale v)."""
    from oracle.oracle import Oracle

    p = tmp_path / "quirks.edgelist"
    p.write_text("1 2\n\n3\t4 extra\nabc def\n7\n  5   6\n")
    u, v = Oracle.load_edge_list(str(p))
    assert u.tolist() == [1, 1, 3, 0, 7, 5] and v.tolist() == [2, 2, 4, 4, 4, 6]
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libsbmbp_ref.so")
    if os.path.exists(ref_so):
        from oracle.oracle import Reference

        ur, vr = Reference.load_edge_list(str(p))
        assert ur.tolist() == u.tolist() and vr.tolist() == v.tolist()


# --------------------------------------------------------------------------- host side of the product

def test_c_abi_exports_every_declared_symbol(built):
    from sbm_bp_b200 import api

    header = open(os.path.join(ROOT, "include", "sbmbp.h")).read()
    declared = sorted(set(re.findall(r"\b(sbmbp_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = api.lib()
    for sym in declared:
        assert hasattr(lib, sym), "libsbmbp.so does not export " + sym
    assert sorted(api.SYMBOLS) == declared
    assert b"sm_100a" in lib.sbmbp_version()


@pytest.mark.parametrize("name", ["sweep_cfg1_eps01", "sweep_hub_q2", "sweep_hub_q4_dc1"])
def test_graph_builder_bit_exact_with_reference(built, name):
    """CSR, reverse index, degrees, E, max degree == graph_neis_ / graph_neis_inv_ / blockmodel_t of the reference."""
    from sbm_bp_b200 import api

    g = load_golden(name)
    bm = api.blockmodel_t(g["sizes"], (g["u"], g["v"]), int(g["dc"]))
    rp, col, rev, deg = bm.csr()
    assert rp.dtype == np.uint64 and col.dtype == np.uint32 and rev.dtype == np.uint32
    assert (rp == g["row_ptr"]).all() and (col == g["col"]).all() and (rev == g["rev_global"]).all()
    assert ((rev - rp[col]) == g["rev_local"]).all()
    assert (deg == np.diff(rp.astype(np.int64))).all()
    assert bm.get_M() == int(g["M"]) and bm.get_E() == int(g["E"]) and bm.get_graph_max_degree() == int(g["max_degree"])
    assert (rev[rev] == np.arange(len(rev))).all()  # the reverse index is an involution


def test_graph_builder_selfloops_duplicates_threads(built):
    from oracle.oracle import Oracle
    from sbm_bp_b200 import api

    rng = np.random.default_rng(0)
    for N, P in ((300, 3000), (50000, 700000)):  # the second size takes the multi-threaded path
        u = rng.integers(0, N, P).astype(np.uint32)
        v = rng.integers(0, N, P).astype(np.uint32)
        bm = api.blockmodel_t([N // 2, N - N // 2], (u, v))
        O = Oracle(u, v, [N // 2, N - N // 2], 0)
        a, b = O.csr(), bm.csr()
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[3] == b[2]).all()
        assert bm.get_E() == O.E and bm.get_graph_max_degree() == O.max_degree


def test_loader_well_formed_and_errors(built, tmp_path):
    from sbm_bp_b200 import api

    p = tmp_path / "ok.edgelist"
    p.write_text("0 1\n\n2\t3 extra columns\n   4    0   \n1 0\n")
    u, v = api.load_edge_list(str(p))
    assert u.tolist() == [0, 2, 4, 1] and v.tolist() == [1, 3, 0, 0]
    bm = api.blockmodel_t([3, 2], str(p))
    assert bm.get_M() == 6 and bm.get_E() == 3
    bad = tmp_path / "bad.edgelist"
    bad.write_text("0 1\n# comment\n")
    with pytest.raises(api.SbmbpError) as ei:
        api.load_edge_list(str(bad))
    assert ei.value.code == 3 and "line 2" in str(ei.value)
    with pytest.raises(api.SbmbpError) as ei:
        api.load_edge_list(str(tmp_path / "missing.edgelist"))
    assert ei.value.code == 2
    with pytest.raises(api.SbmbpError) as ei:
        api.blockmodel_t([1, 1], str(p))  # ids >= N
    assert ei.value.code == 4


def test_parameter_constructors_match_reference_quirks(built):
    """bp_param_from_direct keeps unsigned(int(pa*N)) for every block (no remainder fix-up);
    --cab is the upper triangle; epsilon_c as blockmodel.cpp:251-257."""
    from oracle.oracle import Oracle
    from sbm_bp_b200 import api

    u = np.array([0, 1], np.uint32)
    v = np.array([1, 2], np.uint32)
    for sizes, pa, upper in (([3, 4], [0.3, 0.7], [5, 1, 4]), ([3, 3, 4], [0.33, 0.33, 0.34], [6, 1, .5, 5, .8, 7])):
        bm = api.blockmodel_t(sizes, (u, v))
        st = api.bp_param_from_direct(bm, pa, upper)
        O = Oracle(u, v, sizes, 0)
        O.set_params_direct(pa, upper)
        na, cab, eta = O.get_params()
        assert (st.na == na).all() and (st.cab == cab).all()
        for eps in (0.1, -1.0):
            st = api.bp_param_from_epsilon_c(bm, eps, 3.0)
            O.set_params_epsilon_c(eps, 3.0)
            na, cab, eta = O.get_params()
            assert (st.na == na).all() and (st.cab == cab).all()


def test_engine_fails_loudly_without_gpu(built):
    """No CPU fallback: on a box without a B200, creating an engine is an error, not a slow path."""
    import torch

    from sbm_bp_b200 import api

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    bm = api.blockmodel_t([2, 2], (np.array([0], np.uint32), np.array([3], np.uint32)))
    with pytest.raises(api.SbmbpError) as ei:
        api.belief_propagation(bm)
    assert ei.value.code == 6


def test_cli_validation_messages(built, tmp_path):
    """bin/bp reproduces the reference's option validation (main.cpp:154-206) without needing a device."""
    import subprocess

    exe = os.path.join(ROOT, "bin", "bp")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "BP algorithms for the SBM" in r.stderr
    r = subprocess.run([exe, "-n", "5", "5", "-m", "infer"], capture_output=True, text=True)
    assert r.returncode == 1 and "edge_list_path is required" in r.stderr
    r = subprocess.run([exe, "-l", "x", "-n", "5", "5"], capture_output=True, text=True)
    assert r.returncode == 1 and "mode is required" in r.stderr
    r = subprocess.run([exe, "-l", "x", "-m", "infer"], capture_output=True, text=True)
    assert r.returncode == 1 and "n is required" in r.stderr
    r = subprocess.run([exe, "-l", "x", "-n", "5", "5", "-m", "infer", "--pa", ".5", ".5"], capture_output=True, text=True)
    assert r.returncode == 1 and "input both pa/cab" in r.stderr
    # memberships (main.cpp:176-193, :253-265): --mb of the wrong length is the reference's error; --mb_path is refused
    tiny = str(tmp_path / "tiny.edgelist")
    with open(tiny, "w") as fh:
        fh.write("0 1\n1 2\n2 3\n")
    base = [exe, "-l", tiny, "-n", "2", "2", "-m", "infer", "--epsilon_c", "0.5", "2"]
    r = subprocess.run(base + ["--mb", "0", "1", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "does not fit the number of nodes" in r.stderr
    r = subprocess.run(base + ["--mb_path", "nowhere"], capture_output=True, text=True)
    assert r.returncode == 1 and "mb_path" in r.stderr
    r = subprocess.run(base + ["--mb", "0", "1", "0", "1", "--mb_n"], capture_output=True, text=True)
    assert r.returncode == 1 and "just select one option" in r.stderr


# --------------------------------------------------------------------------- degree-class (ELL) message layout

@pytest.mark.parametrize("region_slots", [0, 700, 3000])
@pytest.mark.parametrize("hubs", [False, True])
def test_ell_layout_invariants(built, region_slots, hubs):
    """What bp_sweep_ell_kernel relies on (csrc/sweep_ell.cuh): pos is a permutation of the slots, the message
    into a node sits in the region of that node's bucket, every node of degree < 32 owns exactly one lane of one
    chunk, and the index words of lane r / slot l of a chunk -- base + 32 d k + 32 l + r -- name the in-message and
    the out-message of that slot."""
    from sbm_bp_b200 import api, generators

    N = 3000
    u, v = generators.planted_sbm([N // 2, N - N // 2], np.array([[5.0, 1.0], [1.0, 5.0]]), seed=11)
    if hubs:  # a few nodes of degree >= 32 (left to the warp / hub kernels) and isolated nodes stay isolated
        rng = np.random.default_rng(3)
        hu = np.repeat(np.array([7, 1500, 2999], np.uint32), [40, 90, 300])
        hv = rng.integers(0, N, hu.size).astype(np.uint32)
        u, v = np.concatenate([u, hu]), np.concatenate([v, hv])
    bm = api.blockmodel_t([N // 2, N - N // 2], (u, v))
    rp, col, rev, deg = bm.csr()
    M = len(col)
    L = api.ell_layout(bm, region_slots)
    pos, gather = L["pos"].astype(np.int64), L["gather"].astype(np.int64)
    assert sorted(pos.tolist()) == list(range(M))
    assert (gather == pos[rev]).all()
    # regions: the message out of slot s goes INTO col[s]; buckets are runs of consecutive nodes
    owner = np.repeat(np.arange(N), np.diff(rp).astype(np.int64))  # node of each slot
    dest_slot_node = col.astype(np.int64)
    order = np.argsort(pos)
    dest_in_buffer_order = dest_slot_node[order]
    # positions grouped by destination bucket: the bucket id along the buffer never decreases, and every bucket is a
    # contiguous node range whose region equals its in-slot range
    nb = L["n_buckets"]
    assert nb >= 1 and (region_slots != 0 or nb == 1)
    if nb > 1:
        starts = [0]
        nxt = region_slots
        for i in range(1, N):
            if rp[i] >= nxt:
                starts.append(i)
                nxt = int(rp[i]) + region_slots
        assert len(starts) == nb
        bucket_of = np.searchsorted(np.array(starts), np.arange(N), side="right") - 1
        b_along = bucket_of[dest_in_buffer_order]
        assert (np.diff(b_along) >= 0).all()
        for b, first in enumerate(starts):
            last = starts[b + 1] if b + 1 < nb else N
            sel = bucket_of[dest_slot_node] == b
            assert pos[sel].min() == rp[first] and pos[sel].max() == rp[last] - 1
    # classes / chunks / index words
    seen = np.zeros(N, np.int64)
    chunk_next, node_next, base_next = 0, 0, 0
    for d, n, node_first, chunk_first, base in L["classes"].astype(np.int64):
        assert 0 <= d < 32 and n > 0
        assert (node_first, chunk_first, base) == (node_next, chunk_next, base_next)
        nodes = L["node"][node_first:node_first + n].astype(np.int64)
        assert (deg[nodes] == d).all() and (np.diff(nodes) > 0).all()
        seen[nodes] += 1
        for r in range(n):
            k, lane = divmod(r, 32)
            ib = base + 32 * d * k + lane
            s0 = int(rp[nodes[r]])
            for l in range(d):
                assert L["pos_idx"][ib + 32 * l] == pos[s0 + l]
                assert L["rev_idx"][ib + 32 * l] == gather[s0 + l]
        nch = (n + 31) // 32
        chunk_next += nch
        node_next += n
        base_next += nch * 32 * d
    assert chunk_next == L["n_chunks"] and base_next == len(L["rev_idx"])
    assert (seen[deg < 32] == 1).all() and (seen[deg >= 32] == 0).all()
    # coalescing property: the out-messages of one (chunk, slot) that go to the same bucket are consecutive
    if nb == 1:
        for d, n, node_first, chunk_first, base in L["classes"].astype(np.int64)[:4]:
            if d == 0:
                continue
            w = L["pos_idx"][base:base + 32 * d].astype(np.int64).reshape(d, 32)[:, :min(n, 32)]
            assert (np.diff(w, axis=1) == 1).all()


# --------------------------------------------------------------------------- init_messages flags 1-3, bp_conditional

@pytest.mark.parametrize("name", golden_names("init_"))
def test_oracle_reproduces_reference_init_flags_and_clamping(built, name):
    """init_messages flags 1-3 (belief_propagation.cpp:132-215, quirks included) and the frozen planted nodes of
    bp_conditional (:1100-1126): bp_oracle.c == the compiled reference, bit for bit."""
    from oracle.oracle import Oracle

    g = load_golden(name)
    O = Oracle(g["u"], g["v"], g["sizes"], 0)
    assert O.init_messages_flag(int(g["flag"]), g["conf"], int(g["seed"])) == 0
    O.set_conditional(not int(g["learn_mode"]))
    O.set_params_raw(g["na"], g["cab"])
    msg0, marg0, _ = O.get_state()
    assert np.array_equal(msg0, g["msg0"]) and np.array_equal(marg0, g["marg0"])
    assert (O.get_conf_planted() == g["conf"]).all()
    nm, ng, nd, md = O.jacobi_sweep(1.0)
    assert np.array_equal(nm, g["new_msg"]) and np.array_equal(ng, g["new_marg"]) and np.array_equal(nd, g["node_diff"])
    assert md == float(g["maxdiff"])
    if "niter" in g:
        O.seed(int(g["seed"]))  # not the reference's generator state: only the fixed point is compared
        assert O.sync_converge(5e-6, 1000, 1.0) >= 0
        marg = O.get_state()[1]
        assert np.max(np.abs(marg - g["marg"])) < 1e-4
        planted = g["conf"] != -1
        assert np.array_equal(marg[planted], g["marg0"][planted])  # frozen nodes never move


def test_oracle_init_flags_refuse_where_the_reference_asserts(built):
    from oracle.oracle import Oracle

    g = load_golden("init_flag1_full_infer")
    O = Oracle(g["u"], g["v"], g["sizes"], 0)
    for flag in (2, 3):
        assert O.init_messages_flag(flag, g["conf"], 1) == -1  # a belief equal to 1: assert(conf_planted_[i] != 1)


def test_greedy_coloring_is_proper(built):
    """The colouring behind the coloured asynchronous schedule: no edge joins two nodes of one colour (self-loops
    aside), and sparse graphs need only a handful of colours."""
    from sbm_bp_b200 import api

    for name in ("sweep_cfg1_eps01", "sweep_hub_q2"):
        g = load_golden(name)
        bm = api.blockmodel_t(g["sizes"], (g["u"], g["v"]), 0)
        rp, col, rev, deg = bm.csr()
        color, nc = api.graph_coloring(bm)
        src = np.repeat(np.arange(bm.get_N()), np.diff(rp).astype(np.int64))
        proper = (color[src] != color[col]) | (src == col)
        assert proper.all()
        assert nc == int(color.max()) + 1 and nc <= 16


# --------------------------------------------------------------------------- legacy generator / GML

def test_microcanonical_generator_and_gml_round_trip(built, tmp_path):
    """The legacy MODE-NET data path (src/old/bm.cpp:41-170, :192-296) in front of the engine's graph builder: the
    micro-canonical generator plants EXACT edge counts per block pair without self-loops or duplicates; a GML file in
    the legacy writer's layout reads back to the same graph and labels; the legacy reader's quirks (ids are strings
    numbered by appearance, values become groups by first appearance, repeated edges dropped in either orientation,
    unknown node keys skipped) hold; the result feeds sbmbp_graph_from_pairs unchanged."""
    from sbm_bp_b200 import api, generators

    sizes = [300, 500, 200]
    N = sum(sizes)
    cab = np.array([[8, 1, .5], [1, 6, 2], [.5, 2, 9.]])
    u, v = generators.microcanonical_sbm(sizes, cab, seed=3)
    grp = np.repeat(np.arange(3), sizes)
    cnt = np.zeros((3, 3), int)
    np.add.at(cnt, (np.minimum(grp[u], grp[v]), np.maximum(grp[u], grp[v])), 1)
    for a in range(3):
        for b in range(a, 3):
            p = cab[a, b] / N
            want = int(p * sizes[a] * sizes[b]) if a != b else int(p * sizes[a] * (sizes[a] - 1) / 2)
            assert cnt[a, b] == want  # old/bm.cpp:221-222
    assert (u != v).all()
    assert len(set(zip(np.minimum(u, v).tolist(), np.maximum(u, v).tolist()))) == len(u)
    path = str(tmp_path / "g.gml")
    generators.write_gml(path, u, v, grp)
    u2, v2, lab, ids = generators.read_gml(path)
    assert np.array_equal(u2, u) and np.array_equal(v2, v) and np.array_equal(lab, grp) and ids[:2] == ["0", "1"]
    a, b = api.blockmodel_t(sizes, (u, v)), api.blockmodel_t(sizes, (u2, v2))
    for x, y in zip(a.csr(), b.csr()):
        assert np.array_equal(x, y)
    assert a.get_M() == 2 * len(u)  # nothing for the set-based builder to merge
    # reader quirks on a hand-written file
    quirky = str(tmp_path / "q.gml")
    with open(quirky, "w") as fh:
        fh.write("""graph [ directed 0
  node [ id n7 label "x" value red ]
  node [ id n3 value blue ]
  node [ id n9 value red ]
  node [ id n1 ]
  edge [ source n7 target n3 ]
  edge [ weight 2 source n3 target n7 ]
  edge [ source n9 target n3 ]
  edge [ source n7 target n3 ]
]
""")
    qu, qv, ql, qi = generators.read_gml(quirky)
    assert qi == ["n7", "n3", "n9", "n1"] and ql.tolist() == [0, 1, 0, -1]
    assert qu.tolist() == [0, 2] and qv.tolist() == [1, 1]
    bad = str(tmp_path / "bad.gml")
    with open(bad, "w") as fh:
        fh.write("graph [ node [ id a ] edge [ source a target zz ] ]\n")
    with pytest.raises(ValueError):
        generators.read_gml(bad)


def test_non_edge_term_against_the_legacy_closed_form(built):
    """Third cross-check of the free energy: the legacy MODE-NET code closes the non-edge term as
    last_term = 1/2 sum_ab c_ab n_a n_b / N^2 with n_a the expected group sizes (src/old/bm.cpp:882-897); the current
    reference sums log(1 - c/N psi psi) over all non-adjacent pairs (belief_propagation.cpp:675-709).  The two agree up
    to the second-order term and the excluded edges, both O(c / N) relative -- checked on the oracle here, on the engine
    at N = 1M in the GPU suite."""
    from oracle.oracle import Oracle
    from sbm_bp_b200 import generators

    rel = []
    for N in (1000, 8000):
        u, v, sizes, upper = generators.planted_sbm_epsilon_c(N, 2, 0.1, 3.0, seed=7)
        O = Oracle(u, v, sizes, 0)
        O.init_messages(1)
        O.set_params_direct([.5, .5], upper)
        assert O.converge(5e-6, 500, 1.0) >= 0
        na, _, _ = O.em_stats()
        cab = np.array([[upper[0], upper[1]], [upper[1], upper[2]]])
        last_term = 0.5 * float(na @ cab @ na) / N ** 2
        rel.append(abs(O.f_non_edge() + last_term) / last_term)
    assert rel[0] < 2e-2 and rel[1] < rel[0] / 4  # O(1 / N)


def test_oracle_reproduces_mb_rand_shuffle(built):
    """--mb_rand (main.cpp:299-301): blockmodel_t::shuffle is std::shuffle with the run's std::mt19937.  The oracle
    restates libstdc++'s algorithm (Lemire bounded draws, two swap positions per draw while n^2 fits 32 bits, one per
    draw beyond) and must leave the generator exactly where the compiled reference's is: golden initial state and niter
    at N = 1000 (even n, paired path); odd n and the one-draw-per-element path (n >= 65536) against the reference live."""
    from oracle.oracle import Oracle, Reference, have_reference
    from sbm_bp_b200 import generators

    g = load_golden("mbrand_cfg1_eps01")
    O = Oracle(g["u"], g["v"], g["sizes"], 0)
    O.init_messages_mb_rand(int(g["seed"]))
    O.set_params_raw(g["na"], g["cab"])
    msg, marg, _ = O.get_state()
    assert np.array_equal(msg, g["msg0"]) and np.array_equal(marg, g["marg0"])
    assert O.converge(5e-6, 1000, 1.0) == int(g["niter"])
    assert np.max(np.abs(O.get_state()[1] - g["marg"])) < 1e-13
    # a permutation comes out, and it is not the identity
    O.seed(3)
    perm = O.shuffle(1000)
    assert sorted(perm.tolist()) == list(range(1000)) and (perm != np.arange(1000)).sum() > 900
    if not have_reference():
        pytest.skip("compiled reference not present")
    for N in (1001, 70000):
        u, v, sizes, _ = generators.planted_sbm_epsilon_c(N, 2, 0.1, 3.0, seed=2)
        R = Reference(u, v, sizes, 0)
        R.init_messages_mb_rand(5)
        O = Oracle(u, v, sizes, 0)
        O.init_messages_mb_rand(5)
        assert np.array_equal(R.get_state()[0], O.get_state()[0]) and np.array_equal(R.get_state()[1], O.get_state()[1])


def test_bench_reference_arm_prints_one_json_line(built):
    """bench.py --impl reference (the reference's own converge() on the host cores) needs no GPU: its stdout must be
    exactly one JSON line carrying the contract's keys -- build chatter and library banners go to stderr."""
    import json
    import subprocess
    import sys

    from oracle.oracle import have_reference

    if not have_reference():
        pytest.skip("compiled reference not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bp_directed_edge_updates_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0 and d["higher_is_better"] is True


@pytest.mark.parametrize("Q,beta", [(2, 1.0), (4, 1.3), (32, 1.0)])
def test_non_edge_series_host_arithmetic(built, Q, beta):
    """The moment series of compute_f_non_edge (belief_propagation.cpp:675-709) as the engine and dist.py evaluate it
    (sbmbp_non_edge_series_order / _term, host only) against the O(N^2) pair sum it replaces, on random marginals."""
    from sbm_bp_b200 import api

    rng = np.random.default_rng(Q)
    N = 400
    psi = rng.dirichlet(np.ones(Q), N)
    cab = rng.uniform(0.1, 0.6 if Q == 32 else 4.0, (Q, Q))
    cab = (cab + cab.T) / 2
    W = (1.0 - cab / N) ** beta
    want = float(np.sum(np.log(psi @ W @ psi.T)))  # all ordered pairs, i == j included (:686)
    K = api.non_edge_series_order(Q, float(N), beta, cab)
    assert 1 <= K <= 8 and Q ** K <= 1 << 20
    import math

    # k = 0: the normalisation defect of the marginals, 2 N sum_i log(sum_q psi_i^q) (exactly summed here)
    defect = np.array([math.fsum(row) - 1.0 for row in psi])
    got = api.non_edge_series_term(Q, float(N), beta, cab, 0, np.array([float(np.sum(np.log1p(defect)))]))
    for k in range(1, K + 1):
        T = np.zeros(Q ** k)  # T_k = sum_i psi_i^(x)k (symmetric: the digit order does not matter)
        for i in range(N):
            t = psi[i]
            for _ in range(k - 1):
                t = np.multiply.outer(psi[i], t).reshape(-1)
            T += t
        got += api.non_edge_series_term(Q, float(N), beta, cab, k, T)
    ymax = float(np.max(1.0 - W))  # the series expands 1 - W with W rounded to double, like the pair sum above
    rem = 0.5 * N * ymax ** (K + 1) / (K + 1) / (1 - ymax) * 2 * N  # bound on the truncated tail of the pair sum
    assert abs(got - want) <= max(1e-12 * abs(want), rem), (got, want, K, rem)


@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_extended_precision_referee_brackets_the_reference(built, name):
    """oracle.referee_sweep (long double, log domain) is the exact value the FP64 evaluations are measured against:
    on product-domain nodes (degree < 50) the reference's own update agrees with it to 1e-13; on log-domain hubs the
    reference's sequential sum of d logarithms is the less accurate side (its error is what the GPU test allows)."""
    from oracle.oracle import Oracle

    g = load_golden(name)
    O = Oracle(g["u"], g["v"], g["sizes"], int(g["dc"]))
    O.init_messages(int(g["seed"]), float(g["beta"]))
    O.set_params_raw(g["na"], g["cab"])
    O.set_state(g["msg0"], g["marg0"])
    ex_msg, ex_marg, skipped = O.referee_sweep(float(g["damping"]))
    assert not skipped.any()
    deg = np.diff(g["row_ptr"].astype(np.int64))
    src_deg = deg[g["col"]]  # slot e of the reference order holds the message col[e] -> i
    err_msg = np.max(np.abs(g["new_msg"] - ex_msg) / np.abs(ex_msg), axis=1)
    err_marg = np.max(np.abs(g["new_marg"] - ex_marg) / np.abs(ex_marg), axis=1)
    assert err_msg[src_deg < 50].max() < 1e-13 and err_marg[deg < 50].max() < 1e-13
    if (deg >= 50).any():
        worst = max(err_msg[src_deg >= 50].max(), err_marg[deg >= 50].max())
        print("%s: reference vs exact at hubs (max degree %d): %.2e" % (name, deg.max(), worst))
        assert worst < 1e-9  # the reference is still a correct evaluation, just not a 1e-12 one at degree 1500


def test_partition_relabel_and_degree_balance(built):
    """The multi-GPU partition of BASELINE.json: random relabelling (a bijection, undone by to_original) and contiguous
    ranges balanced on sum (d_i + const) -- on a power-law graph equal node counts are far from equal work."""
    from sbm_bp_b200 import generators

    N, world = 40000, 8
    u, v, sizes, _ = generators.dc_sbm_powerlaw(N, 4, gamma=2.5, k_min=2.0, ratio=10.0, seed=3)
    P = generators.Partition(N, world, u, v, relabel_seed=7)
    assert sorted(P.new_id.tolist()) == list(range(N)) and np.array_equal(P.old_id[P.new_id], np.arange(N))
    uu, vv = P.relabel(u, v)
    deg = np.bincount(uu, minlength=N) + np.bincount(vv, minlength=N)
    assert P.starts[0] == 0 and P.starts[-1] == N and np.all(np.diff(P.starts.astype(np.int64)) > 0)
    work = np.array([float(np.sum(deg[P.starts[k]:P.starts[k + 1]] + 4.0)) for k in range(world)])
    assert work.max() / work.mean() < 1.02, work
    x = np.arange(N) * 3.0  # a per-node quantity travels there and back
    assert np.array_equal(P.to_original(x[P.old_id]), x)
    # without relabelling the heavy head of a sorted-by-degree labelling would land on one rank: balance still holds
    order = np.argsort(-deg, kind="stable")
    ren = np.empty(N, np.int64)
    ren[order] = np.arange(N)
    P2 = generators.Partition(N, world, ren[uu], ren[vv], relabel_seed=None)
    d2 = np.bincount(ren[uu], minlength=N) + np.bincount(ren[vv], minlength=N)
    w2 = np.array([float(np.sum(d2[P2.starts[k]:P2.starts[k + 1]] + 4.0)) for k in range(world)])
    assert w2.max() / w2.mean() < 1.05, w2
    assert (np.diff(P2.starts.astype(np.int64)).max() > 3 * np.diff(P2.starts.astype(np.int64)).min())  # unequal node counts
    # isolated / tiny inputs keep one node per rank
    s = generators.balanced_ranges(np.zeros(3), 3)
    assert s.tolist() == [0, 1, 2, 3]
