"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsbmbp_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container only: `python tests/golden/make_golden.py`.

The reference has no golden vectors of its own, so these are outputs of the reference's
own compiled routines on fixed inputs:
  * one synchronous sweep computed by bp_iter_update_psi / bp_iter_update_psi_large_degree from a frozen
    random state (level-1 parity target), for dc 0/1/2, several Q, beta and damping values
  * free-energy pieces, entropy, overlap and EM statistics at that same state
  * the end state of the reference's own converge() / inference() / learning() (level-2 parity target)
Inputs (edge pairs, block sizes, parameters, seeds) are stored next to the outputs so the tests need nothing
from /root/reference at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference  # noqa: E402

import importlib  # noqa: E402

generators = importlib.import_module("sbm_bp_b200.generators")

OUT = os.path.dirname(os.path.abspath(__file__))
SHIPPED = "/root/reference/dataset/N_1000-Q_2-method_cab_ec-eps_0.1-c_3.0.edgelist"


def sweep_case(name, u, v, sizes, dc, seed, beta=1.0, damping=1.0, eps_c=None, pa=None, cab_upper=None,
               energy=True, entropy=True):
    R = Reference(u, v, sizes, dc)
    R.init_messages(seed, beta)
    if eps_c is not None:
        R.set_params_epsilon_c(*eps_c)
    else:
        R.set_params_direct(pa, cab_upper)
    R.init_h()
    na, cab, eta = R.get_params()
    msg0, marg0, h0 = R.get_state()
    new_msg, new_marg, node_diff, maxdiff = R.jacobi_sweep(damping)
    out = dict(u=u, v=v, sizes=np.asarray(sizes, np.uint32), dc=dc, seed=seed, beta=beta, damping=damping,
               na=na, cab=cab, msg0=msg0, marg0=marg0, h0=h0, new_msg=new_msg, new_marg=new_marg,
               node_diff=node_diff, maxdiff=maxdiff, M=R.M, E=R.E, max_degree=R.max_degree)
    rp, col, rl, rg = R.csr()
    out.update(row_ptr=rp, col=col, rev_local=rl, rev_global=rg)
    if energy:
        out.update(f_site=R.f_site(), f_edge=R.f_edge(), f_non_edge=R.f_non_edge(), overlap=R.overlap())
        nae, nnae, cabe = R.em_stats()
        out.update(na_expect=nae, nna_expect=nnae, cab_expect=cabe)
        if entropy and dc == 0:
            out.update(entropy=R.entropy(), entropy_site=R.entropy_site(), entropy_edge=R.entropy_edge(),
                       entropy_non_edge=R.entropy_non_edge())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "M", R.M, "maxdeg", R.max_degree, "maxdiff", maxdiff)


def converge_case(name, u, v, sizes, dc, seed, eps_c=None, pa=None, cab_upper=None, crit=5e-6, tmax=1000):
    R = Reference(u, v, sizes, dc)
    R.init_messages(seed, 1.0)
    if eps_c is not None:
        R.set_params_epsilon_c(*eps_c)
    else:
        R.set_params_direct(pa, cab_upper)
    na, cab, eta = R.get_params()
    niter = R.converge(crit, tmax, 1.0)
    msg, marg, h = R.get_state()
    out = dict(u=u, v=v, sizes=np.asarray(sizes, np.uint32), dc=dc, seed=seed, na=na, cab=cab, crit=crit, tmax=tmax,
               niter=niter, marg=marg, f=R.free_energy(), overlap=R.overlap())
    if dc == 0:
        out.update(entropy=R.entropy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "niter", niter, "f", out["f"], "overlap", out["overlap"])


def learn_case(name, u, v, sizes, dc, seed, pa, cab_upper, crit=1e-6, tmax=1000, lr=0.2):
    R = Reference(u, v, sizes, dc, learn_mode=True)
    R.init_messages(seed, 1.0)
    R.set_params_direct(pa, cab_upper)
    na0, cab0, _ = R.get_params()
    na, cab, eta = R.learning(crit, tmax, lr, 1.0)
    out = dict(u=u, v=v, sizes=np.asarray(sizes, np.uint32), dc=dc, seed=seed, na0=na0, cab0=cab0, crit=crit,
               tmax=tmax, lr=lr, na=na, cab=cab, eta=eta, overlap=R.overlap())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "eta", eta, "cab", cab.ravel())


def init_case(name, u, v, sizes, flag, conf, seed, learn_mode, eps_c=None, pa=None, cab_upper=None, converge=False):
    """init_messages flags 1-3 (belief_propagation.cpp:132-215) from a beliefs vector, then one synchronous sweep by
    bp_conditional (infer: planted nodes frozen, :1100-1126) or bp_basic (learn); optionally the reference's converge()."""
    R = Reference(u, v, sizes, 0, learn_mode=learn_mode)
    assert R.init_messages_flag(flag, conf, seed, 1.0) == 0
    if eps_c is not None:
        R.set_params_epsilon_c(*eps_c)
    else:
        R.set_params_direct(pa, cab_upper)
    R.init_h()
    na, cab, eta = R.get_params()
    msg0, marg0, h0 = R.get_state()
    new_msg, new_marg, node_diff, maxdiff = R.jacobi_sweep(1.0)
    out = dict(u=u, v=v, sizes=np.asarray(sizes, np.uint32), dc=0, seed=seed, flag=flag, conf=np.asarray(conf, np.int32),
               learn_mode=int(learn_mode), na=na, cab=cab, msg0=msg0, marg0=marg0, h0=h0, new_msg=new_msg,
               new_marg=new_marg, node_diff=node_diff, maxdiff=maxdiff)
    if converge:
        niter = R.converge(5e-6, 1000, 1.0)
        out.update(niter=niter, marg=R.get_state()[1], overlap=R.overlap())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "flag", flag, "planted", int((np.asarray(conf) != -1).sum()), "maxdiff", maxdiff,
          ("niter %d overlap %.4f" % (out["niter"], out["overlap"])) if converge else "")


def mb_rand_case(name, u, v, sizes, seed, eps_c):
    """--mb_rand (main.cpp:299-301): std::shuffle of the memberships advances the engine before init_messages; the
    initial state and the reference's own converge() from it."""
    R = Reference(u, v, sizes, 0)
    R.init_messages_mb_rand(seed, 1.0)
    R.set_params_epsilon_c(*eps_c)
    na, cab, eta = R.get_params()
    msg0, marg0, _ = R.get_state()
    niter = R.converge(5e-6, 1000, 1.0)
    out = dict(u=u, v=v, sizes=np.asarray(sizes, np.uint32), dc=0, seed=seed, na=na, cab=cab, msg0=msg0, marg0=marg0,
               niter=niter, marg=R.get_state()[1], f=R.free_energy(), overlap=R.overlap())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "niter", niter, "f", out["f"], "overlap", out["overlap"])


def hub_graph(N, Q, seed, hub_degree):
    """power-law DC-SBM plus one node wired to hub_degree others: exercises the >= 50 and the one-CTA-per-hub paths"""
    u, v, sizes, theta = generators.dc_sbm_powerlaw(N, Q, gamma=2.5, k_min=2.0, ratio=10.0, seed=seed)
    rng = np.random.default_rng(seed + 100)
    others = rng.choice(np.arange(1, N), size=hub_degree, replace=False).astype(np.uint32)
    u = np.concatenate([u, np.zeros(hub_degree, np.uint32)])
    v = np.concatenate([v, others])
    return u, v, sizes


def main():
    u, v = Reference.load_edge_list(SHIPPED)
    sizes = [500, 500]
    # ---- level 1: one sweep from a frozen state, shipped graph (BASELINE config #1)
    sweep_case("sweep_cfg1_readme", u, v, sizes, 0, 0, pa=[.5, .5], cab_upper=[3.63, 2.36, 3.63])
    sweep_case("sweep_cfg1_eps01", u, v, sizes, 0, 1, eps_c=(0.1, 3.0))
    sweep_case("sweep_cfg1_dc1", u, v, sizes, 1, 2, pa=[.5, .5], cab_upper=[0.6, 0.06, 0.6])
    sweep_case("sweep_cfg1_dc2", u, v, sizes, 2, 3, pa=[.5, .5], cab_upper=[0.6, 0.06, 0.6])
    sweep_case("sweep_cfg1_beta_damp", u, v, sizes, 0, 4, beta=0.8, damping=0.5, eps_c=(0.1, 3.0))
    # unequal blocks read as Q = 4 on the same graph (parameters need not be the generating ones)
    sweep_case("sweep_cfg1_q4", u, v, [200, 300, 250, 250], 0, 5, pa=[.2, .3, .25, .25],
               cab_upper=[6, 1, .5, .7, 5, .8, 1.2, 7, .9, 4])
    sweep_case("sweep_cfg1_q3", u, v, [300, 300, 400], 0, 6, pa=[.3, .3, .4], cab_upper=[6, 1, .5, 5, .8, 7])
    # ---- hubs: degree >= 50 (log-domain routine) and degree > one tile
    hu, hv, hs = hub_graph(3000, 2, 11, 1500)
    sweep_case("sweep_hub_q2", hu, hv, hs, 0, 7, pa=[.5, .5], cab_upper=[8, 1, 8], entropy=False)
    sweep_case("sweep_hub_q2_dc1", hu, hv, hs, 1, 8, pa=[.5, .5], cab_upper=[0.5, 0.05, 0.5], entropy=False)
    hu4, hv4, hs4 = hub_graph(3000, 4, 12, 700)
    sweep_case("sweep_hub_q4_dc1", hu4, hv4, hs4, 1, 9, pa=[.25] * 4,
               cab_upper=[.5, .05, .05, .05, .5, .05, .05, .5, .05, .5], entropy=False)
    # ---- level 2: the reference's own converge / learning end states
    converge_case("converge_cfg1_readme", u, v, sizes, 0, 0, pa=[.5, .5], cab_upper=[3.63, 2.36, 3.63])
    converge_case("converge_cfg1_eps01", u, v, sizes, 0, 0, eps_c=(0.1, 3.0))
    converge_case("converge_cfg1_eps01_dc1", u, v, sizes, 1, 0, pa=[.5, .5], cab_upper=[0.6, 0.06, 0.6])
    su, sv, ss, supper = generators.planted_sbm_epsilon_c(6000, 3, 0.15, 6.0, seed=3)
    converge_case("converge_sbm_q3", su, sv, ss, 0, 1, pa=[1 / 3.] * 3, cab_upper=supper)
    learn_case("learn_cfg1_515", u, v, sizes, 0, 0, [.5, .5], [5, 1, 5])
    lu, lv, ls, lupper = generators.planted_sbm_epsilon_c(8000, 2, 0.2, 4.0, seed=5)
    learn_case("learn_sbm_n8000", lu, lv, ls, 0, 0, [.5, .5], [5, 2, 5])


def init_cases():
    """planted / noisy / fixed initialisations and bp_conditional clamping, shipped graph."""
    u, v = Reference.load_edge_list(SHIPPED)
    sizes = [500, 500]
    truth = np.repeat(np.arange(2), 500).astype(np.int32)
    rng = np.random.default_rng(42)
    partial = np.where(rng.random(1000) < 0.7, -1, truth).astype(np.int32)  # 30 % of the labels known
    init_case("init_flag1_partial_infer", u, v, sizes, 1, partial, 3, False, eps_c=(0.1, 3.0), converge=True)
    init_case("init_flag1_partial_learn", u, v, sizes, 1, partial, 3, True, eps_c=(0.1, 3.0))
    init_case("init_flag1_full_infer", u, v, sizes, 1, truth, 4, False, eps_c=(0.1, 3.0))
    # flags 2 and 3 assert(conf != 1) in the reference: read the two blocks as labels 0 and 2 of a Q = 3 model
    conf02 = np.where(truth == 1, 2, truth).astype(np.int32)
    s3 = [500, 0, 500]
    init_case("init_flag2_full_infer", u, v, s3, 2, conf02, 5, False, pa=[.5, 0.0, .5], cab_upper=[5, 1, 1, 5, 1, 5])
    init_case("init_flag2_full_learn", u, v, s3, 2, conf02, 5, True, pa=[.5, 0.0, .5], cab_upper=[5, 1, 1, 5, 1, 5])
    init_case("init_flag3_full_infer", u, v, s3, 3, conf02, 6, False, pa=[.5, 0.0, .5], cab_upper=[5, 1, 1, 5, 1, 5])
    # --mb_rand: the shuffle's draws come before init_messages'
    mb_rand_case("mbrand_cfg1_eps01", u, v, sizes, 0, (0.1, 3.0))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "init":
        init_cases()
        sys.exit(0)
    main()
