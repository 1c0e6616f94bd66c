import sys, time; sys.path.insert(0, ".")
import numpy as np
from sbm_bp_b200 import api, generators
# BASELINE configs[2] shape at reduced size: DC-SBM, Q=4, power-law degrees gamma=2.5, --deg_corr_flag 1, -m learn
N, Q = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000000), 4
t0 = time.time(); u, v, sizes, theta = generators.dc_sbm_powerlaw(N, Q, gamma=2.5, k_min=2.0, ratio=10.0, seed=1); t1 = time.time()
bm = api.blockmodel_t(sizes, (u, v), 1); t2 = time.time()
rp, col, rev, deg = bm.csr()
print("gen %.1fs build %.1fs N=%d M=%d maxdeg=%d nodes>=50: %d" % (t1 - t0, t2 - t1, N, bm.get_M(), bm.get_graph_max_degree(), int((deg >= 50).sum())))
# planted parameters in the dc parametrisation: c_ab = N m_ab / (D_a D_b), diagonal 2 N m_aa / D_a^2  (belief_propagation.cpp:974-983)
grp = np.repeat(np.arange(Q), sizes)
D = np.array([deg[grp == a].sum() for a in range(Q)], float)
gu, gv = grp[u], grp[v]
m = np.zeros((Q, Q));
for a in range(Q):
    for b in range(Q):
        m[a, b] = np.sum((gu == a) & (gv == b)) + np.sum((gu == b) & (gv == a))
m[np.diag_indices(Q)] /= 2.0  # m[a, b] = undirected edges between a and b; m[a, a] = edges inside a
cab = np.zeros((Q, Q))
for a in range(Q):
    for b in range(Q):
        cab[a, b] = N * m[a, b] / (D[a] * D[b]) * (2.0 if a == b else 1.0) if a == b else N * (m[a, b]) / (D[a] * D[b])
print("planted cab diag", np.diag(cab), "off", cab[0, 1])
start = cab * (1 + 0.3 * (np.random.default_rng(0).random((Q, Q)) - 0.5)); start = (start + start.T) / 2
for prec in (("f64",) if N > 2000000 else ("f64", "f32")):
    bp = api.belief_propagation(bm, prec)
    bp.init_messages_device(3)
    st = api.bp_blockmodel_state(np.array(sizes, np.uint32), start)
    t0 = time.time()
    eta, cabl, na, iters = bp.learning(st, 1e-6, 100, 0.2, 1.0)
    dt = time.time() - t0
    s = bp.stats()
    print(prec, "learn: %.2fs, %d EM iterations, %d sweeps, %.3e edge-upd/s overall, overlap %.4f" % (dt, iters, s["sweeps"], s["edge_updates"] / dt, bp.compute_overlap()))
    print("  learned diag", np.diag(cabl), "off01", cabl[0, 1], "eta", eta)
