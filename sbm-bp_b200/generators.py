"""Synthetic inputs of the BASELINE shapes: planted SBM and degree-corrected SBM edge lists.

Node ids are block-contiguous, so ``-n n_0 n_1 ...`` is the planted truth (main.cpp:239-252).  Pairs are
returned as two uint32 arrays in shuffled order; duplicates and self-loops are left to the loader exactly as
they would be in an edge-list file (duplicates merge, self-loops are dropped here because the model is simple).
"""
import numpy as np


def epsilon_c_to_cab(Q, epsilon, c):
    """cin, cout of bp_param_from_epsilon_c (blockmodel.cpp:256-257)."""
    cin = c * Q / ((Q - 1) * epsilon + 1)
    return cin, epsilon * cin


def planted_sbm(block_sizes, cab, seed=1):
    """Planted SBM: for every block pair (a <= b) a Poisson number of edges with mean n_a n_b c_ab / N
    (n_a^2 c_aa / 2N inside a block), endpoints uniform in the blocks."""
    rng = np.random.default_rng(seed)
    n = np.asarray(block_sizes, dtype=np.int64)
    Q = len(n)
    N = int(n.sum())
    cab = np.asarray(cab, dtype=np.float64).reshape(Q, Q)
    start = np.concatenate([[0], np.cumsum(n)])
    us, vs = [], []
    for a in range(Q):
        for b in range(a, Q):
            mean = n[a] * n[b] * cab[a, b] / N * (0.5 if a == b else 1.0)
            m = int(rng.poisson(mean))
            if m == 0:
                continue
            u = rng.integers(start[a], start[a + 1], size=m, dtype=np.int64)
            v = rng.integers(start[b], start[b + 1], size=m, dtype=np.int64)
            keep = u != v
            us.append(u[keep])
            vs.append(v[keep])
    if not us:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    u = np.concatenate(us)
    v = np.concatenate(vs)
    perm = rng.permutation(len(u))
    return u[perm].astype(np.uint32), v[perm].astype(np.uint32)


def planted_sbm_epsilon_c(N, Q, epsilon, c, seed=1):
    """Equal blocks, cin/cout from (epsilon, c).  Returns (u, v, block_sizes, cab_upper) -- cab_upper in --cab order."""
    sizes = [N // Q] * Q
    sizes[-1] += N - sum(sizes)
    cin, cout = epsilon_c_to_cab(Q, epsilon, c)
    cab = np.full((Q, Q), cout)
    np.fill_diagonal(cab, cin)
    u, v = planted_sbm(sizes, cab, seed)
    upper = [cab[a, b] for a in range(Q) for b in range(a, Q)]
    return u, v, sizes, upper


def dc_sbm_powerlaw(N, Q, gamma=2.5, k_min=2.0, ratio=10.0, seed=1, theta_cap=None):
    """Degree-corrected SBM (Chung-Lu within/between blocks): expected degrees theta_i ~ Pareto(gamma) above k_min,
    block affinity omega = ratio on the diagonal and 1 off it.  Returns (u, v, block_sizes, theta)."""
    rng = np.random.default_rng(seed)
    sizes = [N // Q] * Q
    sizes[-1] += N - sum(sizes)
    theta = k_min * (1.0 - rng.random(N)) ** (-1.0 / (gamma - 1.0))
    if theta_cap is None:
        theta_cap = float(np.sqrt(N * theta.mean()))  # structural cut-off of a simple graph
    theta = np.minimum(theta, theta_cap)
    start = np.concatenate([[0], np.cumsum(sizes)])
    D = np.array([theta[start[a]:start[a + 1]].sum() for a in range(Q)])
    omega = np.ones((Q, Q))
    np.fill_diagonal(omega, ratio)
    w = np.outer(D, D) * omega
    # expected number of edges between a and b (a < b), and inside a (a == b)
    tot = theta.sum() / 2.0
    pair_w = np.triu(w, 1).sum() + 0.5 * np.trace(w)
    us, vs = [], []
    cdf = [np.cumsum(theta[start[a]:start[a + 1]]) for a in range(Q)]
    for a in range(Q):
        for b in range(a, Q):
            mean = tot * (w[a, b] * (0.5 if a == b else 1.0)) / pair_w
            m = int(rng.poisson(mean))
            if m == 0:
                continue
            u = start[a] + np.searchsorted(cdf[a], rng.random(m) * cdf[a][-1], side="right")
            v = start[b] + np.searchsorted(cdf[b], rng.random(m) * cdf[b][-1], side="right")
            u = np.minimum(u, start[a + 1] - 1)
            v = np.minimum(v, start[b + 1] - 1)
            keep = u != v
            us.append(u[keep])
            vs.append(v[keep])
    u = np.concatenate(us)
    v = np.concatenate(vs)
    perm = rng.permutation(len(u))
    return u[perm].astype(np.uint32), v[perm].astype(np.uint32), sizes, theta


def write_edgelist(path, u, v):
    with open(path, "w") as f:
        for a, b in zip(u.tolist(), v.tolist()):
            f.write("%d %d\n" % (a, b))


def rank_ranges(N, world):
    """Contiguous node ranges of the multi-GPU partition: world + 1 boundaries."""
    return np.array([(N * k) // world for k in range(world + 1)], dtype=np.uint32)


def planted_sbm_rank(N, Q, epsilon, c, rank, world, seed=1):
    """The part of a planted SBM (equal blocks, block-contiguous ids) that rank `rank` of `world` needs: every edge
    with an endpoint in its node range.  Edges between two ranks' ranges are generated from a seed that depends only
    on the pair of ranks, so both owners produce the identical list without communicating.
    Returns (u, v, block_sizes, cab_upper, range_starts)."""
    sizes = [N // Q] * Q
    sizes[-1] += N - sum(sizes)
    gstart = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cin, cout = epsilon_c_to_cab(Q, epsilon, c)
    starts = rank_ranges(N, world).astype(np.int64)

    def segments(p):  # (lo, hi, group) pieces of rank p's range
        out = []
        for g in range(Q):
            lo, hi = max(starts[p], gstart[g]), min(starts[p + 1], gstart[g + 1])
            if lo < hi:
                out.append((int(lo), int(hi), g))
        return out

    us, vs = [], []
    for other in range(world):
        p, q = min(rank, other), max(rank, other)
        for ia, (alo, ahi, ga) in enumerate(segments(p)):
            for ib, (blo, bhi, gb) in enumerate(segments(q)):
                if p == q and ib < ia:
                    continue
                rng = np.random.default_rng([seed, p, q, ia, ib])
                same = p == q and ia == ib
                rate = (cin if ga == gb else cout) / N
                mean = (ahi - alo) * (bhi - blo) * rate * (0.5 if same else 1.0)
                m = int(rng.poisson(mean))
                if m == 0:
                    continue
                u = rng.integers(alo, ahi, size=m, dtype=np.int64)
                v = rng.integers(blo, bhi, size=m, dtype=np.int64)
                keep = u != v
                us.append(u[keep])
                vs.append(v[keep])
    u = np.concatenate(us).astype(np.uint32) if us else np.zeros(0, np.uint32)
    v = np.concatenate(vs).astype(np.uint32) if vs else np.zeros(0, np.uint32)
    upper = [cin if a == b else cout for a in range(Q) for b in range(a, Q)]
    return u, v, sizes, upper, starts.astype(np.uint32)
