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


def balanced_ranges(deg, world, const=4.0):
    """Contiguous node ranges with about equal work sum_i (d_i + const) per rank: world + 1 boundaries.
    const stands for the per-node cost (row pointer, marginal, node phase) in units of one edge slot."""
    w = np.asarray(deg, np.float64) + float(const)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    starts = np.concatenate([[0], cuts, [len(w)]]).astype(np.int64)
    for k in range(1, world + 1):  # every rank keeps at least one node
        starts[k] = max(starts[k], starts[k - 1] + 1) if k < world else len(w)
    for k in range(world - 1, 0, -1):
        starts[k] = min(starts[k], starts[k + 1] - 1)
    return starts.astype(np.uint32)


class Partition:
    """Node partition of the multi-GPU engine as BASELINE.json states it: random relabelling, then contiguous ranges of the
    NEW ids balanced on sum (d_i + const).  new_id[i] = new id of original node i; old_id = its inverse; starts = world + 1
    range boundaries in new ids.  relabel_seed=None keeps the original ids (block-contiguous planted graphs then keep
    their locality: fewer messages cross ranks)."""

    def __init__(self, N, world, u=None, v=None, relabel_seed=None, balance=True, const=4.0):
        self.N, self.world = int(N), int(world)
        if relabel_seed is None:
            self.new_id = np.arange(N, dtype=np.uint32)
        else:
            self.new_id = np.random.default_rng(relabel_seed).permutation(N).astype(np.uint32)
        self.old_id = np.empty(N, np.uint32)
        self.old_id[self.new_id] = np.arange(N, dtype=np.uint32)
        if balance and u is not None:
            uu, vv = self.relabel(u, v)
            deg = np.bincount(uu, minlength=N) + np.bincount(vv, minlength=N)  # multi-edges counted: a weight, not a contract
            self.starts = balanced_ranges(deg, world, const)
        else:
            self.starts = rank_ranges(N, world)

    def relabel(self, u, v):
        return self.new_id[np.asarray(u, np.int64)], self.new_id[np.asarray(v, np.int64)]

    def owned(self, rank):
        """Original ids of the nodes rank owns, in the order of its rows."""
        return self.old_id[int(self.starts[rank]):int(self.starts[rank + 1])]

    def to_original(self, per_new_node):
        """Array indexed by new id -> indexed by original id."""
        return np.asarray(per_new_node)[self.new_id]


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


# ---- legacy MODE-NET side of the data path: its micro-canonical generator and its GML files, so
# graphs made for / by the legacy `sbm` binary can be fed to the engine.  Written from the behaviour of
# src/old/bm.cpp:41-170 (reader), :192-268 (generator), :270-296 (writer); nothing here is on the GPU path.

def microcanonical_sbm(block_sizes, cab, seed=1):
    """Legacy generator (old/bm.cpp:192-268): block-contiguous ids and an EXACT number of edges per block pair --
    int(p_ab n_a n_b) between blocks, int(p_aa n_a (n_a - 1) / 2) inside one, p_ab = c_ab / N -- drawn uniformly
    without self-loops and without duplicates (rejection, like the legacy do-while).  Returns (u, v)."""
    rng = np.random.default_rng(seed)
    n = np.asarray(block_sizes, dtype=np.int64)
    Q = len(n)
    N = int(n.sum())
    cab = np.asarray(cab, dtype=np.float64).reshape(Q, Q)
    start = np.concatenate([[0], np.cumsum(n)])
    us, vs = [], []
    for a in range(Q):
        for b in range(a, Q):
            p = cab[a, b] / N
            want = int(p * n[a] * n[b]) if a != b else int(p * n[a] * (n[a] - 1) / 2)
            cap = n[a] * n[b] if a != b else n[a] * (n[a] - 1) // 2
            if want > cap:
                raise ValueError("block pair (%d, %d): %d edges asked of %d possible" % (a, b, want, cap))
            keys = np.zeros(0, np.int64)
            while len(keys) < want:
                m = want - len(keys)
                x = rng.integers(start[a], start[a + 1], size=m + m // 8 + 8, dtype=np.int64)
                y = rng.integers(start[b], start[b + 1], size=m + m // 8 + 8, dtype=np.int64)
                ok = x != y
                lo, hi = np.minimum(x[ok], y[ok]), np.maximum(x[ok], y[ok])
                fresh = np.unique(np.concatenate([keys, lo * N + hi]))
                # keep the old ones and as many new ones as still needed (np.unique sorts: pick the new ones at random)
                new = np.setdiff1d(fresh, keys, assume_unique=True)
                if len(new) > m:
                    new = rng.choice(new, size=m, replace=False)
                keys = np.concatenate([keys, new])
            us.append(keys // N)
            vs.append(keys % N)
    u = np.concatenate(us) if us else np.zeros(0, np.int64)
    v = np.concatenate(vs) if vs else np.zeros(0, np.int64)
    perm = rng.permutation(len(u))
    return u[perm].astype(np.uint32), v[perm].astype(np.uint32)


def write_gml(path, u, v, labels):
    """The legacy writer's layout (old/bm.cpp:270-296): `graph [ directed 0  node [ id i  value g ] ...  edge [ source
    s  target t ] ... ]`, one token group per line."""
    with open(path, "w") as f:
        f.write("graph [\n  directed 0\n")
        for i, g in enumerate(np.asarray(labels).tolist()):
            f.write("  node\n  [\n    id %d\n    value %s\n  ]\n" % (i, g))
        for a, b in zip(np.asarray(u).tolist(), np.asarray(v).tolist()):
            f.write("  edge\n  [\n    source %d\n    target %d\n  ]\n" % (a, b))
        f.write("]\n")


def read_gml(path):
    """The legacy reader's semantics (old/bm.cpp:41-170), token by token: nodes are numbered in order of appearance of
    their `id` (ids are arbitrary strings), `value` strings become groups 0, 1, ... in order of first appearance, other
    node keys are skipped; an edge is the `source` / `target` pair after `edge [`; an edge already seen in either
    orientation is dropped.  Returns (u, v, labels, ids) with u, v indexing into ids; labels is -1 where a node has no
    value.  Malformed files raise ValueError where the legacy code asserts."""
    with open(path) as f:
        tok = f.read().split()
    id2idx, ids, value_of, colour = {}, [], {}, {}
    k = 0
    while k < len(tok):
        if tok[k] == "node":
            if k + 1 >= len(tok) or tok[k + 1] != "[":
                raise ValueError("[ should follow node")
            k += 2
            myid, closed = None, False
            for _ in range(100):  # the legacy reader scans at most 100 tokens of a node
                if k >= len(tok):
                    break
                t = tok[k]
                if t == "]":
                    closed = True
                    break
                if t == "id":
                    myid = tok[k + 1]
                    if myid in id2idx:
                        raise ValueError("multi-definition of node %s" % myid)
                    id2idx[myid] = len(ids)
                    ids.append(myid)
                    k += 2
                elif t == "value":
                    if myid is None:
                        raise ValueError("id should be given before value")
                    value_of[myid] = tok[k + 1]
                    colour.setdefault(tok[k + 1], len(colour))
                    k += 2
                else:
                    k += 1
            if not closed:
                raise ValueError("unterminated node section")
        k += 1
    us, vs, seen = [], [], set()
    k = 0
    while k < len(tok):
        if tok[k] == "edge":
            if k + 1 >= len(tok) or tok[k + 1] != "[":
                raise ValueError("[ should follow edge")
            k += 2
            while k < len(tok) and tok[k] != "source":
                k += 1
            if k + 3 >= len(tok) or tok[k + 2] != "target":
                raise ValueError("there should be target following source")
            s, t = tok[k + 1], tok[k + 3]
            if s not in id2idx or t not in id2idx:
                raise ValueError("edge endpoint does not exist in the node section")
            i, j = id2idx[s], id2idx[t]
            if (i, j) not in seen and (j, i) not in seen:
                seen.add((i, j))
                us.append(i)
                vs.append(j)
            k += 4
        else:
            k += 1
    labels = np.array([colour[value_of[x]] if x in value_of else -1 for x in ids], dtype=np.int32)
    return np.array(us, np.uint32), np.array(vs, np.uint32), labels, ids
