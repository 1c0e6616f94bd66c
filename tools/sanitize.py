"""Small runs of the newest kernels for compute-sanitizer (memcheck / racecheck / initcheck), no torch needed:

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py

(compute-sanitizer is closed on the round-1 GPU pool, so this has only run plain there -- it doubles as a quick smoke of
the same paths.)  Covers the replay kernel (product and log-domain routines, dc 0 and 1), the degree-class kernel under programmatic
dependent launch with and without the lazy sweep close, and the reductions that follow a converge().
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sbm_bp_b200 import api, generators  # noqa: E402


def replay(dc):
    rng = np.random.default_rng(1)
    sizes = [200, 200]
    cab = np.array([[8.0, 1.0], [1.0, 8.0]])
    u, v = generators.planted_sbm(sizes, cab, seed=2)
    others = rng.choice(np.arange(1, 400), size=80, replace=False).astype(np.uint32)
    u = np.concatenate([u, np.zeros(80, np.uint32)])
    v = np.concatenate([v, others])
    bm = api.blockmodel_t(sizes, (u, v), dc)
    bp = api.belief_propagation(bm, "f64")
    bp.init_messages(3)
    c = cab / (60.0 if dc else 1.0)
    bp.expand_bp_params(api.bp_param_from_direct(bm, [.5, .5], [c[0, 0], c[0, 1], c[1, 1]]))
    bp.set_schedule("replay")
    it = bp.converge(5e-6, 3, 1.0)
    print("replay dc", dc, "niter", it, "f", bp.compute_free_energy())


def ell(lazy):
    os.environ["SBMBP_LAZY_CLOSE"] = "1" if lazy else "0"
    u, v, sizes, upper = generators.planted_sbm_epsilon_c(5000, 2, 0.1, 3.0, seed=4)
    bm = api.blockmodel_t(sizes, (u, v))
    bp = api.belief_propagation(bm, "f64")
    bp.init_messages(9)
    bp.expand_bp_params(api.bp_param_from_direct(bm, [.5, .5], upper))
    it = bp.converge(5e-6, 200, 1.0)
    print("ell lazy", lazy, bp.sweep_kernel_name(), "niter", it, "f", bp.compute_free_energy(), "overlap", bp.compute_overlap())


if __name__ == "__main__":
    replay(0)
    replay(1)
    ell(False)
    ell(True)
