"""Trajectory-level parity at size: the compiled reference's converge() (random-sequential, std::mt19937) against the
engine's replay schedule on the BASELINE configs[1] family -- same seed, same initial messages, same draws.  Reports both
sweep counts, the largest marginal difference and the time per draw of the one-warp replay kernel.

    python tools/replay_at_size.py [N] [seed]          (needs oracle/_ref/libsbmbp_ref.so and a B200)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference  # noqa: E402  (checker, not product)
from sbm_bp_b200 import api, generators  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
u, v, sizes, upper = generators.planted_sbm_epsilon_c(N, 2, 0.1, 3.0, seed=1)
t = time.time()
R = Reference(u, v, sizes, 0)
R.init_messages(seed)
R.set_params_direct([.5, .5], upper)
it_ref, ref_s = R.converge_timed(5e-6, 1000, 1.0)
marg_ref = R.get_state()[1]
print("reference: niter %d, converge %.1f s (setup + converge %.1f s)" % (it_ref, ref_s, time.time() - t), flush=True)
bm = api.blockmodel_t(sizes, (u, v))
bp = api.belief_propagation(bm, "f64")
bp.init_messages(seed)
bp.expand_bp_params(api.bp_param_from_direct(bm, [.5, .5], upper))
bp.set_schedule("replay")
t = time.time()
it = bp.converge(5e-6, 1000, 1.0)
dt = time.time() - t
marg = bp.get_marginals()
out = {"workload": "BASELINE configs[1] family: planted SBM N=%d, Q=2, c=3, eps=0.1" % N, "seed": seed,
       "niter_reference": int(it_ref), "niter_replay": int(it), "max_abs_marginal_diff": float(np.max(np.abs(marg - marg_ref))),
       "replay_seconds": dt, "us_per_draw": 1e6 * dt / (max(it, 0) + 1) / N, "reference_converge_seconds": ref_s,
       "overlap": bp.compute_overlap()}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "replay_at_size_%d.json" % N), "w") as fh:
    fh.write(json.dumps(out) + "\n")
assert it == it_ref
