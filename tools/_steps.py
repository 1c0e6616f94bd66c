import sys, numpy as np, torch
sys.path.insert(0, '.')
from sbm_bp_b200 import api, generators
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12500000
u, v, sizes, upper = generators.planted_sbm_epsilon_c(N, 2, 0.1, 10.0, seed=1)
bm = api.blockmodel_t(sizes, (u, v))
bp = api.belief_propagation(bm, "f64")
bp.expand_bp_params(api.bp_param_from_direct(bm, [.5, .5], upper))
bp.init_messages_device(1234)
print(bp.sweep_kernel_name())
ts = []
for k in range(40):
    ts.append(bp.time_sweep_kernel())
print(" ".join("%.2f" % t for t in ts))
print("tiny", bp.tiny_events(), "overlap", bp.compute_overlap())
