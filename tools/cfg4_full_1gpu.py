"""BASELINE configs[3] AS STATED on ONE B200: planted SBM N = 100M, Q = 2, c = 10 (1.0e9 directed edges; 16 GB per FP64
message buffer), -m infer.  The strong-scaling anchor of the multi-GPU curve: the same graph size the 8-GPU run
partitions.  Device-side initial messages; times the sweep kernel alone (events inside the library) and a converge()
to the reference's default criterion; checks the size-independent properties.  One JSON line into gpurun_out/.

    python tools/cfg4_full_1gpu.py [N]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sbm_bp_b200 import api, generators  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
t0 = time.time()
u, v, sizes, upper = generators.planted_sbm_epsilon_c(N, 2, 0.1, 10.0, seed=1)
t1 = time.time()
bm = api.blockmodel_t(sizes, (u, v), 0)
del u, v
t2 = time.time()
M = bm.get_M()
print("generated %.1f s, graph built %.1f s: N=%d M=%d" % (t1 - t0, t2 - t1, N, M), flush=True)
bp = api.belief_propagation(bm, "f64")
t3 = time.time()
print("engine created %.1f s, kernel %s" % (t3 - t2, bp.sweep_kernel_name()), flush=True)
bp.expand_bp_params(api.bp_param_from_direct(bm, [0.5, 0.5], upper))
bp.init_messages_device(1234)
for _ in range(2):
    bp.time_sweep_kernel()
kms = [bp.time_sweep_kernel() for _ in range(5)]
t = time.perf_counter()
niter = bp.converge(5e-6, 300, 1.0)
conv_s = time.perf_counter() - t
s = bp.stats()
marg = bp.get_marginals()
norm_err = float(np.max(np.abs(marg.sum(axis=1) - 1.0)))
overlap = bp.compute_overlap()
B, kernel_ms = s["bytes_per_edge"], float(np.mean(kms))
out = {"workload": "BASELINE configs[3] at full size on ONE GPU: planted SBM N=%d, Q=2, c=10, eps=0.1, -m infer" % N,
       "N": N, "M": int(M), "kernel": bp.sweep_kernel_name(), "kernel_ms": kernel_ms, "value_edge_updates_per_s": M / (kernel_ms * 1e-3),
       "bytes_per_edge_update": B, "frac": M * B / (kernel_ms * 1e-3) / 1e9 / peak, "converge_niter": niter,
       "time_to_converge_s": conv_s, "overlap": overlap, "marginal_norm_err": norm_err,
       "host_seconds": {"generate": t1 - t0, "graph": t2 - t1, "engine": t3 - t2}, "data": "synthetic, device-side initial messages"}
print(json.dumps(out), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "cfg4_full_1gpu.json"), "w") as fh:
    fh.write(json.dumps(out) + "\n")
assert norm_err < 1e-9 and niter >= 0
