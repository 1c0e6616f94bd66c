"""BASELINE configs[4] at FULL size on one B200: assortative SBM N = 10M, Q = 32, c = 16 (M = 1.6e8 directed edges,
41 GB per message buffer in FP64), -m infer.  Device-side initial messages (the host state of this size is 41 GB; the
e2e leg of bench.py is for configs[1]); times the sweep kernel alone (CUDA events inside the library), a batch of
sweeps, converge() to the reference's default criterion, and checks the size-independent properties (marginals
normalised, overlap with the planted partition).  One JSON line per precision into gpurun_out/.

    python tools/cfg5_full.py [N] [precisions: f64,f32]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sbm_bp_b200 import api, generators  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
precs = (sys.argv[2] if len(sys.argv) > 2 else "f64,f32").split(",")
Q, eps, c = 32, 0.1, 16.0
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
t0 = time.time()
u, v, sizes, upper = generators.planted_sbm_epsilon_c(N, Q, eps, c, seed=1)
t1 = time.time()
bm = api.blockmodel_t(sizes, (u, v), 0)
t2 = time.time()
M = bm.get_M()
print("generated %.1f s, graph built %.1f s: N=%d M=%d max degree %d" % (t1 - t0, t2 - t1, N, M, bm.get_graph_max_degree()), flush=True)
del u, v
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for prec in precs:
    bp = api.belief_propagation(bm, prec)
    bp.expand_bp_params(api.bp_param_from_direct(bm, [1.0 / Q] * Q, upper))
    bp.init_messages_device(1234)
    for _ in range(2):
        bp.sweeps_async(1)
    bp.sync()
    kms = [bp.time_sweep_kernel() for _ in range(5)]
    t = time.perf_counter()
    bp.sweeps_async(5)
    bp.sync()
    batch_ms = (time.perf_counter() - t) * 1e3 / 5
    s0 = bp.stats()
    t = time.perf_counter()
    niter = bp.converge(5e-6, 300, 1.0)
    conv_s = time.perf_counter() - t
    s1 = bp.stats()
    sweeps = s1["sweeps"] - s0["sweeps"]
    marg = bp.get_marginals()
    norm_err = float(np.max(np.abs(marg.sum(axis=1) - 1.0)))
    finite = bool(np.isfinite(marg).all())
    overlap = bp.compute_overlap()
    del marg
    bpe = s1["bytes_per_edge"]
    kernel_ms = float(np.mean(kms))
    out = {"workload": "BASELINE configs[4] at full size: assortative SBM N=%d, Q=32, c=16, eps=0.1, -m infer" % N,
           "precision": prec, "N": N, "M": int(M), "kernel": bp.sweep_kernel_name(), "kernel_ms": kernel_ms,
           "ms_per_sweep_in_batch": batch_ms, "value_edge_updates_per_s": M / (batch_ms * 1e-3),
           "bytes_per_edge_update": bpe, "achieved_gbs": M * bpe / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peak,
           "frac": M * bpe / (kernel_ms * 1e-3) / 1e9 / peak, "converge_niter": niter, "converge_sweeps": int(sweeps),
           "time_to_converge_s": conv_s, "overlap": overlap, "marginal_norm_err": norm_err, "finite": finite,
           "data": "synthetic, device-side initial messages"}
    print(json.dumps(out), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "cfg5_full_%s.json" % prec), "w") as fh:
        fh.write(json.dumps(out) + "\n")
    assert finite and norm_err < 1e-9
    bp.close()
