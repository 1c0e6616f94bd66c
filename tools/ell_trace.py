"""Tuning aid: per-warp timeline of bp_sweep_ell_kernel on BASELINE configs[1] (needs a B200).
Run with SBMBP_ELL_TRACE=1; prints where the sweep's time goes (prologue, chunks, epilogue)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["SBMBP_ELL_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sbm_bp_b200 import api, generators

prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
N = 1000000
cin, cout = 3.0 * 2 / 1.1, 0.1 * 3.0 * 2 / 1.1
u, v = generators.planted_sbm([N // 2, N // 2], np.array([[cin, cout], [cout, cin]]), seed=1)
bm = api.blockmodel_t([N // 2, N // 2], (u, v))
bp = api.belief_propagation(bm, prec)
bp.init_messages_device(0) if hasattr(bp, "init_messages_device") else bp.init_messages(0)
bp.expand_bp_params(api.bp_param_from_direct(bm, [0.5, 0.5], [cin, cout, cin]))
import torch

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(6):
    flush.fill_(it)
    torch.cuda.synchronize()
    ms = bp.time_sweep_kernel()
print("kernel_ms of the traced sweep:", ms)
n = C.c_uint64(0)
api._check(api.lib().sbmbp_debug_trace(bp._e, None, C.c_uint64(0), C.byref(n)))
buf = np.zeros(n.value, np.uint64)
api._check(api.lib().sbmbp_debug_trace(bp._e, buf.ctypes.data_as(C.c_void_p), C.c_uint64(n.value), C.byref(n)))
t = buf.reshape(-1, 16).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0  # us
rel[t == 0] = np.nan
print("warps traced:", len(t))
names = ["entry", "work start"] + ["chunk %d done" % i for i in range(12)] + ["work end", "exit"]
for k, nm in enumerate(names):
    col = rel[:, k]
    if np.all(np.isnan(col)):
        continue
    print("%-14s min %7.2f  median %7.2f  p90 %7.2f  max %7.2f us" % (nm, np.nanmin(col), np.nanmedian(col), np.nanpercentile(col, 90), np.nanmax(col)))
d = np.diff(rel[:, 1:14], axis=1)
print("per-chunk time (us): median %.2f  p10 %.2f  p90 %.2f" % (np.nanmedian(d), np.nanpercentile(d, 10), np.nanpercentile(d, 90)))
print("first chunk (us): median %.2f" % np.nanmedian(rel[:, 2] - rel[:, 1]))
