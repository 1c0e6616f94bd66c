#!/bin/bash
# A/B of the sweep kernels on one B200: parity tests first, then short bench runs per variant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
echo "pytest rc $?" | tee -a gpurun_out/pytest_ab.log
tail -5 gpurun_out/pytest_ab.log
for wl in cfg2; do
  for prec in f64 f32; do
    for var in ell ell24 ell1 ellnp; do
      unset SBMBP_NO_ELL SBMBP_WARP_MAIN SBMBP_REGION_MB; unset SBMBP_ELL_AHEAD_MB; case $var in ell24) export SBMBP_REGION_MB=24;; ell1) export SBMBP_REGION_MB=0;; ellnp) export SBMBP_ELL_AHEAD_MB=0;; esac
      timeout 600 python bench.py --workload $wl --precision $prec --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
        > gpurun_out/ab_${wl}_${prec}_${var}.json 2> gpurun_out/ab_${wl}_${prec}_${var}.err
      echo "$wl $prec $var rc $?"
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${wl}_${prec}_${var}.json").read().strip().splitlines()[-1])
    print("  value %.3e  ms/step %.4f  kernel_ms %.4f  frac %.3f  e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as ex:
    print("  parse failed", ex)
PY
    done
  done
done
