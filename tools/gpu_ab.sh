#!/bin/bash
# A/B of sweep-kernel variants on one B200: parity tests first (unless SKIP_TESTS=1), then short bench runs per variant.
# usage: tools/gpu_ab.sh "<workloads>" "<precisions>" "<variant>=<ENV=val,ENV=val> ..."
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
  echo "pytest rc $?" | tee -a gpurun_out/pytest_ab.log
  tail -3 gpurun_out/pytest_ab.log
fi
WLS=${1:-cfg2}
PRECS=${2:-f64 f32}
VARS=${3:-default=}
for wl in $WLS; do
  for prec in $PRECS; do
    for spec in $VARS; do
      var=${spec%%=*}
      envs=${spec#*=}
      (
        IFS=','
        for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
        unset IFS
        timeout 600 python bench.py --workload $wl --precision $prec --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline --e2e-steps 2 \
          > gpurun_out/ab_${wl}_${prec}_${var}.json 2> gpurun_out/ab_${wl}_${prec}_${var}.err
        echo "$wl $prec $var rc $?"
      )
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${wl}_${prec}_${var}.json").read().strip().splitlines()[-1])
    print("  value %.3e  ms/step %.4f  kernel_ms %.4f  frac %.3f  e2e %.3e warm %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d.get("warm_l2_value", 0)))
except Exception as ex:
    print("  parse failed", ex)
PY
    done
  done
done
