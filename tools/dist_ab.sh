#!/bin/bash
# A/B of the multi-GPU sweep on N GPUs: tools/dist_ab.sh <ngpus> "<name>=<ENV=val,...> ..." [extra bench args]
# (the library must have been built with EXTRA_NVFLAGS=-DSBMBP_TUNING for the SBMBP_DIST_DBG switches)
N=${1:-2}
VARS=${2:-default=}
shift 2
mkdir -p gpurun_out
for spec in $VARS; do
  var=${spec%%=*}
  envs=${spec#*=}
  (
    IFS=','
    for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    unset IFS
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --e2e-steps 1 "$@" > gpurun_out/dist_${N}_${var}.json 2> gpurun_out/dist_${N}_${var}.err
    echo "$var rc $?"
  )
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dist_${N}_${var}.json").read().strip().splitlines()[-1])
    print("  value %.3e  ms/step %.3f  frac/GPU %.3f  e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as ex:
    print("  parse failed", ex)
PY
done
