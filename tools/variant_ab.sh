#!/bin/bash
# A/B of comparison builds (tools/build_variants.sh) on one bench workload: tools/variant_ab.sh <workload> <name> [<name> ...]
# ("default" = the in-tree library).  Prints the sweep kernel's time per variant; full lines land in gpurun_out/ab_<workload>_<name>.json.
w=$1; shift
mkdir -p gpurun_out
for name in "$@"; do
    lib=sbm-bp_b200/libsbmbp.so
    [ "$name" != default ] && lib=build_variants/libsbmbp_${name}.so
    SBMBP_LIB=$PWD/$lib timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
        > gpurun_out/ab_${w}_${name}.json 2> gpurun_out/ab_${w}_${name}.err || tail -3 gpurun_out/ab_${w}_${name}.err
    python - "$name" gpurun_out/ab_${w}_${name}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("%-10s kernel_ms %.4f step_ms %.4f frac %.3f  %s" % (sys.argv[1], d["roofline"]["kernel_ms"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"]))
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
done
