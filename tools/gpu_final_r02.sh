#!/bin/bash
# Round-2 measurement pass on one B200 (gpurun -- 'bash tools/gpu_final_r02.sh'): tests, bench lines, launch list, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02_final.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02_final.log
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r02_n1_cfg2_f64.json 2> gpurun_out/bench_r02_n1_cfg2_f64.err; echo "bench f64 rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_n1_reference.json 2> /dev/null; echo "ref rc $?"
python bench.py --steps 100 --warmup 10 --precision f32 --no-cpu-baseline --no-anchor > gpurun_out/bench_r02_n1_cfg2_f32.json 2> /dev/null; echo "bench f32 rc $?"
SBMBP_COMPACT=0 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-anchor > gpurun_out/bench_r02_n1_cfg2_f64_fullstorage.json 2> /dev/null; echo "bench full rc $?"
python bench.py --workload cfg5small --steps 10 --warmup 3 --no-anchor --cpu-sweeps 2 > gpurun_out/bench_r02_cfg5small_f64.json 2> /dev/null; echo "cfg5small rc $?"
python bench.py --workload cfg5small --steps 10 --warmup 3 --precision f32 --no-cpu-baseline --no-anchor > gpurun_out/bench_r02_cfg5small_f32.json 2> /dev/null; echo "cfg5small f32 rc $?"
# launch list of the bench command, then one full capture of the dominant kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_cfg2_f64.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-anchor --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc $?"
ncu --set full --clock-control none --import-source on -k regex:bp_sweep_ell -s 12 -c 1 -o gpurun_out/prof_ell_cfg2_f64_r02 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-anchor --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
for f in gpurun_out/bench_r02_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value %.3e e2e %.3e kernel_ms %s frac %s" % (d.get("value", 0), d.get("e2e", {}).get("value", 0), r.get("kernel_ms"), r.get("frac")))
except Exception as ex:
    print(sys.argv[1], "parse failed", ex)
PY
done
