#!/bin/bash
# Round-end measurement on one B200: parity tests, the bench lines, the launch list and the ncu captures behind profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final_cfg2_f64.json 2> gpurun_out/bench_final_cfg2_f64.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "ref rc $?"
python bench.py --precision f32 --no-cpu-baseline > gpurun_out/bench_final_cfg2_f32.json 2> gpurun_out/bench_final_cfg2_f32.err
python bench.py --workload cfg5small --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_cfg5small_f64.json 2> gpurun_out/bench_final_cfg5small_f64.err
python bench.py --workload cfg5small --precision f32 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_cfg5small_f32.json 2> gpurun_out/bench_final_cfg5small_f32.err
python bench.py --workload cfg4shard --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_cfg4shard_f64.json 2> gpurun_out/bench_final_cfg4shard_f64.err
for f in gpurun_out/bench_final_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("final_")[1], "value %.3e ms/step %.4f frac %s e2e %s kernel %s" % (d["value"], d["ms_per_step"], r.get("frac"), d.get("e2e", {}).get("value"), r.get("kernel")))
except Exception as ex:
    print(sys.argv[1], "parse failed", ex)
PY
done
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final_cfg2_f64.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
echo "launch list rc $?"
$CMD > gpurun_out/plain_final.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bp_sweep_ell -s 12 -c 1 -f -o gpurun_out/prof_ell_cfg2_f64_r01_final $CMD > gpurun_out/ncu_ell_final.log 2>&1
echo "ncu ell rc $?"
CMD5="python bench.py --workload cfg5small --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD5 > gpurun_out/plain_final5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bp_sweep_wide -s 6 -c 1 -f -o gpurun_out/prof_wide_cfg5small_f64_r01_final $CMD5 > gpurun_out/ncu_wide_final.log 2>&1
echo "ncu wide rc $?"
