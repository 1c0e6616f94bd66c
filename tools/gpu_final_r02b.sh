#!/bin/bash
# Last measurement pass of round 2 (after the tile-kernel rework): test suite, configs[3]-shard bench line, --set full capture of
# the pipeline kernel on that shard, configs[2] bench line.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r02b.log; cat gpurun_out/pytest_r02b.log
timeout 100 python bench.py --workload cfg4shard --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_r02b_cfg4shard_f64.json 2> gpurun_out/bench_r02b_cfg4shard.err; cut -c1-400 gpurun_out/bench_r02b_cfg4shard_f64.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bp_sweep_pipe -s 5 -c 1 -o gpurun_out/prof_pipe_cfg4shard_r02b -f python bench.py --workload cfg4shard --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r02b.log 2>&1; tail -1 gpurun_out/ncu_r02b.log
timeout 150 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_r02b_cfg3_f64.json 2> gpurun_out/bench_r02b_cfg3.err; cut -c1-300 gpurun_out/bench_r02b_cfg3_f64.json
