#!/bin/bash
# Round-2 closing evidence (under gpurun): GPU test suite, the bench line, and the ncu launch list of
# the bench command with the concurrent cluster-pair launch in place.  Outputs land in gpurun_out/.
set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
B="python bench.py --steps 5 --warmup 3 --ttt-budget 0 --cpu-budget 2"
timeout 200 $B > gpurun_out/bench_r2c_short.log 2> gpurun_out/bench_r2c_short.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/launches_r2c.csv $B > gpurun_out/ncu_bench_r2c.log 2>&1
tail -c 400 gpurun_out/bench_r2c.json; tail -2 gpurun_out/bench_r2c.err; tail -2 gpurun_out/ncu_bench_r2c.log
