#!/bin/bash
# Round-1 evidence run (under gpurun): bench line, ncu launch list of the same command, and one
# ncu --set full capture of the dominant kernel.  Outputs land in gpurun_out/.
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.log 2> gpurun_out/bench_r1.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/prof.py tc > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 1 -c 1 \
    -o gpurun_out/prof_r1_tc python tools/prof.py tc > gpurun_out/ncu_prof.log 2>&1
tail -2 gpurun_out/bench_r1.log; tail -3 gpurun_out/ncu_prof.log
