#!/bin/bash
# Round-2 evidence run (under gpurun): ncu launch list of the bench command, one ncu --set full
# capture of the dominant kernel and one of the L2 stream probe.  Outputs land in gpurun_out/.
set -x
B="python bench.py --steps 5 --warmup 3 --ttt-budget 0 --cpu-budget 2"
$B > gpurun_out/bench_r2_short.log 2> gpurun_out/bench_r2_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_bench_r2.log 2>&1
python tools/prof.py tc > gpurun_out/prof_plain_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_tc -s 1 -c 1 \
    -o gpurun_out/prof_r2_tc python tools/prof.py tc > gpurun_out/ncu_prof_r2.log 2>&1
python tools/prof.py tma > gpurun_out/prof_tma_plain_r2.log 2>&1 &&
ncu --set full --clock-control none -k regex:tma_probe -c 4 \
    -o gpurun_out/prof_r2_tma python tools/prof.py tma > gpurun_out/ncu_tma_r2.log 2>&1
tail -c 600 gpurun_out/bench_r2_short.log; tail -3 gpurun_out/ncu_prof_r2.log; tail -3 gpurun_out/ncu_tma_r2.log; cat gpurun_out/prof_tma_plain_r2.log
