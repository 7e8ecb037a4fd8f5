#!/bin/bash
timeout 200 python tools/mma_bench.py 2>&1 | tail -6
timeout 300 python tools/tc_check.py replay 2>&1 | tail -3
timeout 300 python tools/tc_check.py vs 2>&1 | tail -2
timeout 300 python tools/tc_timeline.py 3 2>&1 | grep "item phases\|period\|mma:\|quarter:\|decision:"
timeout 300 python tools/tc_check.py perf 2>&1 | tail -8
