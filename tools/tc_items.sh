#!/bin/bash
timeout 300 python tools/tc_check.py replay 2>&1 | tail -1
timeout 300 python tools/tc_check.py vs 2>&1 | tail -2
timeout 300 python tools/tc_check.py perf 2>&1 | grep "P=3\|P=1"
timeout 120 python tools/tc_timeline.py 3 2>&1 | grep "period\|decision:\|mma:"
SG_TC_CLUSTER=1 timeout 300 python tools/tc_check.py perf 2>&1 | grep "R=2368 P=3 sweeps=5 T=1.0"
