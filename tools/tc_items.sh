#!/bin/bash
timeout 300 python tools/tc_check.py vs 2>&1 | tail -2
timeout 300 python tools/tc_check.py replay 2>&1 | tail -1
timeout 300 python tools/tc_check.py perf 2>&1 | grep "P=3"
timeout 120 python tools/tc_timeline.py 3 2>&1 | grep "period\|quarter:\|decision:\|mma:\|CTA duration"
