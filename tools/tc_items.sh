#!/bin/bash
timeout 300 python tools/tc_check.py replay 2>&1 | tail -3
timeout 300 python tools/tc_check.py vs 2>&1 | tail -4
SG_TC_SM=2 timeout 300 python tools/tc_check.py vs 2>&1 | tail -4
SG_TC_SM=2 SG_TC_SPI=1 timeout 300 python tools/tc_check.py vs 2>&1 | tail -4
SG_TC_VERBOSE=1 timeout 300 python tools/tc_check.py perf 2>&1 | grep "P=3\|grid=148 smem=228304 NS=4 sweeps/item=[25]" | sort -u
timeout 300 python tools/tc_check.py drift 2>&1 | tail -4
