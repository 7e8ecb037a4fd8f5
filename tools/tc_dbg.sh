#!/bin/bash
# timing experiments on the TC kernel (results are wrong with SG_TC_DBG set): constant operand
# chunk (bit 0), no MMA (bit 1).  Never use bit 2 with cluster pairs (the peer waits for the
# skipped remote stores).
for d in 0 1 2 3; do
  echo "== SG_TC_DBG=$d"; SG_TC_DBG=$d timeout 120 python tools/tc_check.py perf 2>&1 | grep "R=2368 P=3 sweeps=5 T=1.0\|R=2368 P=1"
done
for ns in 2 3; do
  echo "== SG_TC_STAGES=$ns"; SG_TC_STAGES=$ns timeout 120 python tools/tc_check.py perf 2>&1 | grep "R=2368 P=3 sweeps=5 T=1.0"
done
