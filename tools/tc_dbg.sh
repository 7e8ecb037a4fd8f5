#!/bin/bash
# timing experiments for the TC kernel (results with SG_TC_DBG != 0 are wrong by construction)
for d in 0 1 2 3 4 7; do
  echo "== SG_TC_DBG=$d"; SG_TC_DBG=$d timeout 120 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["x", "none"]
exec(open("tools/tc_check.py").read().split("if what in")[0])
for P in (3, 1):
    perf(planes=P, reps=2)
PY
done
