import sys, os
sys.path.insert(0, os.getcwd()); sys.argv = ["x", "none"]
exec(open("tools/tc_check.py").read().split("if what in")[0])
print("SG_TC_DBG", os.environ.get("SG_TC_DBG"), "SG_TC_CLUSTER", os.environ.get("SG_TC_CLUSTER"))
perf(R=2368, planes=3, reps=2)
perf(R=1024, planes=3, reps=2)
