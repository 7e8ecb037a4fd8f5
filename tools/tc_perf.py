"""Device-timed throughput of the tensor-core sweep at the headline shape for every cluster size
(development aid): python tools/tc_perf.py [planes] [replicas] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
P = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
n = 4096
rs = np.random.RandomState(3003)
Gm = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
J = ((Gm + Gm.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
eng = Engine(0)
eng.set_model(torch.from_numpy(J).cuda(), torch.zeros(n, device="cuda"))
eng.alloc_replicas(R)
S0 = (torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for cl in os.environ.get("TC_PERF_CLUSTERS", "4,2,1").split(","):
    if cl == "4": os.environ.pop("SG_TC_CLUSTER", None)
    else: os.environ["SG_TC_CLUSTER"] = cl
    for T in (1.0, 0.2):
        eng.set_spins(S0); eng.init_fields()
        eng.sweep(S, np.array([T]), seed=1, kernel="tc", coupling_planes=P)
        torch.cuda.synchronize()
        best = 1e9
        eng.set_profiling(True); eng.profile()
        for i in range(3):
            t0.record(); eng.sweep(S, np.array([T]), seed=1, sweep_base=S * (i + 1), kernel="tc", coupling_planes=P); t1.record()
            torch.cuda.synchronize(); best = min(best, t0.elapsed_time(t1))
        pr = eng.profile(); eng.set_profiling(False)
        acc = eng.accepted().double().mean().item() / (4 * S * n)
        print(f"C={eng.tc_cluster_size()} P={P} R={R} T={T}: {best:.3f} ms / {S} sweeps = {R * n * S / best / 1e6:.2f} G attempts/s "
              f"(sweep kernel {pr['sweep_ms'] / 3:.3f} ms, gather {pr['gather_ms'] / 3:.3f} ms) acc~{acc:.3f}", flush=True)
