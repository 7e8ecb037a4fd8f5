"""cfg5 (scheduling 500 x 100, 1024 replicas) on the two group kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instances as inst
from spin_glass_anneal_rl_b200.engine import Engine
rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
n = 50000
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for R in (1024, 4096):
    for flag in ("0", "1"):
        os.environ["SG_GRP_PART"] = flag
        eng = Engine(0)
        eng.set_model_groups((np.arange(n) // 100).astype(np.int32), np.full(500, 50.0, np.float32), h); eng.alloc_replicas(R)
        g = torch.Generator(device="cuda").manual_seed(1)
        eng.set_spins((torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8)); eng.init_fields()
        eng.sweep(2, np.array([40.0]), seed=1); torch.cuda.synchronize()
        for tb in (False, True):
            t0.record(); eng.sweep(10, np.array([40.0]), seed=1, sweep_base=2, track_best=tb); t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
            print(f"cfg5 groups kernel part={flag} R={R} track_best={tb}: {R * n * 10 / ms / 1e6:.2f} G attempts/s ({ms / 10:.3f} ms/sweep), E/N={eng.energies().mean().item() / n:.3f}")
