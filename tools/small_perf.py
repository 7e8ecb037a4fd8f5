"""cfg1-shaped models (N = 100 dense) on the small-model kernel against the big kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for n, R, sweeps in ((100, 32, 1000), (100, 1024, 200), (100, 16384, 50), (224, 4096, 50), (32, 8192, 100)):
    g = torch.Generator().manual_seed(1001)
    A = torch.randn(n, n, generator=g); J = (A + A.T) / 2; J.fill_diagonal_(0.0); h = 0.5 * torch.randn(n, generator=g)
    for kern in ("small", "tc", "simt"):
        if kern == "tc" and n < 16: continue
        eng = Engine(0); eng.set_model(J, h); eng.alloc_replicas(R)
        eng.set_spins((torch.randint(0, 2, (R, n), generator=g) * 2 - 1).to(torch.int8)); eng.init_fields()
        temps = np.geomspace(5.0, 0.01, sweeps)
        eng.sweep(sweeps, temps, temps_sweep_stride=1, seed=1, kernel=kern); torch.cuda.synchronize()
        t0.record(); eng.sweep(sweeps, temps, temps_sweep_stride=1, seed=1, sweep_base=sweeps, kernel=kern); t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        print(f"n={n} R={R} sweeps={sweeps} kernel={kern}: {ms:.3f} ms, {R * n * sweeps / ms / 1e6:.3f} G attempts/s, best {eng.best_energies().min().item():.3f}")
