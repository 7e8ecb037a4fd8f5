"""Throughput of the sparse sweep kernel on BASELINE cfg2 (EA L=256, 4096 replicas) and cfg5
(scheduling 500 x 100, 1024 replicas)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instances as inst
from spin_glass_anneal_rl_b200.engine import Engine
for name, model, R, T, sweeps in (("cfg2 EA L=256", inst.ea_lattice(256), 4096, 1.0, 3),
                                  ("cfg5 sched 500x100", inst.scheduling_ising(*inst.random_scheduling(500, 100)), 1024, 40.0, 3)):
    rowptr, colidx, val, h = model
    n = h.shape[0]
    eng = Engine(0)
    eng.set_model_csr(rowptr, colidx, val, h); eng.alloc_replicas(R)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8)); eng.init_fields()
    eng.sweep(1, np.array([T]), seed=1); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    a0 = eng.accepted().sum().item()
    t0.record(); eng.sweep(sweeps, np.array([T]), seed=1, sweep_base=1, track_best=False); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1); a1 = eng.accepted().sum().item()
    t0.record(); eng.sweep(sweeps, np.array([T]), seed=1, sweep_base=1 + sweeps, track_best=True); t1.record(); torch.cuda.synchronize()
    ms2 = t0.elapsed_time(t1)
    att = R * n * sweeps
    print(f"{name}: n={n} R={R}: {att / ms / 1e6:.3f} G attempts/s (track_best: {att / ms2 / 1e6:.3f}), acc={(a1 - a0) / att:.3f}, E/N={eng.energies().mean().item() / n:.4f}")
# ---- cfg2 on the checkerboard multi-spin-coded lattice kernel
Jx, Jy = inst.ea_lattice_bonds(256)
n, R = 65536, 4096
eng = Engine(0)
eng.set_model_lattice2d(Jx, Jy); eng.alloc_replicas(R)
eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8)); eng.init_fields()
temps = np.tile(np.geomspace(3.0, 0.1, 32), R // 32)
eng.sweep(2, temps, temps_replica_stride=1, seed=1, site_order="checkerboard"); torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for tb in (False, True):
    t0.record(); eng.sweep(20, temps, temps_replica_stride=1, seed=1, sweep_base=2, site_order="checkerboard", track_best=tb); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print(f"cfg2 EA L=256 lattice kernel: R={R} track_best={tb}: {R * n * 20 / ms / 1e6:.1f} G attempts/s ({ms / 20:.3f} ms/sweep), E/N={eng.energies().mean().item() / n:.4f}")
# ---- cfg5 on the block-clique (group-sum) kernel
rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
n, R = 50000, 1024
eng = Engine(0)
eng.set_model_groups((np.arange(n) // 100).astype(np.int32), np.full(500, 50.0, np.float32), h); eng.alloc_replicas(R)
eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8)); eng.init_fields()
eng.sweep(2, np.array([40.0]), seed=1); torch.cuda.synchronize()
for tb in (False, True):
    t0.record(); eng.sweep(20, np.array([40.0]), seed=1, sweep_base=2, track_best=tb); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print(f"cfg5 sched 500x100 groups kernel: R={R} track_best={tb}: {R * n * 20 / ms / 1e6:.2f} G attempts/s ({ms / 20:.3f} ms/sweep), E/N={eng.energies().mean().item() / n:.3f}")
