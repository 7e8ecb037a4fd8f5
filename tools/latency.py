"""Per-call latency of the reference-shaped API on small models (the RL environment's use:
repeated short constant-temperature anneals, reference rl_integration/environment.py:318-336)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spin_glass_anneal_rl_b200 as sg
from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType

for n, R, sweeps in ((50, 1, 10), (100, 1, 10), (100, 32, 10), (100, 32, 1000), (500, 64, 100)):
    rng = np.random.default_rng(n)
    a = rng.normal(size=(n, n)).astype(np.float32)
    J = np.triu(a, 1); J = J + J.T
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=False))
    m.set_couplings_from_matrix(torch.from_numpy(J))
    cfg = sg.GPUAnnealerConfig(n_sweeps=sweeps, initial_temp=1.0, final_temp=1.0,
                               schedule_type=ScheduleType.GEOMETRIC, schedule_params={"alpha": 1.0},
                               record_interval=sweeps, n_replicas=R, random_seed=1)
    ann = sg.GPUAnnealer(cfg)
    ann.anneal(m); torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 20
    for _ in range(reps):
        res = ann.anneal(m)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"n={n:4d} R={R:3d} sweeps={sweeps:5d}: {dt * 1e3:8.3f} ms per anneal() call, "
          f"{R * n * sweeps / dt / 1e6:9.3f} M attempts/s, best {res.best_energy:.3f}")

# ---- many models per launch (BatchProcessor): 64 models of N=100, 10 sweeps at constant T
from spin_glass_anneal_rl_b200.annealing.batch_processor import BatchConfig, BatchProcessor
rng = np.random.default_rng(0)
models = []
for _ in range(64):
    a = rng.normal(size=(100, 100)).astype(np.float32); J = np.triu(a, 1); J = J + J.T
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=100, use_sparse=False)); m.set_couplings_from_matrix(torch.from_numpy(J)); models.append(m)
for R in (1, 32):
    cfg = sg.GPUAnnealerConfig(n_sweeps=10, initial_temp=1.0, final_temp=1.0, schedule_type=ScheduleType.GEOMETRIC,
                               schedule_params={"alpha": 1.0}, record_interval=10, n_replicas=R, random_seed=1)
    bp = BatchProcessor(BatchConfig(batch_size=64), cfg)
    bp.process_models_batch(models); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): bp.process_models_batch(models)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    ann = sg.GPUAnnealer(cfg); ann.anneal(models[0]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for m in models: ann.anneal(m)
    torch.cuda.synchronize(); dt1 = time.perf_counter() - t0
    print(f"64 models N=100 R={R} 10 sweeps: stacked {dt * 1e3:.3f} ms per batch ({dt / 64 * 1e6:.1f} us per model), one anneal() per model {dt1 * 1e3:.3f} ms")
