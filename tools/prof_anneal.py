import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spin_glass_anneal_rl_b200 as sg
from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType
n = 100
rng = np.random.default_rng(n)
a = rng.normal(size=(n, n)).astype(np.float32); J = np.triu(a, 1); J = J + J.T
m = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=False)); m.set_couplings_from_matrix(torch.from_numpy(J))
cfg = sg.GPUAnnealerConfig(n_sweeps=10, initial_temp=1.0, final_temp=1.0, schedule_type=ScheduleType.GEOMETRIC,
                           schedule_params={"alpha": 1.0}, record_interval=10, n_replicas=1, random_seed=1)
ann = sg.GPUAnnealer(cfg)
for _ in range(5): ann.anneal(m)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): ann.anneal(m)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
