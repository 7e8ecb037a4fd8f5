import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spin_glass_anneal_rl_b200 as sg
from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType
from spin_glass_anneal_rl_b200.annealing.batch_processor import BatchConfig, BatchProcessor
rng = np.random.default_rng(0)
models = []
for _ in range(64):
    a = rng.normal(size=(100, 100)).astype(np.float32); J = np.triu(a, 1); J = J + J.T
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=100, use_sparse=False)); m.set_couplings_from_matrix(torch.from_numpy(J)); models.append(m)
cfg = sg.GPUAnnealerConfig(n_sweeps=10, initial_temp=1.0, final_temp=1.0, schedule_type=ScheduleType.GEOMETRIC,
                           schedule_params={"alpha": 1.0}, record_interval=10, n_replicas=32, random_seed=1)
bp = BatchProcessor(BatchConfig(batch_size=64), cfg)
bp.process_models_batch(models); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): bp.process_models_batch(models)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
