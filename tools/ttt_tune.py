"""Time to target (E <= 0.97 x Parisi) of parallel tempering on SK N=4096, 8192 replicas, for a few
ladders / exchange intervals (tuning aid for bench.py's time_to_target)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instances as inst
from spin_glass_anneal_rl_b200.engine import Engine
n, R = 4096, 8192
J, h = inst.sk(n)
eng = Engine(0); eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda()); eng.alloc_replicas(R)
e_target = 0.97 * (-0.7632 / np.sqrt(2.0)) * n
g = torch.Generator(device="cuda")
def run(tmax, tmin, K, spe, seed):
    g.manual_seed(seed)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8)); eng.init_fields()
    eng.set_ladder(np.geomspace(tmax, tmin, K)); torch.cuda.synchronize()
    w0 = time.perf_counter(); rounds = 0
    while time.perf_counter() - w0 < 3.0:
        for _ in range(2):
            eng.sweep(spe, None, seed=42 + seed, sweep_base=rounds * spe, site_order="random", track_best=True, kernel="tc", coupling_planes=3)
            eng.refresh_fields(); eng.exchange(rounds & 1, seed=7 + seed, round=rounds); rounds += 1
        if eng.best_energies().min().item() <= e_target: break
    torch.cuda.synchronize(); return time.perf_counter() - w0, rounds * spe
for tmax, tmin, K, spe in ((2.0, 0.05, 64, 10), (1.0, 0.1, 64, 10), (2.0, 0.1, 32, 10), (2.0, 0.1, 64, 5), (0.5, 0.05, 64, 10),
                           (3.0, 0.1, 64, 10), (1.0, 0.05, 32, 5), (2.0, 0.1, 16, 10)):
    rs = [run(tmax, tmin, K, spe, s) for s in range(4)]
    print(f"T {tmax}->{tmin} K={K} sweeps/exchange={spe}: median {np.median([r[0] for r in rs]):.3f} s, sweeps {[r[1] for r in rs]}")
