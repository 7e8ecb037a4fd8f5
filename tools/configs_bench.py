"""Throughput of the five BASELINE.json configurations through the engine (device-resident inputs,
CUDA events, after a warm-up launch); writes profiles/<tag>_configs.json.
usage: python tools/configs_bench.py [tag]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import instances as inst
from spin_glass_anneal_rl_b200.engine import Engine

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
out = []


def run(name, eng, n, R, sweeps, reps=3, **kw):
    g = torch.Generator(device="cuda").manual_seed(7)
    eng.alloc_replicas(R)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8))
    eng.init_fields()
    eng.sweep(sweeps, seed=1, **kw); torch.cuda.synchronize()
    a0 = eng.accepted().sum().item()
    best = 1e30
    for i in range(reps):
        t0.record(); eng.sweep(sweeps, seed=1, sweep_base=(i + 1) * sweeps, **kw); t1.record()
        torch.cuda.synchronize(); best = min(best, t0.elapsed_time(t1))
    a1 = eng.accepted().sum().item()
    att = R * n * sweeps
    rec = {"config": name, "n_spins": n, "replicas": R, "sweeps_per_launch": sweeps, "ms_per_launch": best,
           "attempts_per_s": att / best * 1e3, "acceptance": (a1 - a0) / (att * reps),
           "mean_energy_per_spin": eng.energies().mean().item() / n}
    out.append(rec)
    print(f"{name}: n={n} R={R}: {rec['attempts_per_s'] / 1e9:.2f} G attempts/s ({best / sweeps:.3f} ms/sweep), acc={rec['acceptance']:.3f}")


# cfg1: N = 100 dense Gaussian, 32 replicas, geometric schedule 5.0 -> 0.01, 1000 sweeps
J, h = inst.random_dense(100)
eng = Engine(0); eng.set_model(J, h)
run("cfg1 N=100 dense, 32 replicas, 1000 sweeps geometric (K1-SMALL)", eng, 100, 32, 1000,
    temps=np.maximum(5.0 * 0.95 ** np.arange(1000), 0.01), temps_sweep_stride=1)
# cfg2: EA +-J L = 256 open boundaries, 4096 replicas, checkerboard, ladder T in [0.1, 3.0]
Jx, Jy = inst.ea_lattice_bonds(256)
eng = Engine(0); eng.set_model_lattice2d(Jx, Jy)
run("cfg2 EA +-J L=256, 4096 replicas, checkerboard multi-spin (K1-LAT)", eng, 65536, 4096, 20,
    temps=np.tile(np.geomspace(3.0, 0.1, 32), 128), temps_replica_stride=1, site_order="checkerboard")
# cfg3: SK N = 4096, 8192 replicas (the headline; bench.py measures it with the full contract)
J, h = inst.sk(4096)
eng = Engine(0); eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda())
run("cfg3 SK N=4096, 8192 replicas, T=1 (K1-TC, 3 planes)", eng, 4096, 8192, 10, temps=np.array([1.0]), kernel="tc")
# cfg4: TSP 64 cities position encoding (4096 spins, dense, penalties >> distances), 2048 replicas
J, h = inst.tsp_ising(inst.random_tsp(64))
eng = Engine(0); eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda())
run("cfg4 TSP-64 QUBO (4096 spins), 2048 replicas, T=3000 (K1-TC, 3 planes)", eng, 4096, 2048, 10,
    temps=np.array([3000.0]), kernel="tc")
# cfg5: scheduling QUBO 500 tasks x 100 agents (50 000 spins, block cliques), 1024 replicas
rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
eng = Engine(0); eng.set_model_groups((np.arange(50000) // 100).astype(np.int32), np.full(500, 50.0, np.float32), h)
run("cfg5 scheduling 500x100 (50k spins), 1024 replicas, T=3000 (K1-GRP partitioned)", eng, 50000, 1024, 10,
    temps=np.array([3000.0]))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_configs.json"), "w"), indent=1)
