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

from oracle import oracle as orc   # the CPU arm beside every number (test infrastructure, timing only)

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
CPU_BUDGET = float(os.environ.get("CPU_BUDGET", "4"))
try:
    CORES = len(os.sched_getaffinity(0))
except AttributeError:
    CORES = os.cpu_count() or 1


def cpu_arm(n, T, dense=None, csr=None):
    """Reference algorithm (oracle port, all host cores) on a bounded sample of the same instance:
    one replica per core, as many sweeps as fit CPU_BUDGET seconds (at least one)."""
    import time
    rng = np.random.default_rng(3)
    S = (rng.integers(0, 2, size=(CORES, n)) * 2 - 1).astype(np.float32)
    run = (lambda sw, seed: orc.baseline_run(dense[0], dense[1], S, sw, T, seed=seed, n_threads=CORES)) if dense \
        else (lambda sw, seed: orc.baseline_run_csr(csr[0], csr[1], csr[2], csr[3], S, sw, T, seed=seed, n_threads=CORES))
    w = time.perf_counter(); run(1, 1); dt = time.perf_counter() - w
    sw = max(1, int(CPU_BUDGET / max(dt, 1e-4)))
    w = time.perf_counter(); att, _ = run(sw, 2); dt = time.perf_counter() - w
    return {"attempts_per_s": att / dt, "cores": CORES, "kind": "port" + ("" if dense else " (CSR rows instead of dense dots)"),
            "sample": f"{CORES} replicas x {sw} sweeps at T={T} ({dt:.1f} s)"}

t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
out = []


def run(name, eng, n, R, sweeps, reps=3, cpu=None, **kw):
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
    if cpu is not None:
        rec["cpu_baseline"] = cpu
        rec["gpu_over_cpu"] = rec["attempts_per_s"] / cpu["attempts_per_s"]
    out.append(rec)
    print(f"{name}: n={n} R={R}: {rec['attempts_per_s'] / 1e9:.2f} G attempts/s ({best / sweeps:.3f} ms/sweep), acc={rec['acceptance']:.3f}"
          + (f" | CPU {cpu['attempts_per_s'] / 1e6:.2f} M attempts/s on {cpu['cores']} cores" if cpu else ""), flush=True)


# cfg1: N = 100 dense Gaussian, 32 replicas, geometric schedule 5.0 -> 0.01, 1000 sweeps
J, h = inst.random_dense(100)
eng = Engine(0); eng.set_model(J, h)
run("cfg1 N=100 dense, 32 replicas, 1000 sweeps geometric (K1-SMALL)", eng, 100, 32, 1000,
    cpu=cpu_arm(100, 1.0, dense=(J, h)), temps=np.maximum(5.0 * 0.95 ** np.arange(1000), 0.01), temps_sweep_stride=1)
# cfg2: EA +-J L = 256 open boundaries, 4096 replicas, checkerboard, ladder T in [0.1, 3.0]
Jx, Jy = inst.ea_lattice_bonds(256)
eng = Engine(0); eng.set_model_lattice2d(Jx, Jy)
_rp, _ci, _v, _h0 = inst.lattice_csr(Jx, Jy)
run("cfg2 EA +-J L=256, 4096 replicas, checkerboard multi-spin (K1-LAT)", eng, 65536, 4096, 20,
    cpu=cpu_arm(65536, 1.0, csr=(_rp, _ci, _v, _h0)), temps=np.tile(np.geomspace(3.0, 0.1, 32), 128), temps_replica_stride=1, site_order="checkerboard")
# cfg3: SK N = 4096, 8192 replicas (the headline; bench.py measures it with the full contract)
J, h = inst.sk(4096)
eng = Engine(0); eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda())
run("cfg3 SK N=4096, 8192 replicas, T=1 (K1-TC, 3 planes)", eng, 4096, 8192, 10, cpu=cpu_arm(4096, 1.0, dense=(J, h)),
    temps=np.array([1.0]), kernel="tc")
# cfg4: TSP 64 cities position encoding (4096 spins, dense, penalties >> distances), 2048 replicas
J, h = inst.tsp_ising(inst.random_tsp(64))
eng = Engine(0); eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda())
run("cfg4 TSP-64 QUBO (4096 spins), 2048 replicas, T=3000 (K1-TC, 3 planes)", eng, 4096, 2048, 10,
    cpu=cpu_arm(4096, 3000.0, dense=(J, h)), temps=np.array([3000.0]), kernel="tc")
# cfg5: scheduling QUBO 500 tasks x 100 agents (50 000 spins, block cliques), 1024 replicas
rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
eng = Engine(0); eng.set_model_groups((np.arange(50000) // 100).astype(np.int32), np.full(500, 50.0, np.float32), h)
run("cfg5 scheduling 500x100 (50k spins), 1024 replicas, T=3000 (K1-GRP partitioned)", eng, 50000, 1024, 10,
    cpu=cpu_arm(50000, 3000.0, csr=(rowptr, colidx, val, h)), temps=np.array([3000.0]))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_configs.json"), "w"), indent=1)
