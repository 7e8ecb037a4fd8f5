"""Tiny run of every kernel family (for compute-sanitizer --tool memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instances as inst
from spin_glass_anneal_rl_b200.engine import Engine
rng = np.random.default_rng(0)
eng = Engine(0)
# dense: K2-TC, TC sweep, SIMT sweep, exchange
n, R = 300, 40
a = rng.integers(-2, 3, size=(n, n)); J = np.triu(a, 1); J = (J + J.T).astype(np.float32)
h = rng.integers(-2, 3, size=n).astype(np.float32)
S = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
eng.set_model(J, h); eng.alloc_replicas(R); eng.set_spins(S); eng.init_fields()
eng.sweep(3, np.array([1.5]), seed=1, kernel="tc")
eng.sweep(2, np.array([1.5]), seed=1, sweep_base=3, kernel="simt")
eng.set_ladder(np.geomspace(3, 0.3, 8)); eng.sweep(2, None, seed=2, sweep_base=5, kernel="tc"); eng.exchange(0, seed=3, round=0)
eng.refresh_fields(); e1 = eng.energies().cpu().numpy(); e2 = eng.batch_energies(eng.spins()).cpu().numpy()
assert np.array_equal(e1, e2)
# dense, cluster forms of the tensor-core sweep (clusters of 4: 64 replicas per group; pairs; ragged groups)
n, R = 1024, 70
a = rng.integers(-2, 3, size=(n, n)); J = np.triu(a, 1); J = (J + J.T).astype(np.float32)
S = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
eng.set_model(J, np.zeros(n, np.float32)); eng.alloc_replicas(R); eng.set_spins(S); eng.init_fields()
assert eng.tc_cluster_size() == 4
eng.sweep(2, np.array([1.5]), seed=1, kernel="tc")
os.environ["SG_TC_CLUSTER"] = "2"
eng.sweep(1, np.array([1.5]), seed=1, sweep_base=2, kernel="tc")
os.environ.pop("SG_TC_CLUSTER")
eng.refresh_fields(); e1 = eng.energies().cpu().numpy(); e2 = eng.batch_energies(eng.spins()).cpu().numpy()
assert np.array_equal(e1, e2)
hit = torch.full((2,), -1, dtype=torch.int32, device="cuda"); eng.check_target(0.0, 0, hit); eng.best_config()
eng.set_ladder(np.geomspace(3, 0.3, 7)); eng.exchange(0, seed=3, round=0); eng.exchange(0, seed=3, round=1, method="all_pairs")
# sparse
rowptr, colidx, val, hh = inst.scheduling_ising(*inst.random_scheduling(12, 8, seed=3))
eng.set_model_csr(rowptr, colidx, val, np.round(hh)); eng.alloc_replicas(37)
eng.set_spins((rng.integers(0, 2, size=(37, 96)) * 2 - 1).astype(np.int8)); eng.init_fields()
eng.sweep(3, np.array([30.0]), seed=4)
assert np.array_equal(eng.energies().cpu().numpy(), eng.batch_energies(eng.spins()).cpu().numpy())
# lattice
Jx, Jy = inst.ea_lattice_bonds(12, seed=1)
eng.set_model_lattice2d(Jx, Jy); eng.alloc_replicas(45)
eng.set_spins((rng.integers(0, 2, size=(45, 144)) * 2 - 1).astype(np.int8)); eng.init_fields()
eng.sweep(3, np.array([1.0]), seed=5, site_order="checkerboard")
assert np.array_equal(eng.energies().cpu().numpy(), eng.batch_energies(eng.spins()).cpu().numpy())
torch.cuda.synchronize()
print("sanitize_small ok")
