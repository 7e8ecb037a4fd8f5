"""BASELINE.json configurations 4 and 5 through the reference-shaped Python API
(IsingModel -> GPUAnnealer.anneal -> AnnealingResult), as the reference's callers use it
(problems/base.py:118-146)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, has_cuda

sys.path.insert(0, os.path.join(ROOT, "tools"))
import instances as inst  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_cuda():
        pytest.skip("needs a CUDA device")


def test_cfg4_tsp64_dense_model():
    """TSP QUBO, 64 cities = 4096 spins, penalty-encoded (float distances, penalties >> distances)."""
    import spin_glass_anneal_rl_b200 as sg
    J, h = inst.tsp_ising(inst.random_tsp(64, 4004))
    n = h.shape[0]
    assert n == 4096
    model = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=False))
    model.set_couplings_from_matrix(torch.from_numpy(J))
    model.external_fields = torch.from_numpy(h)
    e_start = float(-0.5 * model.spins.double() @ torch.from_numpy(J).double() @ model.spins.double()
                    - torch.from_numpy(h).double() @ model.spins.double())
    cfg = sg.GPUAnnealerConfig(n_sweeps=60, initial_temp=400.0, final_temp=1.0, random_seed=7,
                               schedule_params={"alpha": 0.9}, n_replicas=128, record_interval=10)
    res = sg.GPUAnnealer(cfg).anneal(model)
    s = res.best_configuration.double()
    assert s.shape == (n,) and set(torch.unique(s).tolist()) <= {-1.0, 1.0}
    e_best = float(-0.5 * s @ torch.from_numpy(J).double() @ s - torch.from_numpy(h).double() @ s)
    assert abs(res.best_energy - e_best) <= 1e-5 * abs(e_best)      # north_star float tolerance
    assert res.best_energy < e_start and res.n_sweeps == 60
    assert len(res.energy_history) == len(res.temperature_history) == 1 + 6


def test_cfg5_scheduling_sparse_model_through_the_api():
    """500 tasks x 100 agents = 50 000 spins, block cliques: a torch sparse COO model, as
    problems/base.py builds them; dense J would be 10 GB."""
    import spin_glass_anneal_rl_b200 as sg
    rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
    n = h.shape[0]
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    model = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=True))
    model.couplings = torch.sparse_coo_tensor(np.stack([rows, colidx]), torch.from_numpy(val), (n, n))
    model.external_fields = torch.from_numpy(h)
    cfg = sg.GPUAnnealerConfig(n_sweeps=6, initial_temp=200.0, final_temp=20.0, random_seed=3,
                               schedule_params={"alpha": 0.7}, n_replicas=64, record_interval=2)
    res = sg.GPUAnnealer(cfg).anneal(model)
    s = res.best_configuration.numpy().astype(np.float64)
    assert s.shape == (n,)
    # exact energy from the block structure: within task t, sum_{a != b} s_a s_b = (sum s)^2 - A
    blocks = s.reshape(500, 100)
    e_best = -0.5 * 50.0 * float(((blocks.sum(1) ** 2) - 100).sum()) - float(h.astype(np.float64) @ s)
    assert abs(res.best_energy - e_best) <= 1e-5 * abs(e_best)
    assert res.energy_history[-1] < res.energy_history[0]
    assert torch.equal(model.spins.abs(), torch.ones(n))


def test_cfg2_lattice_model_is_detected_and_swept_in_checkerboard_order():
    """A sparse COO model with the structure of the reference's 2D Edwards-Anderson generator is
    routed to the multi-spin-coded lattice kernel by the host API."""
    import spin_glass_anneal_rl_b200 as sg
    L = 80
    Jx, Jy = inst.ea_lattice_bonds(L, seed=8)
    rowptr, colidx, val, h = inst.lattice_csr(Jx, Jy)
    n = L * L
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    model = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=True))
    model.couplings = torch.sparse_coo_tensor(np.stack([rows, colidx]), torch.from_numpy(val), (n, n))
    cfg = sg.GPUAnnealerConfig(n_sweeps=40, initial_temp=3.0, final_temp=0.2, random_seed=5,
                               schedule_params={"alpha": 0.93}, n_replicas=96, record_interval=10)
    res = sg.GPUAnnealer(cfg).anneal(model)
    assert model._sg_engine[1].kind == "lattice"
    s = res.best_configuration.numpy().astype(np.float64).reshape(L, L)
    e = -float((Jx[:-1] * s[:-1] * s[1:]).sum() + (Jy[:, :-1] * s[:, :-1] * s[:, 1:]).sum())
    assert res.best_energy == e                      # integer energies: exact
    assert res.best_energy < -1.2 * n                # well below the random-start energy (~0)
    assert res.energy_history[-1] < res.energy_history[0]
    # a perturbed copy (one coupling 2.0) is no longer a +-J lattice -> sparse kernel
    val2 = val.copy()
    val2[0] = 2.0
    val2[np.where((rows == colidx[0]) & (colidx == rows[0]))[0][0]] = 2.0
    model2 = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=True))
    model2.couplings = torch.sparse_coo_tensor(np.stack([rows, colidx]), torch.from_numpy(val2), (n, n))
    sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=2, n_replicas=8, random_seed=1)).anneal(model2)
    assert model2._sg_engine[1].kind == "csr"


def test_dense_model_above_4096_spins_stays_dense(oracle):
    """Dense models with 4097..7168 spins run on the sequential-FMA sweep kernel (register-resident
    fields) instead of falling to the latency-bound sparse kernel."""
    import torch
    import spin_glass_anneal_rl_b200 as sg
    rng = np.random.default_rng(4500)
    n = 4500
    a = rng.integers(-1, 2, size=(n, n))
    J = np.triu(a, 1)
    J = (J + J.T).astype(np.float32)
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=False))
    m.set_couplings_from_matrix(torch.from_numpy(J))
    e0 = m.compute_energy()
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=5, initial_temp=2.0, final_temp=0.5, record_interval=2,
                                              random_seed=3, n_replicas=4)).anneal(m)
    assert m._sg_engine[1].kind == "dense"
    h = np.zeros(n, np.float32)
    assert oracle.energy(J, h, res.best_configuration.numpy()) == res.best_energy
    assert res.best_energy < e0 - 1000
    assert oracle.energy(J, h, m.spins.numpy()) == res.energy_history[-1]
