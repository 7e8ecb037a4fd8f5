"""Edge cases of the C ABI on the GPU: empty launches, single replicas, smallest sizes, the
multi-GPU wrapper without a process group, model-kind switches on one engine."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, has_cuda

sys.path.insert(0, os.path.join(ROOT, "tools"))
import instances as inst  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    from spin_glass_anneal_rl_b200.engine import Engine
    return Engine(0)


def _models():
    rng = np.random.default_rng(0)
    a = rng.integers(-2, 3, size=(20, 20))
    J = np.triu(a, 1)
    J = (J + J.T).astype(np.float32)
    h = rng.integers(-1, 2, size=20).astype(np.float32)
    yield "dense", 20, "random", lambda e: e.set_model(J, h)
    rp, ci, v, hh = inst.scheduling_ising(*inst.random_scheduling(4, 5, seed=1))
    yield "csr", 20, "random", lambda e: e.set_model_csr(rp, ci, v, np.round(hh))
    yield "groups", 20, "random", lambda e: e.set_model_groups((np.arange(20) // 5).astype(np.int32),
                                                              np.full(4, 50.0, np.float32), np.round(hh))
    Jx, Jy = inst.ea_lattice_bonds(5, seed=1)
    yield "lattice", 25, "checkerboard", lambda e: e.set_model_lattice2d(Jx, Jy)


@pytest.mark.parametrize("R", [1, 33])
def test_every_model_kind_on_one_engine(engine, R):
    """Switching kinds on one engine, R = 1 and a ragged R, zero-sweep launches, energy bookkeeping."""
    rng = np.random.default_rng(R)
    for kind, n, order, setter in _models():
        setter(engine)
        engine.alloc_replicas(R)
        S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
        engine.set_spins(S0)
        engine.init_fields()
        e0 = engine.energies().cpu().numpy()
        engine.sweep(0, np.array([1.0]), site_order=order)          # no-op
        assert np.array_equal(engine.spins().cpu().numpy(), S0), kind
        assert engine.accepted().sum().item() == 0
        engine.sweep(4, np.array([2.0]), seed=R, site_order=order)
        e1 = engine.energies().cpu().numpy()
        assert np.array_equal(e1, engine.batch_energies(engine.spins()).cpu().numpy()), kind
        be, bs = engine.best()
        assert np.array_equal(be.cpu().numpy(), engine.batch_energies(bs).cpu().numpy()), kind
        assert (be.cpu().numpy() <= np.minimum(e0, e1)).all(), kind
        assert set(np.unique(engine.spins().cpu().numpy())) <= {-1, 1}


def test_wrong_call_order_and_bad_arguments(engine):
    from spin_glass_anneal_rl_b200._lib import SGError
    rng = np.random.default_rng(1)
    J = np.zeros((6, 6), np.float32)
    engine.set_model(J, np.zeros(6, np.float32))
    engine.alloc_replicas(2)
    engine.set_spins(np.ones((2, 6), np.int8))
    with pytest.raises(SGError):          # fields not initialised
        engine.sweep(1, np.array([1.0]))
    engine.init_fields()
    with pytest.raises(SGError):          # checkerboard is a lattice order
        engine.sweep(1, np.array([1.0]), site_order="checkerboard")
    with pytest.raises(SGError):          # group ids out of range
        engine.set_model_groups(np.array([0, 3], np.int32), np.ones(2, np.float32), np.zeros(2, np.float32))
    with pytest.raises(SGError):          # too big for shared memory -> caller must use CSR
        engine.set_model_groups(np.zeros(70000, np.int32), np.ones(1, np.float32), np.zeros(70000, np.float32))
    with pytest.raises(SGError):          # periodic lattice with odd L has no checkerboard
        engine.set_model_lattice2d(np.ones((5, 5), np.int8), np.ones((5, 5), np.int8))


def test_multi_gpu_annealer_without_process_group():
    import torch
    import spin_glass_anneal_rl_b200 as sg
    from spin_glass_anneal_rl_b200.annealing import MultiGPUAnnealer, MultiGPUConfig
    J, h = inst.random_dense(100, 1001)
    model = sg.IsingModel(sg.IsingModelConfig(n_spins=100, use_sparse=False))
    model.set_couplings_from_matrix(torch.from_numpy(J))
    model.external_fields = torch.from_numpy(h)
    for strategy, acfg in (("data_parallel", sg.GPUAnnealerConfig(n_sweeps=60, random_seed=1, initial_temp=5.0)),
                           ("replica_exchange", sg.ParallelTemperingConfig(n_sweeps=60, random_seed=1))):
        mg = MultiGPUAnnealer(MultiGPUConfig(n_replicas=64, strategy=strategy, n_rungs=8), acfg)
        res = mg.anneal(model)
        s = res.best_configuration.double()
        e = float(-0.5 * s @ torch.from_numpy(J).double() @ s - torch.from_numpy(h).double() @ s)
        assert abs(res.best_energy - e) <= 1e-5 * abs(e)
        assert res.best_energy < -60.0          # ground state of this instance is about -88
