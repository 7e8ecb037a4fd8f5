"""GPU parity of the sparse (CSR) sweep path through the C ABI: scaled-down cfg2 (2D +-J lattice)
and cfg5 (scheduling QUBO) against the oracle (dense restatement of the reference) in replay
mode, against the dense kernels in Philox mode, and size-independent properties at full size."""
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


def _sched(T, A, integer_h):
    rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(T, A, seed=T * A))
    if integer_h:
        h = np.round(h).astype(np.float32)
    return rowptr, colidx, val, h


def _random_sparse(rng, n, deg, symmetric=True, diag=False):
    J = np.zeros((n, n), np.float32)
    for i in range(n):
        cols = rng.choice(n, size=deg, replace=False)
        J[i, cols] = rng.integers(-3, 4, size=deg)
    if symmetric:
        J = np.triu(J, 1)
        J = J + J.T
    if not diag:
        np.fill_diagonal(J, 0.0)
    rowptr = np.zeros(n + 1, np.int64)
    colidx, val = [], []
    for i in range(n):
        nz = np.nonzero(J[i])[0]
        colidx.append(nz)
        val.append(J[i, nz])
        rowptr[i + 1] = rowptr[i] + len(nz)
    h = rng.integers(-2, 3, size=n).astype(np.float32)
    return rowptr, np.concatenate(colidx).astype(np.int32), np.concatenate(val).astype(np.float32), h, J


CASES = {
    "ea_L8": lambda: inst.ea_lattice(8, seed=1),
    "ea_L16": lambda: inst.ea_lattice(16, seed=2),
    "sched_6x5": lambda: _sched(6, 5, True),
    "sched_12x8": lambda: _sched(12, 8, True),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("rule", ["metropolis", "glauber"])
def test_csr_replay_is_bit_exact(engine, oracle, case, rule):
    rowptr, colidx, val, h = CASES[case]()
    n = h.shape[0]
    J = inst.csr_to_dense(rowptr, colidx, val, n)
    rng = np.random.default_rng(n)
    R, ns = 37, 3
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
    uni = rng.random((R, ns, n), dtype=np.float32)
    temps = np.array([30.0, 10.0, 2.0]) if "sched" in case else np.array([2.5, 1.5, 0.7])
    engine.set_model_csr(rowptr, colidx, val, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    Fo, Eo = oracle.batch_fields_energies(J, h, S0.astype(np.float32))
    assert np.array_equal(engine.fields().cpu().numpy().astype(np.float64), Fo)
    assert np.array_equal(engine.energies().cpu().numpy().astype(np.float64), Eo)
    trace = engine.sweep(ns, temps, temps_sweep_stride=1, rule=rule, sites=sites, uniforms=uni,
                         energy_trace=True).cpu().numpy()
    final = engine.spins().cpu().numpy()
    acc = engine.accepted().cpu().numpy()
    best_e, best_s = engine.best()
    for r in range(R):
        s = S0[r].astype(np.float32).copy()
        e0 = oracle.energy(J, h, s)
        es, ac = oracle.sweeps_scheduled(J, h, s, temps, rule, sites, uni[r])
        assert np.array_equal(final[r], s.astype(np.int8)), f"replica {r} trajectory differs"
        assert np.array_equal(trace[:, r].astype(np.float64), es)
        assert int(acc[r]) == int(ac.sum())
        assert float(best_e[r]) == min(e0, es.min())
    Fo, Eo = oracle.batch_fields_energies(J, h, final.astype(np.float32))
    assert np.array_equal(engine.fields().cpu().numpy().astype(np.float64), Fo)
    assert np.array_equal(engine.batch_energies(best_s).cpu().numpy(), best_e.cpu().numpy())


def test_csr_asymmetric_couplings_and_diagonal(engine, oracle):
    """Rows of J define the local field (core/ising_model.py:176-185), also for sparse models."""
    rng = np.random.default_rng(9)
    rowptr, colidx, val, h, J = _random_sparse(rng, 60, 6, symmetric=False, diag=True)
    R, ns = 5, 2
    S0 = (rng.integers(0, 2, size=(R, 60)) * 2 - 1).astype(np.int8)
    sites = rng.integers(0, 60, size=(ns, 60)).astype(np.int32)
    uni = rng.random((R, ns, 60), dtype=np.float32)
    temps = np.array([3.0, 1.0])
    engine.set_model_csr(rowptr, colidx, val, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    trace = engine.sweep(ns, temps, temps_sweep_stride=1, sites=sites, uniforms=uni,
                         energy_trace=True).cpu().numpy()
    final = engine.spins().cpu().numpy()
    for r in range(R):
        s = S0[r].astype(np.float32).copy()
        es, _ = oracle.sweeps_scheduled(J, h, s, temps, "metropolis", sites, uni[r])
        assert np.array_equal(final[r], s.astype(np.int8))
        assert np.array_equal(trace[:, r].astype(np.float64), es)


def test_csr_equals_dense_kernels_in_philox_mode(engine):
    """Same Philox counters and site order: the sparse kernel, the sequential-FMA kernel and the
    tensor-core kernel walk identical trajectories on integer couplings."""
    rng = np.random.default_rng(21)
    n, R, ns = 400, 70, 4
    rowptr, colidx, val, h, J = _random_sparse(rng, n, 12)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    outs = []
    for mode in ("csr", "simt", "tc"):
        if mode == "csr":
            engine.set_model_csr(rowptr, colidx, val, h)
        else:
            engine.set_model(J, h)
        engine.alloc_replicas(R)
        engine.set_spins(S0)
        engine.init_fields()
        kw = {} if mode == "csr" else {"kernel": mode, "coupling_planes": 1}
        tr = engine.sweep(ns, np.array([2.0]), seed=5, sweep_base=3, site_order="random",
                          energy_trace=True, **kw).cpu().numpy()
        outs.append((engine.spins().cpu().numpy(), tr, engine.accepted().cpu().numpy(),
                     engine.best()[0].cpu().numpy(), engine.best()[1].cpu().numpy()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("which", ["cfg2_ea_L256", "cfg5_sched_500x100"])
def test_csr_full_size_properties(engine, which):
    """BASELINE cfg2 / cfg5 at full size (fewer replicas): incremental fields and energies equal an
    exact recomputation, the best configuration has the best energy, energies went down."""
    import torch
    if which.startswith("cfg2"):
        rowptr, colidx, val, h = inst.ea_lattice(256)
        temps, exact = np.array([1.0]), True
    else:
        rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
        temps, exact = np.array([40.0]), False   # h is not integer-valued: fp32 incremental adds round
    n, R = h.shape[0], 96
    engine.set_model_csr(rowptr, colidx, val, h)
    engine.alloc_replicas(R)
    g = torch.Generator(device="cuda").manual_seed(3)
    engine.set_spins((torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8))
    engine.init_fields()
    e_start = engine.energies().double()
    engine.sweep(3, temps, seed=11, site_order="random")
    e_inc, f_inc = engine.energies().double(), engine.fields().double()
    acc = engine.accepted().sum().item()
    assert 0 < acc < 3 * n * R
    e_ref, f_ref = engine.batch_energies(engine.spins(), want_fields=True)
    if exact:
        assert torch.equal(f_inc, f_ref.double()) and torch.equal(e_inc, e_ref.double())
    else:
        assert (f_inc - f_ref.double()).abs().max().item() < 2e-2      # |f| ~ 5e3, ulp 5e-4
        assert ((e_inc - e_ref.double()).abs() / e_ref.double().abs()).max().item() < 1e-5
    assert (e_ref.double() < e_start).all()
    best_e, best_s = engine.best()
    eb = engine.batch_energies(best_s).double()
    assert ((eb - best_e.double()).abs() <= 1e-5 * eb.abs() + (0 if exact else 1.0)).all()
    assert (best_e.double() <= e_start + 1e-3).all()
