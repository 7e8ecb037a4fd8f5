"""GPU parity of the checkerboard multi-spin-coded lattice kernel (BASELINE cfg2) through the
C ABI.  The checkerboard sequence is a legal explicit site order of the sequential algorithm
(same-colour sites do not interact), so replay with injected uniforms against the oracle is
bit-exact; in Philox mode the kernel is compared statistically with the sparse kernel
(random site order) and checked for exact energy bookkeeping at full size."""
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


@pytest.mark.parametrize("L,periodic,rule", [(8, False, "metropolis"), (7, False, "metropolis"),
                                             (6, True, "metropolis"), (12, False, "glauber"),
                                             (8, True, "heat_bath")])
def test_lattice_replay_is_bit_exact(engine, oracle, L, periodic, rule):
    Jx, Jy = inst.ea_lattice_bonds(L, seed=L, periodic=periodic)
    rowptr, colidx, val, h = inst.lattice_csr(Jx, Jy)
    n = L * L
    J = inst.csr_to_dense(rowptr, colidx, val, n)
    seq = inst.checkerboard_sequence(L)
    for (x, y) in [(0, 0), (1, 0), (L - 1, L - 1), (2, 3)]:
        assert seq[engine._lib.sg_lattice_sequence_index(L, x, y)] == x * L + y
    rng = np.random.default_rng(L)
    R, ns = 45, 3
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    uni = rng.random((R, ns, n), dtype=np.float32)
    temps = np.array([2.5, 1.2, 0.6])
    engine.set_model_lattice2d(Jx, Jy)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    assert np.array_equal(engine.spins().cpu().numpy(), S0)
    _, Eo = oracle.batch_fields_energies(J, h, S0.astype(np.float32))
    assert np.array_equal(engine.energies().cpu().numpy().astype(np.float64), Eo)
    trace = engine.sweep(ns, temps, temps_sweep_stride=1, rule=rule, site_order="checkerboard",
                         uniforms=uni, energy_trace=True).cpu().numpy()
    final = engine.spins().cpu().numpy()
    acc = engine.accepted().cpu().numpy()
    best_e, best_s = engine.best()
    sites = np.tile(seq, (ns, 1))
    for r in range(R):
        s = S0[r].astype(np.float32).copy()
        e0 = oracle.energy(J, h, s)
        es, ac = oracle.sweeps_scheduled(J, h, s, temps, rule, sites, uni[r])
        assert np.array_equal(final[r], s.astype(np.int8)), f"replica {r} trajectory differs"
        assert np.array_equal(trace[:, r].astype(np.float64), es)
        assert int(acc[r]) == int(ac.sum())
        assert float(best_e[r]) == min(e0, es.min())
    assert np.array_equal(engine.batch_energies(best_s).cpu().numpy(), best_e.cpu().numpy())


def test_lattice_philox_statistics_match_sparse_kernel(engine):
    """Equilibrium energy per spin at T = 1.2 on L = 32: checkerboard multi-spin kernel vs the
    sparse kernel with random site order (different dynamics, same stationary distribution)."""
    import torch
    L, R, T = 32, 256, 1.2
    Jx, Jy = inst.ea_lattice_bonds(L, seed=5)
    n = L * L
    g = torch.Generator(device="cuda").manual_seed(1)
    S0 = (torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8)
    res = {}
    for mode in ("lattice", "csr"):
        if mode == "lattice":
            engine.set_model_lattice2d(Jx, Jy)
        else:
            engine.set_model_csr(*inst.lattice_csr(Jx, Jy))
        engine.alloc_replicas(R)
        engine.set_spins(S0)
        engine.init_fields()
        order = "checkerboard" if mode == "lattice" else "random"
        engine.sweep(150, np.array([T]), seed=9, site_order=order, track_best=False)
        a0 = engine.accepted().sum().item()
        engine.sweep(50, np.array([T]), seed=9, sweep_base=150, site_order=order, track_best=False)
        acc = (engine.accepted().sum().item() - a0) / (50 * n * R)
        e = engine.batch_energies(engine.spins()).double() / n
        res[mode] = (e.mean().item(), e.std().item() / np.sqrt(R), acc)
    (ml, sl, al), (mc, sc, ac) = res["lattice"], res["csr"]
    assert abs(ml - mc) < 4.5 * np.hypot(sl, sc) + 2e-3, res
    assert abs(al - ac) < 0.01, res


def test_lattice_full_size_cfg2(engine):
    """L = 256, 4096 replicas, a temperature ladder inside every word: exact energy bookkeeping."""
    import torch
    L, R = 256, 4096
    Jx, Jy = inst.ea_lattice_bonds(L)
    n = L * L
    engine.set_model_lattice2d(Jx, Jy)
    engine.alloc_replicas(R)
    g = torch.Generator(device="cuda").manual_seed(2)
    engine.set_spins((torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8))
    engine.init_fields()
    e0 = engine.energies().double()
    temps = np.tile(np.geomspace(3.0, 0.1, 32), R // 32)            # per-replica temperatures
    engine.sweep(4, temps, temps_replica_stride=1, seed=3, site_order="checkerboard")
    e1 = engine.energies().double()
    assert torch.equal(e1, engine.batch_energies(engine.spins()).double())
    assert (e1 < e0).all()
    best_e, best_s = engine.best()
    assert torch.equal(engine.batch_energies(best_s).double(), best_e.double())
    assert (best_e.double() <= e1).all()
    acc = engine.accepted().double().reshape(-1, 32).mean(0) / (4 * n)
    assert acc[0] > acc[-1] and 0 < acc[-1] < 0.3 and acc[0] > 0.4   # hot rungs accept more
    with pytest.raises(Exception):
        engine.sweep(1, np.array([1.0]), site_order="random")
    with pytest.raises(Exception):
        engine.fields()
