"""IsingModel coupling construction (host logic, CPU).

The reference's problem encoders build their models with one ``set_coupling`` call per pair
(reference problems/routing.py:275-294, core/constraints.py:360-388) and read ``couplings``
afterwards; on a sparse model every call there is a dense round trip (core/ising_model.py:94-99).
Here the calls are queued and folded in on the next read: same resulting matrix, O(1) per call."""
import time

import numpy as np
import pytest
import torch

from spin_glass_anneal_rl_b200.core.ising_model import IsingModel, IsingModelConfig


def _pair(n):
    return (IsingModel(IsingModelConfig(n_spins=n, use_sparse=True)),
            IsingModel(IsingModelConfig(n_spins=n, use_sparse=False)))


def test_queued_set_coupling_equals_dense_model():
    rng = np.random.default_rng(0)
    n = 40
    ms, md = _pair(n)
    for k in range(2500):
        i, j = (int(x) for x in rng.integers(0, n, 2))
        v = float(rng.integers(-2, 3))          # zeros delete entries
        ms.set_coupling(i, j, v)
        md.set_coupling(i, j, v)
        if k % 311 == 0:
            assert ms.get_coupling(i, j) == md.get_coupling(i, j) == v
            assert ms.get_coupling(j, i) == v
            assert torch.equal(ms.couplings.to_dense(), md.couplings)
    J = ms.couplings
    assert J.is_sparse and torch.equal(J.to_dense(), md.couplings)
    assert bool((J.coalesce().values() != 0).all())        # no stored zeros, like to_sparse()
    assert torch.equal(J.to_dense(), J.to_dense().T)
    ms.spins = md.spins.clone()
    assert ms.compute_energy() == pytest.approx(md.compute_energy(), abs=1e-5)
    for i in (0, 7, n - 1):
        assert ms.get_local_field(i) == pytest.approx(md.get_local_field(i), abs=1e-5)
        assert ms.get_coupling(i, (i + 3) % n) == md.get_coupling(i, (i + 3) % n)


def test_assigning_couplings_discards_queued_writes_and_copy_sees_them():
    ms, _ = _pair(6)
    ms.set_coupling(0, 1, 2.0)
    c = ms.copy()                                 # reads the attribute: the write is in the copy
    assert c.get_coupling(1, 0) == 2.0
    ms.set_coupling(2, 3, 1.0)
    ms.couplings = torch.zeros(6, 6)              # public attribute assignment wins
    assert float(ms.couplings.abs().sum()) == 0.0
    ms.set_coupling(4, 5, -1.0)                   # dense now: written in place
    assert ms.couplings[5, 4] == -1.0
    with pytest.raises(ValueError):
        ms.set_coupling(0, 6, 1.0)
    with pytest.raises(ValueError):
        ms.get_coupling(-1, 0)


def test_energy_cache_invalidated_by_queued_write():
    ms, _ = _pair(5)
    ms.spins = torch.ones(5)
    e0 = ms.compute_energy()
    ms.set_coupling(0, 1, 3.0)
    assert ms.compute_energy() == pytest.approx(e0 - 3.0)


def test_encoder_call_pattern_builds_tsp_model_quickly():
    """The position-encoded TSP of the reference's encoder, written pair by pair through
    set_coupling on a sparse model (12 cities = 144 spins, ~4.5 k pairs), equals the vectorised
    construction used for the benchmark instances; 64 cities' worth of calls stays in seconds."""
    from tools.instances import cardinality_terms, random_tsp, tsp_ising
    xy = random_tsp(12, seed=11)
    n = xy.shape[0]
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    field, coupling = cardinality_terms(n, 1, 100.0)
    m = IsingModel(IsingModelConfig(n_spins=n * n, use_sparse=True))
    for c in range(n):
        for p in range(n):
            m.set_external_field(c * n + p, 2.0 * field)
            for q in range(p + 1, n):
                m.set_coupling(c * n + p, c * n + q, coupling)      # one position per city
                m.set_coupling(p * n + c, q * n + c, coupling)      # one city per position
    for p in range(n):
        q = (p + 1) % n
        for c in range(n):
            for c2 in range(n):
                if c != c2:
                    m.set_coupling(c * n + p, c2 * n + q, -float(d[c, c2]))
    J, h = tsp_ising(xy, penalty=100.0)
    assert np.allclose(m.couplings.to_dense().numpy(), J, atol=1e-5)
    assert np.allclose(m.external_fields.numpy(), h)

    big = IsingModel(IsingModelConfig(n_spins=4096, use_sparse=True))
    t0 = time.time()
    for k in range(300_000):
        big.set_coupling(k % 4096, (k * 13 + 5) % 4096, 1.0 + (k & 3))
    nnz = big.couplings._nnz()
    assert nnz > 0 and time.time() - t0 < 30.0
