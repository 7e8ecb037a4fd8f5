"""Stacked small models (sg_set_model_dense_batch + the small-model kernel) and the
BatchProcessor mirror built on them (reference annealing/batch_processor.py:180-288)."""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    from spin_glass_anneal_rl_b200.engine import Engine
    return Engine(0)


def _int_models(rng, M, n):
    a = rng.integers(-2, 3, size=(M, n, n))
    J = np.triu(a, 1)
    J = (J + J.transpose(0, 2, 1)).astype(np.float32)
    h = rng.integers(-2, 3, size=(M, n)).astype(np.float32)
    return J, h


@pytest.mark.parametrize("M,n,r", [(3, 20, 5), (7, 100, 8), (2, 224, 3), (16, 33, 1)])
def test_stacked_models_equal_one_engine_per_model(engine, oracle, M, n, r):
    """Every model of a stack must evolve exactly as it does alone: same Philox seed, the replica
    index inside the stack is the Philox replica id, so model k alone is run with its replicas
    placed at the same global indices (a dummy prefix), and against the oracle in replay mode."""
    rng = np.random.default_rng(M * n + r)
    J, h = _int_models(rng, M, n)
    S0 = (rng.integers(0, 2, size=(M * r, n)) * 2 - 1).astype(np.int8)
    ns = 4
    temps = np.linspace(2.5, 0.6, ns)
    sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
    uni = rng.random((M * r, ns, n), dtype=np.float32)

    engine.set_models(J, h)
    engine.alloc_replicas(M * r)
    engine.set_spins(S0)
    engine.init_fields()
    E0 = engine.energies().cpu().numpy()
    trace = engine.sweep(ns, temps, temps_sweep_stride=1, sites=sites, uniforms=uni,
                         energy_trace=True).cpu().numpy()
    final = engine.spins().cpu().numpy()
    acc = engine.accepted().cpu().numpy()
    best_e, best_s = engine.best()
    fields = engine.fields().cpu().numpy()
    for k in range(M * r):
        m = k // r
        s = S0[k].astype(np.float32).copy()
        assert float(E0[k]) == oracle.energy(J[m], h[m], s)
        es, ac = oracle.sweeps_scheduled(J[m], h[m], s, temps, "metropolis", sites, uni[k])
        assert np.array_equal(final[k], s.astype(np.int8)), f"replica {k} (model {m}) differs"
        assert np.array_equal(trace[:, k].astype(np.float64), es)
        assert int(acc[k]) == int(ac.sum())
        assert float(best_e[k]) == min(float(E0[k]), es.min())
        Fo, _ = oracle.batch_fields_energies(J[m], h[m], final[k:k + 1].astype(np.float32))
        assert np.array_equal(fields[k].astype(np.float64), Fo[0])
    assert np.array_equal(engine.batch_energies(best_s).cpu().numpy(), best_e.cpu().numpy())

    # Philox mode: the stack against each model in an engine of its own
    engine.set_models(J, h)
    engine.alloc_replicas(M * r)
    engine.set_spins(S0)
    engine.init_fields()
    engine.sweep(ns, temps, temps_sweep_stride=1, seed=31, site_order="random")
    stacked = engine.spins().cpu().numpy()
    from spin_glass_anneal_rl_b200.engine import Engine
    solo = Engine(0)
    for m in (0, M - 1):
        # the same replicas at the same global indices: pad with the preceding replicas' slots
        solo.set_model(J[m], h[m])
        solo.alloc_replicas((m + 1) * r)
        pad = np.ones(((m + 1) * r, n), np.int8)
        pad[m * r:] = S0[m * r:(m + 1) * r]
        solo.set_spins(pad)
        solo.init_fields()
        solo.sweep(ns, temps, temps_sweep_stride=1, seed=31, site_order="random", kernel="small")
        assert np.array_equal(solo.spins().cpu().numpy()[m * r:], stacked[m * r:(m + 1) * r])


def test_stacked_models_errors(engine):
    from spin_glass_anneal_rl_b200._lib import SGError
    rng = np.random.default_rng(1)
    J, h = _int_models(rng, 3, 16)
    engine.set_models(J, h)
    with pytest.raises(SGError):
        engine.alloc_replicas(7)                 # not a multiple of the number of models
    engine.alloc_replicas(6)
    engine.set_spins(np.ones((6, 16), np.int8))
    engine.init_fields()
    with pytest.raises(SGError):
        engine.sweep(1, np.array([1.0]), kernel="tc")
    with pytest.raises(SGError):
        engine.sweep(1, np.array([1.0]), site_order="random_per_block")
    Jb, hb = _int_models(rng, 2, 300)
    with pytest.raises(SGError):
        engine.set_models(Jb, hb)                # n > 224
    engine.set_model(J[0], h[0])                 # back to a single model
    engine.alloc_replicas(5)
    engine.set_spins(np.ones((5, 16), np.int8))
    engine.init_fields()
    engine.sweep(1, np.array([1.0]))


def test_batch_processor_stacks_small_models(oracle):
    """process_models_batch: results in input order, best energies exact, every model at least
    as good as its start, mixed sizes (stacked groups + a large model on its own)."""
    import torch
    import spin_glass_anneal_rl_b200 as sg
    from spin_glass_anneal_rl_b200.annealing.batch_processor import BatchConfig, BatchProcessor, plan_stacks
    from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType

    assert plan_stacks([10, 20, 10, 300, 10, 20], [True, True, True, False, True, True], 2) == \
        [[0, 2], [1, 5], [3], [4]]
    rng = np.random.default_rng(9)
    models, Js, hs = [], [], []
    for n in (24, 40, 24, 260, 24, 40):
        a = rng.normal(size=(n, n)).astype(np.float32)
        J = np.triu(a, 1)
        J = J + J.T
        h = (0.3 * rng.normal(size=n)).astype(np.float32)
        m = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=(n == 40)))
        m.set_couplings_from_matrix(torch.from_numpy(J))
        m.set_external_fields(torch.from_numpy(h))
        models.append(m)
        Js.append(J)
        hs.append(h)
    start = [m.compute_energy() for m in models]
    cfg = sg.GPUAnnealerConfig(n_sweeps=300, initial_temp=3.0, final_temp=0.05,
                               schedule_type=ScheduleType.GEOMETRIC, schedule_params={"alpha": 0.985},
                               record_interval=50, n_replicas=16, random_seed=3)
    bp = BatchProcessor(BatchConfig(batch_size=8), cfg)
    seen = []
    results = bp.process_models_batch(models, callback=lambda rs: seen.append(len(rs)))
    assert seen == [6] and len(results) == 6 and bp.get_processing_stats()["processed_batches"] == 1
    for m, J, h, res, e_start in zip(models, Js, hs, results, start):
        n = m.n_spins
        assert res.best_configuration.shape == (n,) and res.n_sweeps == 300
        assert oracle.energy(J, h, res.best_configuration.numpy()) == pytest.approx(res.best_energy, rel=1e-5, abs=1e-4)
        assert res.best_energy <= e_start + 1e-4
        assert len(res.energy_history) == len(res.temperature_history) == len(res.acceptance_rate_history)
        assert set(np.unique(m.spins.numpy())) <= {-1.0, 1.0}
    # the same three 24-spin models alone reach energies of the same quality (not identical runs)
    solo = sg.GPUAnnealer(cfg).anneal(models[0])
    assert abs(solo.best_energy - results[0].best_energy) <= 0.1 * abs(solo.best_energy)
