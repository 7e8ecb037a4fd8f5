"""GPU parity: the CUDA sweep path (through the C ABI) against the CPU oracle and the
golden traces recorded from the reference.

Bar: bit-exact for integer couplings (spins, energies, best configuration); for float
couplings the same spin trajectory and energies within 1e-5 relative (north_star).
"""
import numpy as np
import pytest

from conftest import golden_names, has_cuda, load_golden

pytestmark = pytest.mark.gpu

REL = 1e-5


@pytest.fixture(scope="module")
def engine():
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    from spin_glass_anneal_rl_b200.engine import Engine
    return Engine(0)


def _is_integer(J, h):
    return bool(np.all(J == np.round(J)) and np.all(h == np.round(h)))


def _close(a, b, exact, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if exact:
        assert np.array_equal(a, b), what
    else:
        assert np.allclose(a, b, rtol=REL, atol=REL), what


def _rand_instance(rng, n, integer):
    if integer:
        a = rng.integers(-2, 3, size=(n, n))
        J = np.triu(a, 1)
        J = (J + J.T).astype(np.float32)
        h = rng.integers(-2, 3, size=n).astype(np.float32)
    else:
        a = rng.standard_normal((n, n)) / np.sqrt(n)
        J = ((a + a.T) / 2).astype(np.float32)
        np.fill_diagonal(J, 0.0)
        h = (0.3 * rng.standard_normal(n)).astype(np.float32)
    return J, h


# ------------------------------------------------------------------ K2: fields + energies
@pytest.mark.parametrize("n,R,integer", [(48, 5, True), (100, 33, False), (1024, 70, True),
                                         (1500, 9, False), (4096, 16, True)])
def test_fields_and_energies(engine, oracle, n, R, integer):
    rng = np.random.default_rng(n * 7 + R)
    J, h = _rand_instance(rng, n, integer)
    S = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    engine.set_model(J, h)
    engine.alloc_replicas(R)
    engine.set_spins(S)
    engine.init_fields()
    F = engine.fields().cpu().numpy()
    E = engine.energies().cpu().numpy()
    assert np.array_equal(engine.spins().cpu().numpy(), S)
    Fo, Eo = oracle.batch_fields_energies(J, h, S.astype(np.float32))
    if integer:
        assert np.array_equal(F.astype(np.float64), Fo)
        assert np.array_equal(E.astype(np.float64), Eo)
    else:
        assert np.allclose(F, Fo, rtol=1e-5, atol=1e-5)
        assert np.allclose(E, Eo, rtol=REL, atol=1e-4)
    # stand-alone batched API (process_batch_energies / vectorized_local_fields)
    e2, f2 = engine.batch_energies(S, want_fields=True)
    assert np.array_equal(e2.cpu().numpy(), E) and np.array_equal(f2.cpu().numpy(), F)


def test_asymmetric_couplings_follow_reference_rows(engine, oracle):
    """Local field of spin i uses ROW i of J, as IsingModel.get_local_field does."""
    rng = np.random.default_rng(5)
    n = 40
    J = rng.integers(-3, 4, size=(n, n)).astype(np.float32)  # not symmetric, non-zero diagonal
    h = rng.integers(-1, 2, size=n).astype(np.float32)
    S = (rng.integers(0, 2, size=(3, n)) * 2 - 1).astype(np.int8)
    engine.set_model(J, h)
    engine.alloc_replicas(3)
    engine.set_spins(S)
    engine.init_fields()
    F = engine.fields().cpu().numpy()
    for b in range(3):
        for i in range(n):
            assert F[b, i] == oracle.local_field(J, h, S[b].astype(np.float32), i)


# ------------------------------------------------------------------ K1: replay of reference traces
def _replay_golden(engine, oracle, name):
    g = load_golden(name)
    c = g["config"]
    J, h = g["J"], g["h"]
    n = J.shape[0]
    stream = oracle.RawStream(oracle.mt_raw_stream(c["seed"], 2 * n * c["n_sweeps"] + 16))
    ores = oracle.anneal(J, h, g["spins0"], n_sweeps=c["n_sweeps"], T0=c["T0"], Tf=c["Tf"],
                         schedule=c["schedule"], schedule_params=c["params"],
                         record_interval=c["record_interval"], energy_tolerance=c["tol"],
                         rule=c["rule"], stream=stream, trace=True)
    ns = ores.n_sweeps
    engine.set_model(J, h)
    engine.alloc_replicas(1)
    engine.set_spins(g["spins0"].reshape(1, n))
    engine.init_fields()
    uni = np.nan_to_num(ores.extra["uniforms"], nan=0.5).astype(np.float32)
    trace = engine.sweep(ns, ores.extra["temps"], temps_sweep_stride=1, rule=c["rule"],
                         sites=ores.extra["sites"].astype(np.int32), uniforms=uni,
                         energy_trace=True, track_best=True, replicas_per_block=1)
    return g, ores, trace.cpu().numpy()[:, 0]


@pytest.mark.parametrize("name", golden_names("sa_"))
def test_sweep_replays_reference_trace(engine, oracle, name):
    g, ores, trace = _replay_golden(engine, oracle, name)
    exact = _is_integer(g["J"], g["h"])
    final = engine.spins().cpu().numpy()[0]
    best_e, best_s = engine.best()
    # the reference's own outputs (golden) ...
    assert np.array_equal(final, g["final_spins"]), "final spins differ from the reference"
    assert np.array_equal(best_s.cpu().numpy()[0], g["best_configuration"])
    _close(best_e.cpu().numpy()[0], g["best_energy"], exact, "best energy")
    # ... and the oracle's per-sweep view of the same run
    _close(trace, ores.sweep_energies, exact, "per-sweep energies")
    acc = int(engine.accepted().cpu().numpy()[0])
    n_att = ores.n_sweeps * g["J"].shape[0]
    assert abs(acc / n_att - g["acceptance_rate_history"][-1]) < 1e-12 or \
        (ores.n_sweeps - 1) % g["config"]["record_interval"] != 0


@pytest.mark.parametrize("n,G,integer,rule", [
    (64, 32, True, "metropolis"), (200, 7, False, "metropolis"), (1100, 24, True, "metropolis"),
    (1100, 5, False, "glauber"), (2100, 12, True, "metropolis"), (2100, 3, True, "heat_bath"),
    (4200, 6, True, "metropolis"),
])
def test_shared_site_order_many_replicas(engine, oracle, n, G, integer, rule):
    """G replicas per block share one site order but have their own spins and uniforms:
    every replica must follow the oracle's trajectory for (sites, its uniforms)."""
    rng = np.random.default_rng(n + G)
    J, h = _rand_instance(rng, n, integer)
    R = 2 * G + 1  # three blocks, the last one ragged
    ns = 3
    temps = np.array([2.5, 1.5, 0.8])
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)  # one order for every block
    uni = rng.random((R, ns, n), dtype=np.float32)
    engine.set_model(J, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    trace = engine.sweep(ns, temps, temps_sweep_stride=1, rule=rule, sites=sites,
                         sites_block_stride=0, uniforms=uni, energy_trace=True,
                         replicas_per_block=G).cpu().numpy()
    final = engine.spins().cpu().numpy()
    acc = engine.accepted().cpu().numpy()
    best_e, best_s = engine.best()
    for r in range(R):
        s = S0[r].astype(np.float32).copy()
        e0 = oracle.energy(J, h, s)
        es, ac = oracle.sweeps_scheduled(J, h, s, temps, rule, sites, uni[r])
        assert np.array_equal(final[r], s.astype(np.int8)), f"replica {r} trajectory differs"
        _close(trace[:, r], es, integer, f"replica {r} energies")
        assert int(acc[r]) == int(ac.sum())
        _close(best_e.cpu().numpy()[r], min(e0, es.min()), integer, "best energy")
    # best configuration really has the best energy
    eb = engine.batch_energies(best_s).cpu().numpy()
    _close(eb, best_e.cpu().numpy(), integer)


def test_launch_chunking_is_invisible(engine):
    """Philox counters are absolute: 6 sweeps in one launch == 2 + 4 sweeps in two."""
    rng = np.random.default_rng(3)
    n, R = 300, 50
    J, h = _rand_instance(rng, n, True)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    temps = np.linspace(3.0, 0.5, 6)
    outs = []
    for chunks in ([6], [2, 4], [1, 1, 1, 3]):
        engine.set_model(J, h)
        engine.alloc_replicas(R)
        engine.set_spins(S0)
        engine.init_fields()
        base = 0
        for c in chunks:
            engine.sweep(c, temps[base:base + c].copy(), temps_sweep_stride=1, seed=99,
                         sweep_base=base, site_order="random")
            base += c
        outs.append((engine.spins().cpu().numpy(), engine.energies().cpu().numpy(),
                     engine.best_energies().cpu().numpy(), engine.accepted().cpu().numpy()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    assert outs[0][3].sum() > 0


def test_philox_energy_consistency(engine):
    """After Philox-mode sweeps the incrementally maintained fields/energies equal a fresh
    evaluation from the final spins (integer couplings: exactly)."""
    rng = np.random.default_rng(11)
    n, R = 1500, 40
    J, h = _rand_instance(rng, n, True)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    engine.set_model(J, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    engine.sweep(5, np.array([4.0]), site_order="sequential", seed=5)
    e = engine.energies().cpu().numpy()
    f = engine.fields().cpu().numpy()
    s = engine.spins().cpu().numpy()
    e2, f2 = engine.batch_energies(s, want_fields=True)
    assert np.array_equal(e, e2.cpu().numpy()) and np.array_equal(f, f2.cpu().numpy())
    assert not np.array_equal(s, S0)


def test_philox_matches_reference_statistics(engine, oracle):
    """Philox mode vs the reference algorithm (oracle, mt19937 streams) at fixed T:
    equilibrium mean energy and acceptance rate agree within sampling error."""
    rng = np.random.default_rng(21)
    n = 64
    J, h = _rand_instance(rng, n, True)
    T, warm, meas = 3.0, 60, 60
    # oracle: 48 independent chains
    eo, ao = [], []
    for seed in range(48):
        st = oracle.RawStream(oracle.mt_raw_stream(1000 + seed, 2 * n * (warm + meas) + 8))
        s = oracle.raw_to_spins(st.take(n)).copy()
        es, ac, _, _ = oracle.sweeps(J, h, s, [T] * (warm + meas), "metropolis", st)
        eo.append(es[warm:].mean())
        ao.append(ac[warm:].sum() / (meas * n))
    R = 512
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    engine.set_model(J, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    engine.sweep(warm, np.array([T]), seed=7, site_order="random")
    a0 = engine.accepted().cpu().numpy().copy()
    tr = engine.sweep(meas, np.array([T]), seed=7, sweep_base=warm, site_order="random",
                      energy_trace=True).cpu().numpy()
    a1 = engine.accepted().cpu().numpy()
    eg = tr.mean(axis=0)
    ag = (a1 - a0) / (meas * n)
    se = np.sqrt(np.var(eo) / len(eo) + np.var(eg) / len(eg))
    assert abs(np.mean(eo) - np.mean(eg)) < 5 * se + 1e-9, (np.mean(eo), np.mean(eg), se)
    sa = np.sqrt(np.var(ao) / len(ao) + np.var(ag) / len(ag))
    assert abs(np.mean(ao) - np.mean(ag)) < 5 * sa + 1e-9, (np.mean(ao), np.mean(ag), sa)


# ------------------------------------------------------------------ K3: exchange
def test_exchange_matches_reference_rule(engine):
    """sg_exchange with injected uniforms == the reference's nearest-neighbour rule."""
    rng = np.random.default_rng(8)
    n, K, Lad = 32, 6, 3
    R = K * Lad
    J, h = _rand_instance(rng, n, True)
    S = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    ladder = np.geomspace(5.0, 0.4, K)
    engine.set_model(J, h)
    engine.alloc_replicas(R)
    engine.set_spins(S)
    engine.init_fields()
    engine.set_ladder(list(ladder))
    E = engine.energies().cpu().numpy().astype(np.float64)
    rep_at = np.arange(R).reshape(Lad, K).copy()
    att = np.zeros((Lad, K - 1), np.int64)
    acc = np.zeros((Lad, K - 1), np.int64)
    for rnd, parity in enumerate([0, 1, 1, 0, 1]):
        u = rng.random((Lad, K // 2))
        engine.exchange(parity, uniforms=u)
        for l in range(Lad):
            for m, k in enumerate(range(parity, K - 1, 2)):
                ra, rb = rep_at[l, k], rep_at[l, k + 1]
                p = min(1.0, np.exp((1 / ladder[k + 1] - 1 / ladder[k]) * (E[rb] - E[ra])))
                att[l, k] += 1
                if u[l, m] < p:
                    rep_at[l, k], rep_at[l, k + 1] = rb, ra
                    acc[l, k] += 1
    g_rep, g_T, g_att, g_acc = [t.cpu().numpy() for t in engine.ladder_state()]
    assert np.array_equal(g_rep.reshape(Lad, K), rep_at)
    assert np.array_equal(g_att, att) and np.array_equal(g_acc, acc)
    want_T = np.empty(R)
    for l in range(Lad):
        for k in range(K):
            want_T[rep_at[l, k]] = ladder[k]
    assert np.array_equal(g_T, want_T)
    assert acc.sum() > 0
