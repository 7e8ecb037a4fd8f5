"""GPU tests of the reference-shaped Python API (GPUAnnealer / ParallelTempering /
SpinDynamics / batch energies), checked against the oracle and the golden fixtures."""
import numpy as np
import pytest
import torch

from conftest import golden_names, has_cuda, load_golden

pytestmark = pytest.mark.gpu

import spin_glass_anneal_rl_b200 as sg
from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType
from spin_glass_anneal_rl_b200.core.spin_dynamics import UpdateRule


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_cuda():
        pytest.skip("needs a CUDA device")


def _model(J, h, spins=None, sparse=False):
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=J.shape[0], use_sparse=sparse))
    m.set_couplings_from_matrix(torch.from_numpy(np.asarray(J, np.float32)))
    m.set_external_fields(torch.from_numpy(np.asarray(h, np.float32)))
    if spins is not None:
        m.set_spins(torch.from_numpy(np.asarray(spins, np.float32)))
    return m


def _cfg1():
    g = load_golden("sa_cfg1_float_n100")
    return g["J"], g["h"], g["spins0"]


def test_anneal_result_contract(oracle):
    J, h, s0 = _cfg1()
    m = _model(J, h, s0)
    e0 = m.compute_energy()
    ann = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=200, initial_temp=5.0, final_temp=0.01,
                                              random_seed=7))
    res = ann.anneal(m)
    assert isinstance(res, sg.AnnealingResult) and res.algorithm == "simulated_annealing"
    assert res.n_sweeps == 200 and res.total_time > 0
    assert len(res.energy_history) == len(res.temperature_history) == len(res.acceptance_rate_history) == 21
    assert res.energy_history[0] == pytest.approx(e0, rel=1e-5)
    assert res.temperature_history[0] == 5.0 and res.temperature_history[-1] < res.temperature_history[1]
    assert res.best_configuration.dtype == torch.float32 and res.best_configuration.device.type == "cpu"
    assert set(res.best_configuration.tolist()) <= {-1.0, 1.0}
    # the best configuration really has the reported energy, and it is a big improvement
    assert oracle.energy(J, h, res.best_configuration.numpy()) == pytest.approx(res.best_energy, rel=1e-5)
    assert res.best_energy <= min(res.energy_history) + 1e-4 and res.best_energy < e0 - 100
    # anneal() leaves the model at the final configuration (reference mutates model.spins)
    assert oracle.energy(J, h, m.spins.numpy()) == pytest.approx(res.energy_history[-1], rel=1e-4, abs=1e-3) \
        or res.n_sweeps % 10 != 1
    assert 0.0 < res.acceptance_rate_history[1] <= 1.0


def test_anneal_is_reproducible_and_seed_sensitive():
    J, h, s0 = _cfg1()
    out = []
    for seed in (11, 11, 12):
        m = _model(J, h, s0)
        r = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=60, initial_temp=3.0, final_temp=0.05,
                                                random_seed=seed)).anneal(m)
        out.append((r.best_energy, r.energy_history, r.best_configuration.clone()))
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1] and torch.equal(out[0][2], out[1][2])
    assert out[0][1] != out[2][1]


def test_anneal_matches_reference_statistics(oracle):
    """Same instance, schedule and budget as the reference run recorded in the golden file:
    the distribution of best energies over seeds agrees with the oracle's (reference
    algorithm, mt19937 streams) within sampling error."""
    g = load_golden("sa_cfg1_float_n100")
    c = g["config"]
    J, h, s0 = g["J"], g["h"], g["spins0"]
    ref = []
    for seed in range(24):
        st = oracle.RawStream(oracle.mt_raw_stream(500 + seed, 2 * 100 * 120 + 16))
        ref.append(oracle.anneal(J, h, s0, n_sweeps=120, T0=c["T0"], Tf=c["Tf"], schedule="geometric",
                                 schedule_params=c["params"], record_interval=10, stream=st).best_energy)
    m = _model(J, h, s0)
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(
        n_sweeps=120, initial_temp=c["T0"], final_temp=c["Tf"], schedule_params=c["params"],
        random_seed=1, n_replicas=256)).anneal(m)
    eng = m._sg_engine[1]
    mine = eng.best_energies().cpu().numpy()
    # replica 0 starts from s0 like the oracle runs; the others from random spins (the start is
    # forgotten at T0 = 5): compare means with a 5-sigma band
    se = np.sqrt(np.var(ref) / len(ref) + np.var(mine) / len(mine))
    assert abs(np.mean(ref) - np.mean(mine)) < 5 * se + 1e-6, (np.mean(ref), np.mean(mine), se)
    assert res.best_energy == pytest.approx(mine.min(), rel=1e-6)
    assert res.best_energy <= np.mean(ref)


@pytest.mark.parametrize("sched,params", [(ScheduleType.LINEAR, {}), (ScheduleType.EXPONENTIAL, {}),
                                          (ScheduleType.LOGARITHMIC, {"c": 2.0}),
                                          (ScheduleType.POWER_LAW, {"k": 0.7}), (ScheduleType.FAST, {}),
                                          (ScheduleType.BOLTZMANN, {}),
                                          (ScheduleType.ADAPTIVE, {"alpha": 0.9, "adaptation_window": 5,
                                                                   "target_acceptance": 0.3})])
def test_every_schedule_type_runs(sched, params):
    g = load_golden("sa_sched_linear_n24")
    m = _model(g["J"], g["h"], g["spins0"])
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=40, initial_temp=4.0, final_temp=0.2,
                                              schedule_type=sched, schedule_params=params,
                                              record_interval=3, random_seed=3)).anneal(m)
    gold = load_golden(f"sa_sched_{sched.value}_n24")
    if sched is not ScheduleType.ADAPTIVE:
        assert np.allclose(res.temperature_history, gold["temperature_history"], rtol=1e-15, atol=0)
    assert len(res.energy_history) == len(gold["energy_history"])
    assert res.best_energy <= gold["energy_history"][0]
    assert res.best_energy == int(res.best_energy)  # integer couplings -> exact integer energies


@pytest.mark.parametrize("window,target", [(5, 0.3), (12, 0.6), (100, 0.44)])
def test_adaptive_schedule_on_device_equals_host_evaluation(monkeypatch, window, target):
    """ADAPTIVE (reference annealing/temperature_scheduler.py:206-249): the feedback step runs as a
    device kernel between sweeps; the same run with the schedule evaluated on the host after a
    read-back of the acceptance counter per sweep (SG_ADAPTIVE_HOST=1) must give the same
    temperatures, energies and best configuration."""
    g = load_golden("sa_sched_linear_n24")
    out = []
    for host in ("1", ""):
        if host:
            monkeypatch.setenv("SG_ADAPTIVE_HOST", host)
        else:
            monkeypatch.delenv("SG_ADAPTIVE_HOST", raising=False)
        m = _model(g["J"], g["h"], g["spins0"])
        cfg = sg.GPUAnnealerConfig(n_sweeps=150, initial_temp=4.0, final_temp=0.2,
                                   schedule_type=ScheduleType.ADAPTIVE,
                                   schedule_params={"alpha": 0.97, "adaptation_window": window,
                                                    "target_acceptance": target, "adaptation_rate": 0.2},
                                   record_interval=1, random_seed=3, energy_tolerance=0.0)
        out.append((sg.GPUAnnealer(cfg).anneal(m), m.spins.clone()))
    (a, sa), (b, sb) = out
    assert len(a.temperature_history) == len(b.temperature_history) == 151
    assert np.allclose(a.temperature_history, b.temperature_history, rtol=1e-14, atol=0)
    assert len(set(np.round(a.temperature_history, 12))) > 20         # the feedback really acts
    assert a.energy_history == b.energy_history
    assert a.acceptance_rate_history == b.acceptance_rate_history
    assert a.best_energy == b.best_energy and torch.equal(a.best_configuration, b.best_configuration)
    assert torch.equal(sa, sb)


@pytest.mark.parametrize("rule", [UpdateRule.GLAUBER, UpdateRule.HEAT_BATH])
def test_other_update_rules(rule, oracle):
    g = load_golden("sa_glauber_int_n32")
    m = _model(g["J"], g["h"], g["spins0"])
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=50, initial_temp=3.0, final_temp=0.3,
                                              random_seed=5, n_replicas=64)).anneal(m, rule)
    assert oracle.energy(g["J"], g["h"], res.best_configuration.numpy()) == res.best_energy
    assert res.best_energy <= float(g["best_energy"]) + 8


def test_early_stop_like_reference():
    g = load_golden("sa_converge_n12")
    c = g["config"]
    m = _model(g["J"], g["h"], g["spins0"])
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(
        n_sweeps=c["n_sweeps"], initial_temp=c["T0"], final_temp=c["Tf"], schedule_params=c["params"],
        record_interval=1, random_seed=3)).anneal(m)
    # frozen system: the relative-std test fires once 50 energies are recorded (sweep 48)
    assert res.n_sweeps < c["n_sweeps"] and res.n_sweeps >= 49
    assert res.best_energy == float(g["best_energy"])


def test_sparse_models_just_work(oracle):
    """Every ProblemTemplate builds use_sparse=True models; the reference fails on them."""
    rng = np.random.default_rng(0)
    n = 60
    a = rng.integers(-1, 2, size=(n, n)) * (rng.random((n, n)) < 0.1)
    J = np.triu(a, 1)
    J = (J + J.T).astype(np.float32)
    h = rng.integers(-1, 2, size=n).astype(np.float32)
    m = _model(J, h, sparse=True)
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=80, initial_temp=3.0, final_temp=0.05,
                                              random_seed=2, n_replicas=32)).anneal(m)
    assert oracle.energy(J, h, res.best_configuration.numpy()) == res.best_energy
    # the engine is cached and refreshed when the couplings change
    e1 = m._sg_engine[1]
    sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=5, random_seed=2)).anneal(m)
    assert m._sg_engine[1] is e1
    m.set_coupling(0, 1, 5.0)
    r2 = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=40, initial_temp=3.0, final_temp=0.05,
                                             random_seed=2)).anneal(m)
    J[0, 1] = J[1, 0] = 5.0
    assert oracle.energy(J, h, r2.best_configuration.numpy()) == r2.best_energy


def test_invalid_model_raises():
    with pytest.raises((AttributeError, TypeError)):
        sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=5)).anneal("not a model")


def test_parallel_tempering_contract(oracle):
    g = load_golden("pt_int_n40_r6")
    c = g["config"]
    J, h = g["J"], g["h"]
    m = _model(J, h)
    pt = sg.ParallelTempering(sg.ParallelTemperingConfig(
        n_replicas=6, n_sweeps=200, temp_min=c["tmin"], temp_max=c["tmax"], exchange_interval=5,
        record_interval=5, random_seed=21, n_ladders=8))
    res = pt.run(m)
    assert res.algorithm == "parallel_tempering" and res.n_sweeps == 200
    assert oracle.energy(J, h, res.best_configuration.numpy()) == res.best_energy
    assert res.best_energy <= float(g["best_energy"])  # 8 ladders x 200 sweeps vs 1 x 60
    assert len(pt.energy_histories) == 6 and len(pt.energy_histories[0]) == 40
    assert res.energy_history == pt.energy_histories[0]
    assert np.mean(pt.energy_histories[0]) > np.mean(pt.energy_histories[5])  # hot rung is higher
    # 39 exchange rounds x 8 ladders, alternating parity at random
    assert pt.exchange_attempts.shape == (5,) and pt.exchange_attempts.sum() > 39 * 8 * 2 - 1
    rates = pt.get_exchange_rates()
    assert np.all(rates >= 0) and np.all(rates <= 1) and rates.max() > 0
    assert len(res.acceptance_rate_history) == 6
    assert res.acceptance_rate_history[0] > res.acceptance_rate_history[-1]  # hot rung accepts more
    assert pt.get_statistics()["temperatures"] == pt.temperatures
    res2 = sg.ParallelTempering(sg.ParallelTemperingConfig(
        n_replicas=6, n_sweeps=50, temp_min=0.5, temp_max=6.0, random_seed=1)).anneal(m)
    assert res2.algorithm == "parallel_tempering"


def test_pt_exchange_rates_match_reference_statistics(oracle):
    """Exchange acceptance per rung pair vs the reference algorithm (oracle) on the same ladder."""
    g = load_golden("pt_int_n40_r6")
    c = g["config"]
    J, h = g["J"], g["h"]
    acc = np.zeros(5)
    att = np.zeros(5)
    for seed in range(6):
        st = oracle.RawStream(oracle.mt_raw_stream(900 + seed, 2 * 40 * 6 * 302 + 16))
        r = oracle.parallel_tempering(J, h, n_replicas=6, n_sweeps=300, temp_min=c["tmin"],
                                      temp_max=c["tmax"], exchange_interval=3, record_interval=50,
                                      stream=st, np_rng=np.random.RandomState(900 + seed))
        acc += r.extra["exchange_accepts"]
        att += r.extra["exchange_attempts"]
    pt = sg.ParallelTempering(sg.ParallelTemperingConfig(
        n_replicas=6, n_sweeps=300, temp_min=c["tmin"], temp_max=c["tmax"], exchange_interval=3,
        record_interval=50, random_seed=4, n_ladders=32))
    pt.run(_model(J, h))
    mine, ref = pt.get_exchange_rates(), acc / att
    se = np.sqrt(ref * (1 - ref) / att + mine * (1 - mine) / pt.exchange_attempts)
    assert np.all(np.abs(mine - ref) < 5 * se + 0.02), (mine, ref, se)


def test_readme_shaped_anneal_and_batch_api(oracle):
    J, h, s0 = _cfg1()
    m = _model(J, h, s0)
    res = sg.anneal(m, n_replicas=64, n_sweeps=100, beta_schedule="geometric", initial_temp=5.0,
                    final_temp=0.05, random_seed=9)
    assert oracle.energy(J, h, res.best_configuration.numpy()) == pytest.approx(res.best_energy, rel=1e-5)
    betas = np.linspace(0.2, 20.0, 50)
    res2 = sg.anneal(m, n_replicas=8, n_sweeps=50, beta_schedule=betas, random_seed=9)
    assert np.allclose(res2.temperature_history[1:], (1.0 / betas)[::1][:len(res2.temperature_history) - 1])
    for name in golden_names("kat_"):
        k = load_golden(name)
        S = torch.from_numpy(k["S"].astype(np.float32))
        E = sg.batch_energies(S, torch.from_numpy(k["J"]), torch.from_numpy(k["h"])).cpu().numpy()
        F = sg.batch_local_fields(S, torch.from_numpy(k["J"]), torch.from_numpy(k["h"])).cpu().numpy()
        assert np.allclose(E, k["E"], rtol=1e-5, atol=1e-4) and np.allclose(F, k["F"], rtol=1e-5, atol=1e-5)


def test_vectorized_operations_mirror(oracle):
    """VectorizedOperations / BatchProcessor (reference optimization/high_performance_computing.py
    :98-165, 338-386): batched fields on the tensor-core GEMM, dE = 2 s f for chosen sites equals
    the energy difference of actually flipping them one at a time."""
    k = load_golden("kat_energy_int_n64")
    J, h = torch.from_numpy(k["J"]), torch.from_numpy(k["h"])
    S = torch.from_numpy(k["S"].astype(np.float32))
    B, n = S.shape
    F = sg.VectorizedOperations.vectorized_local_fields(S, J, h)
    assert np.array_equal(F.cpu().numpy().astype(np.float64), k["F"])
    E = sg.BatchProcessor().process_batch_energies(S, J, h)
    assert np.array_equal(E.cpu().numpy().astype(np.float64), k["E"])
    e1 = sg.BatchProcessor().process_batch_energies(S[0], J, h)
    assert e1.dim() == 0 and float(e1) == k["E"][0]
    idx = torch.from_numpy(np.random.default_rng(2).integers(0, n, size=(B, 3)))
    dE = sg.VectorizedOperations.vectorized_energy_differences(S.to(F.device), F, idx).cpu().numpy()
    for j in range(3):
        one = sg.VectorizedOperations.vectorized_spin_flips(S, idx[:, j:j + 1])
        assert int((one != S).sum()) == B
        e_after = np.array([oracle.energy(k["J"], k["h"], one[b].numpy()) for b in range(B)])
        # the diagonal of this instance is zero, so dE(flip i) = E_after - E_before exactly
        assert np.array_equal(dE[:, j], e_after - k["E"])


def test_spin_dynamics_facade(oracle):
    g = load_golden("sa_pm1_n48")
    m = _model(g["J"], g["h"], g["spins0"])
    dyn = sg.SpinDynamics(m, temperature=2.0, random_seed=5)
    e = dyn.sweep()
    assert e == oracle.energy(g["J"], g["h"], m.spins.numpy()) == m.compute_energy()
    assert dyn.total_flips == 48 and 0 < dyn.n_accepted <= 48
    dyn.set_temperature(0.0)
    assert dyn.temperature == 1e-10
    out = dyn.run_dynamics(5)
    assert out["n_sweeps"] == 5 and len(dyn.energy_history) == 6
    assert out["final_energy"] <= e  # T -> 0: only downhill moves


@pytest.mark.parametrize("name", golden_names("ec_"))
def test_energy_computer_matches_reference(name):
    """EnergyComputer on K2 against the known answers recorded from the reference's own
    EnergyComputer (core/energy_computer.py:50-231): exact for integer couplings, 1e-5 for float."""
    g = load_golden(name)
    J, h, S = g["J"], g["h"], g["S"].astype(np.float32)
    exact = bool(np.all(J == np.round(J)) and np.all(h == np.round(h)))

    def close(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert np.array_equal(a, b) if exact else np.allclose(a, b, rtol=1e-5, atol=1e-5), (a, b)

    m = _model(J, h, S[0])
    for mode in sg.ComputeMode:
        ec = sg.EnergyComputer(m, mode)
        close(ec.compute_total_energy(), g["total_" + mode.value])
    ec = sg.EnergyComputer(m, sg.ComputeMode.FULL)
    close(ec.compute_total_energy(torch.from_numpy(S[1])), g["total_other"])
    close([ec.compute_energy_change(i) for i in range(0, J.shape[0], 7)], g["dE"][::7])
    close(ec.compute_energy_changes().numpy(), g["dE"])
    st = ec.compute_energy_stats()
    close([st.total_energy, st.interaction_energy, st.field_energy], g["stats"])
    close(st.per_spin_energy.numpy(), g["per_spin"])
    close(ec.compute_energy_gradient().numpy(), g["gradient"])
    be = ec.compute_batch_energies(torch.from_numpy(S))
    assert be.device.type == "cpu" and be.shape == (S.shape[0],)
    close(be.numpy(), g["batch"])
    # incremental-mode cache semantics (reference :160-167, :298-301)
    inc = sg.EnergyComputer(m, sg.ComputeMode.INCREMENTAL)
    e0 = inc.compute_total_energy()
    inc.update_incremental_cache(3, 2.5)
    assert inc.compute_total_energy() == e0 + 2.5
    inc.invalidate_cache()
    assert inc.compute_total_energy() == e0
    with pytest.raises(IndexError):
        ec.compute_energy_change(J.shape[0])


@pytest.mark.parametrize("integer", [True, False])
def test_mid_run_checkpoint_resumes_bit_for_bit(integer):
    """(spins, best records, counters, ladder state) saved between two launches and
    restored into a FRESH engine continue exactly like the uninterrupted run -- the Philox streams
    are keyed on (seed, absolute sweep, replica), fields are recomputed exactly on restore."""
    import io
    from spin_glass_anneal_rl_b200.engine import Engine
    rng = np.random.default_rng(31)
    n, K, Lad = 300, 8, 5
    R = K * Lad
    if integer:
        a = rng.integers(-2, 3, size=(n, n))
        J = np.triu(a, 1)
        J = (J + J.T).astype(np.float32)
    else:
        a = rng.standard_normal((n, n)) / np.sqrt(n)
        J = ((a + a.T) / 2).astype(np.float32)
        np.fill_diagonal(J, 0.0)
    h = np.zeros(n, np.float32)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    ladder = list(np.geomspace(3.0, 0.3, K))

    def segment(eng, first):
        for r in range(first, first + 3):
            eng.sweep(4, None, seed=5, sweep_base=4 * r, site_order="random", track_best=True)
            eng.refresh_fields()
            eng.exchange(r & 1, seed=6, round=r)

    def state(eng):
        rep_at, temps, att, acc = eng.ladder_state()
        be, bs = eng.best()
        return [t.cpu().numpy() for t in (eng.spins(), eng.energies(), be, bs, eng.accepted(), rep_at, temps, att, acc)]

    def fresh():
        e = Engine(0)
        e.set_model(J, h)
        return e

    a_eng = fresh()
    a_eng.alloc_replicas(R)
    a_eng.set_spins(S0)
    a_eng.init_fields()
    a_eng.set_ladder(ladder)
    segment(a_eng, 0)
    buf = io.BytesIO()
    torch.save(a_eng.checkpoint(), buf)          # through a file image, like a real checkpoint
    segment(a_eng, 3)
    want = state(a_eng)

    b_eng = fresh()
    buf.seek(0)
    b_eng.restore(torch.load(buf), ladder_temps=ladder)
    segment(b_eng, 3)
    got = state(b_eng)
    for w, g_ in zip(want, got):
        assert np.array_equal(w, g_)
    assert want[4].sum() > 0 and want[8].sum() > 0
