"""GPU parity of the small-model sweep kernel (sg_sweep_small.cu: n <= 224, J in shared memory, one
warp per replica, fields in registers) through the C ABI.  Its arithmetic is the sequential
algorithm's (one fp32 FMA per field per accepted flip), so on integer couplings it must be
BIT-EXACT against the oracle in replay mode and identical to the other kernels in Philox mode."""
import numpy as np
import pytest

from conftest import golden_names, has_cuda, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    from spin_glass_anneal_rl_b200.engine import Engine
    return Engine(0)


def _int_instance(rng, n, amp=2, diag=False):
    a = rng.integers(-amp, amp + 1, size=(n, n))
    J = np.triu(a, 1)
    J = (J + J.T).astype(np.float32)
    if diag:
        J[np.arange(n), np.arange(n)] = rng.integers(-amp, amp + 1, size=n)
    h = rng.integers(-amp, amp + 1, size=n).astype(np.float32)
    return J, h


def _setup(engine, J, h, S0):
    engine.set_model(J, h)
    engine.alloc_replicas(S0.shape[0])
    engine.set_spins(S0)
    engine.init_fields()


@pytest.mark.parametrize("n,R,ns,rule,diag", [
    (2, 3, 4, "metropolis", False), (5, 9, 6, "metropolis", True), (33, 8, 5, "glauber", False),
    (100, 33, 4, "metropolis", False), (129, 5, 3, "heat_bath", True), (224, 17, 3, "metropolis", False),
])
def test_small_replay_is_bit_exact_for_integer_couplings(engine, oracle, n, R, ns, rule, diag):
    rng = np.random.default_rng(7 * n + R)
    J, h = _int_instance(rng, n, diag=diag)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
    uni = rng.random((R, ns, n), dtype=np.float32)
    temps = np.linspace(3.0, 0.5, ns)
    _setup(engine, J, h, S0)
    trace = engine.sweep(ns, temps, temps_sweep_stride=1, rule=rule, sites=sites, uniforms=uni,
                         energy_trace=True, kernel="small").cpu().numpy()
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
    assert np.array_equal(engine.energies().cpu().numpy().astype(np.float64), Eo)
    assert np.array_equal(engine.batch_energies(best_s).cpu().numpy(), best_e.cpu().numpy())


@pytest.mark.parametrize("name", [g for g in golden_names("sa_") if "int" in g or "pm1" in g])
def test_small_replays_integer_reference_traces(engine, oracle, name):
    """The golden traces recorded from the reference (integer couplings, n <= 224)."""
    g = load_golden(name)
    c = g["config"]
    J, h = g["J"], g["h"]
    n = J.shape[0]
    if n > 224 or not (np.all(J == np.round(J)) and np.all(h == np.round(h))):
        pytest.skip("small-model kernel: n <= 224, integer couplings for bit-exactness")
    stream = oracle.RawStream(oracle.mt_raw_stream(c["seed"], 2 * n * c["n_sweeps"] + 16))
    ores = oracle.anneal(J, h, g["spins0"], n_sweeps=c["n_sweeps"], T0=c["T0"], Tf=c["Tf"],
                         schedule=c["schedule"], schedule_params=c["params"],
                         record_interval=c["record_interval"], energy_tolerance=c["tol"],
                         rule=c["rule"], stream=stream, trace=True)
    ns = ores.n_sweeps
    _setup(engine, J, h, g["spins0"].reshape(1, n).astype(np.int8))
    uni = np.nan_to_num(ores.extra["uniforms"], nan=0.5).astype(np.float32)
    trace = engine.sweep(ns, ores.extra["temps"], temps_sweep_stride=1, rule=c["rule"],
                         sites=ores.extra["sites"].astype(np.int32), uniforms=uni,
                         energy_trace=True, track_best=True, kernel="small").cpu().numpy()[:, 0]
    best_e, best_s = engine.best()
    assert np.array_equal(engine.spins().cpu().numpy()[0], g["final_spins"])
    assert np.array_equal(best_s.cpu().numpy()[0], g["best_configuration"])
    assert float(best_e[0]) == float(g["best_energy"])
    assert np.array_equal(trace.astype(np.float64), ores.sweep_energies)


@pytest.mark.parametrize("n,R,ns,order", [(100, 70, 6, "random"), (64, 9, 5, "sequential"), (200, 40, 3, "random")])
def test_small_equals_other_kernels_in_philox_mode(engine, n, R, ns, order):
    """Same Philox counters and site order => identical trajectories (integer couplings), whether
    the launch is one piece or cut in two."""
    rng = np.random.default_rng(3 * n + R)
    J, h = _int_instance(rng, n)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    temps = np.linspace(2.5, 0.7, ns)
    outs = []
    for kern, cuts in (("simt", [ns]), ("tc", [ns]), ("small", [ns]), ("small", [2, ns - 2]), ("auto", [ns])):
        _setup(engine, J, h, S0)
        base, traces = 0, []
        for c in cuts:
            traces.append(engine.sweep(c, temps[base:base + c].copy(), temps_sweep_stride=1, seed=77,
                                       sweep_base=5 + base, site_order=order, energy_trace=True,
                                       kernel=kern, coupling_planes=1 if kern == "tc" else 0).cpu().numpy())
            base += c
        outs.append((engine.spins().cpu().numpy(), np.concatenate(traces), engine.accepted().cpu().numpy(),
                     engine.best()[0].cpu().numpy(), engine.best()[1].cpu().numpy(),
                     engine.fields().cpu().numpy()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    assert outs[0][2].sum() > 0


def test_small_float_couplings_are_the_sequential_fma_chain(engine):
    """Float couplings (cfg1: N = 100 Gaussian J): every field is updated by one fp32 FMA per
    accepted flip, so after many sweeps the resident fields stay within 2e-4 of an exact
    recomputation and the carried energies within 1e-5 relative."""
    import torch
    g = torch.Generator().manual_seed(1001)
    n, R = 100, 32
    A = torch.randn(n, n, generator=g)
    J = ((A + A.T) / 2)
    J.fill_diagonal_(0.0)
    h = 0.5 * torch.randn(n, generator=g)
    S0 = (torch.randint(0, 2, (R, n), generator=g) * 2 - 1).to(torch.int8)
    engine.set_model(J, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    engine.sweep(200, np.geomspace(5.0, 0.05, 200), temps_sweep_stride=1, seed=3, kernel="small")
    f = engine.fields().double()
    e = engine.energies().double()
    e2, f2 = engine.batch_energies(engine.spins(), want_fields=True)
    assert (f - f2.double()).abs().max().item() < 2e-4
    assert ((e - e2.double()).abs() / e2.double().abs()).max().item() < 1e-5
    assert engine.accepted().sum().item() > 0


def test_small_site_energy_changes_and_errors(engine):
    import torch
    from spin_glass_anneal_rl_b200._lib import SGError
    rng = np.random.default_rng(4)
    n, R = 50, 6
    J, h = _int_instance(rng, n)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    outs = []
    for kern in ("simt", "small"):
        _setup(engine, J, h, S0)
        de = torch.zeros((R, n), dtype=torch.float32, device="cuda")
        engine.sweep(3, np.array([1.5]), seed=2, site_order="sequential", kernel=kern, site_energy_changes=de)
        outs.append((de.cpu().numpy(), engine.spins().cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.abs(outs[0][0]).sum() > 0
    with pytest.raises(SGError):       # per-block site orders need the big SIMT kernel
        engine.sweep(1, np.array([1.0]), site_order="random_per_block", kernel="small")
    J3, h3 = _int_instance(rng, 300)
    _setup(engine, J3, h3, (rng.integers(0, 2, size=(2, 300)) * 2 - 1).astype(np.int8))
    with pytest.raises(SGError):       # n > 224 does not fit
        engine.sweep(1, np.array([1.0]), kernel="small")
    engine.sweep(1, np.array([1.0]), kernel="auto")
