"""UpdateRule.WOLFF on the GPU (csrc/sg_wolff.cu, C ABI sg_sweep_wolff) against the oracle's
restatement of the reference's dense cluster move (core/spin_dynamics.py:193-262), which is pinned
to tests/golden/wolff_*.npz (recorded from the unmodified reference by make_wolff_golden.py).

Injected mode: the kernel gets the start sites and the flat list of uniforms the reference drew
and must walk the same clusters -- same spins after every sweep, same number of uniforms consumed,
same cluster sizes; energies bit-exact on integer couplings, 1e-5 relative on float couplings.
Philox mode: distribution of the energy after a few sweeps against independent oracle chains.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, has_cuda, load_golden

pytestmark = pytest.mark.gpu

import spin_glass_anneal_rl_b200 as sg
from spin_glass_anneal_rl_b200._lib import SGError
from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType
from spin_glass_anneal_rl_b200.core.spin_dynamics import SpinDynamics, UpdateRule
from spin_glass_anneal_rl_b200.engine import Engine

REL = 1e-5


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_cuda():
        pytest.skip("needs a CUDA device")


def _is_integer(J, h):
    return bool(np.all(J == np.round(J)) and np.all(h == np.round(h)))


def _close(a, b, exact, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, what
    if exact:
        assert np.array_equal(a, b), what
    else:
        assert np.allclose(a, b, rtol=REL, atol=REL), what


def _model(J, h, spins=None):
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=J.shape[0], use_sparse=False))
    m.set_couplings_from_matrix(torch.from_numpy(np.asarray(J, np.float32)))
    m.set_external_fields(torch.from_numpy(np.asarray(h, np.float32)))
    if spins is not None:
        m.set_spins(torch.from_numpy(np.asarray(spins, np.float32)))
    return m


def _oracle_run(oracle, g):
    c = g["config"]
    stream = oracle.RawStream(oracle.mt_raw_stream(c["seed"], int(g["raw_consumed"]) + 4096))
    return oracle.anneal(g["J"], g["h"], g["spins0"], n_sweeps=c["n_sweeps"], T0=c["T0"], Tf=c["Tf"],
                         schedule=c["schedule"], schedule_params=c["params"],
                         record_interval=c["record_interval"], energy_tolerance=c["tol"],
                         rule="wolff", stream=stream, trace=True)


def _sparse_neg_model(n, seed, degree=3, integer=True, asym=False):
    """Random sparse couplings, about `degree` negative and as many positive ones per row."""
    rs = np.random.RandomState(seed)
    J = np.zeros((n, n), np.float32)
    m = n * degree
    i, j = rs.randint(0, n, m), rs.randint(0, n, m)
    v = -np.ones(m, np.float32) if integer else -np.abs(rs.standard_normal(m)).astype(np.float32)
    J[i, j] = v
    i, j = rs.randint(0, n, m), rs.randint(0, n, m)
    J[i, j] = rs.randint(1, 3, m) if integer else np.abs(rs.standard_normal(m))
    if not asym:
        J = np.triu(J, 1)
        J = J + J.T
    np.fill_diagonal(J, 0.0)
    h = rs.randint(-1, 2, n).astype(np.float32) if integer else (0.2 * rs.standard_normal(n)).astype(np.float32)
    return J.astype(np.float32), h


# ------------------------------------------------------------------ replay of the reference's runs
def _set_form(monkeypatch, form):
    """list: neighbour lists (the default where every row has at most 32 negative couplings);
    row: the row-walking forms (a warp per replica up to n = 1792, a CTA above); cta: a CTA always."""
    if form == "list":
        monkeypatch.delenv("SG_WOLFF_FORM", raising=False)
    else:
        monkeypatch.setenv("SG_WOLFF_FORM", form)


@pytest.mark.parametrize("form", ["list", "row"])
@pytest.mark.parametrize("name", golden_names("wolff_"))
def test_engine_replay_walks_the_reference_clusters(oracle, monkeypatch, name, form):
    _set_form(monkeypatch, form)
    g = load_golden(name)
    ores = _oracle_run(oracle, g)
    n = g["J"].shape[0]
    ns = ores.n_sweeps
    exact = _is_integer(g["J"], g["h"])
    eng = Engine(0)
    eng.set_model(g["J"], g["h"])
    eng.alloc_replicas(1)
    eng.set_spins(g["spins0"].astype(np.int8).reshape(1, n))
    eng.init_fields()
    uni = ores.extra["uniforms"]
    off = ores.extra["uniform_offsets"]
    cursor = torch.zeros(1, dtype=torch.int64, device=eng.device)
    # cut into launches of 1, 2, 3, ... sweeps: the cursor carries the stream position across them
    done, k, energies = 0, 1, []
    spins_ref = g["spins0"].astype(np.float32).copy()
    while done < ns:
        k = min(k, ns - done)
        tr = eng.sweep_wolff(k, ores.extra["temps"][done:done + k], temps_sweep_stride=1,
                             sites=ores.extra["sites"][done:done + k], uniforms=uni, cursor=cursor,
                             energy_trace=True, track_best=True, sweep_base=done)
        energies += tr[:, 0].cpu().tolist()
        done += k
        assert int(cursor.item()) == int(off[done]), "uniforms consumed"
        oracle.wolff_sweeps_scheduled(g["J"], g["h"], spins_ref, ores.extra["temps"][done - k:done],
                                      ores.extra["sites"][done - k:done], uni[off[done - k]:off[done]])
        assert np.array_equal(eng.spins()[0].cpu().numpy(), spins_ref.astype(np.int8)), f"spins after sweep {done}"
        k += 1
    _close(energies, ores.sweep_energies, exact, "energy after every sweep")
    assert np.array_equal(eng.spins()[0].cpu().numpy(), g["final_spins"])
    be, bs = eng.best()
    _close(float(be[0].item()), float(g["best_energy"]), exact, "best energy")
    assert np.array_equal(bs[0].cpu().numpy(), g["best_configuration"])
    _close(eng.fields()[0].cpu().numpy(),
           oracle.batch_fields_energies(g["J"].T.copy(), g["h"], g["final_spins"].astype(np.float32)[None])[0][0],
           exact, "local fields after the run")


@pytest.mark.parametrize("name", golden_names("wolff_"))
def test_anneal_replay_reproduces_the_reference(oracle, name):
    """GPUAnnealer.anneal(model, UpdateRule.WOLFF) in replay mode == the reference's own result."""
    g = load_golden(name)
    c = g["config"]
    ores = _oracle_run(oracle, g)
    m = _model(g["J"], g["h"], g["spins0"])
    cfg = sg.GPUAnnealerConfig(
        n_sweeps=c["n_sweeps"], initial_temp=c["T0"], final_temp=c["Tf"],
        schedule_type=ScheduleType(c["schedule"]), schedule_params=dict(c["params"]),
        record_interval=c["record_interval"], energy_tolerance=c["tol"], random_seed=c["seed"],
        rng_mode="replay", replay={"sites": ores.extra["sites"], "uniforms": ores.extra["uniforms"]})
    res = sg.GPUAnnealer(cfg).anneal(m, UpdateRule.WOLFF)
    exact = _is_integer(g["J"], g["h"])
    assert res.n_sweeps == int(g["n_sweeps_done"])
    assert np.array_equal(res.best_configuration.numpy().astype(np.int8), g["best_configuration"])
    assert np.array_equal(m.spins.cpu().numpy().astype(np.int8), g["final_spins"])
    _close(res.best_energy, g["best_energy"], exact, "best energy")
    _close(res.energy_history, g["energy_history"], exact, "energy history")
    _close(res.acceptance_rate_history, g["acceptance_rate_history"], True, "acceptance rates")
    assert np.allclose(res.temperature_history, g["temperature_history"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", golden_names("ptw_"))
def test_parallel_tempering_replay_with_the_cluster_move(oracle, name):
    """ParallelTempering.run(model, UpdateRule.WOLFF) in replay mode == the reference's run whose
    replicas all use _wolff_update (n_threads=1, device='cpu'): configurations per temperature slot,
    exchange statistics, per-slot energy histories, best configuration, acceptance rates."""
    g = load_golden(name)
    c = g["config"]
    n, K = g["J"].shape[0], c["n_replicas"]
    stream = oracle.RawStream(oracle.mt_raw_stream(c["seed"], int(g["raw_consumed"]) + 4096))
    ores = oracle.parallel_tempering(
        g["J"], g["h"], n_replicas=K, n_sweeps=c["n_sweeps"], temp_min=c["tmin"], temp_max=c["tmax"],
        temp_distribution=c["dist"], exchange_interval=c["exchange_interval"],
        record_interval=c["record_interval"], rule="wolff", stream=stream,
        np_rng=np.random.RandomState(c["seed"]), trace=True)
    assert ores.raw_consumed == int(g["raw_consumed"])
    cfg = sg.ParallelTemperingConfig(
        n_replicas=K, n_sweeps=c["n_sweeps"], temp_min=c["tmin"], temp_max=c["tmax"],
        temp_distribution=c["dist"], exchange_interval=c["exchange_interval"],
        record_interval=c["record_interval"], random_seed=c["seed"], rng_mode="replay",
        replay={k: ores.extra[k] for k in ("spins0", "sites", "uniforms", "exchange_draws")})
    pt = sg.ParallelTempering(cfg)
    res = pt.run(_model(g["J"], g["h"]), UpdateRule.WOLFF)
    exact = _is_integer(g["J"], g["h"])
    assert np.array_equal(pt.exchange_attempts, g["exchange_attempts"])
    assert np.array_equal(pt.exchange_accepts, g["exchange_accepts"])
    slot_spins = pt._final_spins.cpu().numpy()[pt._rung_replica[:K]]
    assert np.array_equal(slot_spins, g["final_spins"]), "configurations per temperature slot differ"
    assert np.array_equal(res.best_configuration.numpy().astype(np.int8), g["best_configuration"])
    _close(res.best_energy, g["best_energy"], exact, "best energy")
    _close(np.array(pt.energy_histories), g["energy_histories"], exact, "per-slot energy histories")
    assert np.allclose(res.acceptance_rate_history, g["acceptance_rate_history"], rtol=0, atol=1e-12)
    assert g["exchange_accepts"].sum() > 0


# ------------------------------------------------------------------ wider shapes against the oracle
@pytest.mark.parametrize("n,R,integer,asym,T,form", [
    (300, 5, True, False, 5.0, "list"),    # neighbour lists (rows with at most 32 negative couplings)
    (300, 5, True, False, 5.0, "row"),     # n_pad = 896: a warp per replica walks the row
    (300, 11, True, False, 5.0, "cta"),    # the same through the CTA-per-replica form (one pass)
    (1100, 3, True, True, 3.5, "list"),    # asymmetric couplings (row of the dequeued site)
    (1100, 3, True, True, 3.5, "row"),     # n_pad = 1792, 10 float4 loads per lane
    (1100, 3, True, True, 3.5, "cta"),     # two passes of 1024 columns
    (1000, 4, False, False, 3.0, "list"),  # float couplings
    (1000, 4, False, False, 3.0, "row"),
    (4100, 2, True, False, 5.0, "list"),   # a model beyond the warp-per-replica row walk
    (4100, 2, True, False, 5.0, "row"),    # five passes (n_pad = 4480: the CTA form)
])
def test_replicas_with_their_own_streams(oracle, monkeypatch, n, R, integer, asym, T, form):
    _set_form(monkeypatch, form)
    J, h = _sparse_neg_model(n, 100 + n, integer=integer, asym=asym)
    rs = np.random.RandomState(n)
    ns = 2
    temps = np.array([T, 0.8 * T])
    spins0 = (rs.randint(0, 2, (R, n)) * 2 - 1).astype(np.float32)
    sites = np.zeros((R, ns, n), np.int32)
    streams, want_spins, want_e, want_flips = [], [], [], []
    for r in range(R):
        s = spins0[r].copy()
        st = oracle.RawStream(oracle.mt_raw_stream(7000 + r, 64 * n * ns + 200000))
        es, flips, tr = oracle.wolff_sweeps(J, h, s, temps, st, trace=True)
        sites[r] = tr["sites"]
        streams.append(tr["uniforms"])
        want_spins.append(s.astype(np.int8))
        want_e.append(es)
        want_flips.append(int(flips.sum()))
    m = max(len(u) for u in streams) + 1
    assert min(len(u) for u in streams) > n // 4, "the case must exercise the cluster growth"
    uni = np.full((R, m), 2.0, np.float32)   # 2.0 beyond the end: never accepted, and never reached
    for r, u in enumerate(streams):
        uni[r, :len(u)] = u
    eng = Engine(0)
    eng.set_model(J, h)
    eng.alloc_replicas(R)
    eng.set_spins(spins0.astype(np.int8))
    eng.init_fields()
    tr = eng.sweep_wolff(ns, temps, temps_sweep_stride=1, sites=sites, sites_replica_stride=ns * n,
                         sites_sweep_stride=n, uniforms=uni, energy_trace=True)
    assert np.array_equal(eng.wolff_cursor.cpu().numpy(), np.array([len(u) for u in streams]))
    assert np.array_equal(eng.spins().cpu().numpy(), np.stack(want_spins))
    assert np.array_equal(eng.accepted().cpu().numpy().astype(np.int64), np.array(want_flips))
    _close(tr.cpu().numpy().T, np.stack(want_e), integer, "energies per sweep")


def test_stream_too_short_is_reported(oracle):
    g = load_golden("wolff_dense_int_n20")
    ores = _oracle_run(oracle, g)
    n = g["J"].shape[0]
    eng = Engine(0)
    eng.set_model(g["J"], g["h"])
    eng.alloc_replicas(1)
    eng.set_spins(g["spins0"].astype(np.int8).reshape(1, n))
    eng.init_fields()
    with pytest.raises(SGError, match="more uniforms"):
        eng.sweep_wolff(2, ores.extra["temps"][:2], temps_sweep_stride=1, sites=ores.extra["sites"][:2],
                        uniforms=ores.extra["uniforms"][:5])
    with pytest.raises(SGError, match="own entry point"):
        from spin_glass_anneal_rl_b200 import _lib
        import ctypes
        p = _lib.SweepParams()
        p.struct_size = ctypes.sizeof(_lib.SweepParams)
        p.n_sweeps, p.rule = 1, _lib.SG_RULE["wolff"]
        _lib.check(eng._lib.sg_sweep(eng._h, ctypes.byref(p), eng.stream), "sg_sweep")


# ------------------------------------------------------------------ Philox mode
def test_philox_mode_matches_the_oracle_distribution(oracle):
    """Energy and cluster volume after 4 sweeps at fixed T: 512 GPU replicas (in-kernel Philox)
    against 192 oracle chains on independent mt19937 streams, same start configuration."""
    n, T, ns = 64, 4.0, 4
    J, h = _sparse_neg_model(n, 5, degree=3, integer=True)
    s0 = (np.random.RandomState(1).randint(0, 2, n) * 2 - 1).astype(np.float32)
    oe, of = [], []
    for c in range(192):
        s = s0.copy()
        es, flips, _ = oracle.wolff_sweeps(J, h, s, [T] * ns, oracle.RawStream(oracle.mt_raw_stream(900 + c, 400000)))
        oe.append(es[-1])
        of.append(flips.sum())
    R = 512
    eng = Engine(0)
    eng.set_model(J, h)
    eng.alloc_replicas(R)
    eng.set_spins(np.tile(s0.astype(np.int8), (R, 1)))
    eng.init_fields()
    tr = eng.sweep_wolff(ns, np.array([T]), site_order="random", seed=1234, energy_trace=True)
    ge = tr[-1].cpu().numpy().astype(np.float64)
    gf = eng.accepted().cpu().numpy().astype(np.float64)
    for a, b, what in ((ge, np.array(oe), "energy"), (gf, np.array(of, np.float64), "cluster volume")):
        z = (a.mean() - b.mean()) / np.sqrt(a.var() / len(a) + b.var() / len(b))
        assert abs(z) < 4.5, f"{what}: GPU {a.mean():.3f} vs oracle {b.mean():.3f} (z = {z:.2f})"
    # exact energies, reproducible, and independent of how the replicas are cut over engines
    assert np.array_equal(ge, oracle.batch_fields_energies(J, h, eng.spins().cpu().numpy().astype(np.float32))[1])
    first = eng.spins().clone()
    eng.set_spins(np.tile(s0.astype(np.int8), (R, 1)))
    eng.init_fields()
    eng.sweep_wolff(2, np.array([T]), site_order="random", seed=1234)
    eng.sweep_wolff(2, np.array([T]), site_order="random", seed=1234, sweep_base=2)
    assert torch.equal(eng.spins(), first)
    half = Engine(0)
    half.set_model(J, h)
    half.alloc_replicas(R // 2)
    half.set_spins(np.tile(s0.astype(np.int8), (R // 2, 1)))
    half.init_fields()
    half.sweep_wolff(ns, np.array([T]), site_order="random", seed=1234, replica_base=R // 2)
    assert torch.equal(half.spins(), first[R // 2:])
    assert len({tuple(r) for r in first.cpu().numpy()[:64].tolist()}) > 32   # replicas differ


@pytest.mark.parametrize("n,dense", [(200, False), (1500, False), (200, True)])
def test_kernel_forms_agree_in_philox_mode(monkeypatch, n, dense):
    """Neighbour lists, a warp per replica and a CTA per replica draw the same Philox numbers per
    (update, visit, column quad): same clusters, same spins.  A dense model (about n / 2 negative
    couplings per row) has no neighbour lists: the default is then the row-walking form."""
    if dense:
        rs = np.random.RandomState(n)
        G = rs.normal(0.0, 1.0, size=(n, n)).astype(np.float32)
        J = ((G + G.T) / 2).astype(np.float32)
        np.fill_diagonal(J, 0.0)
        h = np.zeros(n, np.float32)
        temps = np.array([100.0, 80.0])   # (clusters of a few sites)
    else:
        J, h = _sparse_neg_model(n, 40 + n, integer=False)
        temps = np.array([3.0, 2.5])
    R = 19
    S0 = (np.random.RandomState(n).randint(0, 2, (R, n)) * 2 - 1).astype(np.int8)
    out = []
    for form in ("list", "row", "cta"):
        _set_form(monkeypatch, form)
        eng = Engine(0)
        eng.set_model(J, h)
        eng.alloc_replicas(R)
        eng.set_spins(S0)
        eng.init_fields()
        tr = eng.sweep_wolff(2, temps, temps_sweep_stride=1, seed=77, energy_trace=True)
        out.append((eng.spins().cpu().numpy(), tr.cpu().numpy(), eng.accepted().cpu().numpy()))
    assert out[0][2].sum() > 2.2 * R * n, "the case must exercise the cluster growth"
    for o in out[1:]:
        for a, b in zip(out[0], o):
            assert np.array_equal(a, b)


def test_api_paths_take_wolff(oracle):
    J, h = _sparse_neg_model(48, 9, degree=3, integer=True)
    s0 = (np.random.RandomState(2).randint(0, 2, 48) * 2 - 1).astype(np.float32)
    m = _model(J, h, s0)
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=12, initial_temp=4.0, final_temp=0.5, random_seed=4,
                                              record_interval=3, n_replicas=8)).anneal(m, UpdateRule.WOLFF)
    assert res.best_energy == oracle.energy(J, h, res.best_configuration.numpy())
    assert res.acceptance_rate_history[0] == 0.0 and set(res.acceptance_rate_history[1:]) == {1.0}
    assert res.best_energy <= min(res.energy_history)
    # adaptive schedule: the reference feeds it the acceptance rate 1.0 of cluster moves
    res = sg.GPUAnnealer(sg.GPUAnnealerConfig(
        n_sweeps=8, initial_temp=4.0, final_temp=0.5, random_seed=4, record_interval=1,
        schedule_type=ScheduleType.ADAPTIVE,
        schedule_params={"alpha": 0.9, "adaptation_window": 3, "target_acceptance": 0.3})).anneal(m, UpdateRule.WOLFF)
    want = oracle.AdaptiveState(4.0, 0.5, alpha=0.9, target_acceptance=0.3, adaptation_window=3)
    temps = [want.update(s, 0.0 if s == 0 else 1.0) for s in range(8)]
    assert np.allclose(res.temperature_history[1:], temps, rtol=1e-12)
    # SpinDynamics facade
    m2 = _model(J, h, s0)
    dyn = SpinDynamics(m2, temperature=2.0, update_rule=UpdateRule.WOLFF, random_seed=3)
    e = dyn.sweep()
    assert e == pytest.approx(oracle.energy(J, h, m2.spins.cpu().numpy()), abs=1e-4)
    assert dyn.n_rejected == 0 and dyn.n_accepted >= 48 and dyn.get_acceptance_rate() == 1.0
    # parallel tempering sweeps with the cluster move
    pt = sg.ParallelTempering(sg.ParallelTemperingConfig(n_replicas=4, n_sweeps=10, temp_min=0.5, temp_max=4.0,
                                                         exchange_interval=2, record_interval=2, random_seed=5))
    r = pt.run(_model(J, h, s0), UpdateRule.WOLFF)
    assert r.best_energy == oracle.energy(J, h, r.best_configuration.numpy())
    # models that are not dense say so
    big = sg.IsingModel(sg.IsingModelConfig(n_spins=9000, use_sparse=True))
    idx = torch.tensor([[0, 1], [1, 0]])
    big.couplings = torch.sparse_coo_tensor(idx, torch.tensor([-1.0, -1.0]), (9000, 9000)).coalesce()
    with pytest.raises(NotImplementedError):
        sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=1)).anneal(big, UpdateRule.WOLFF)
