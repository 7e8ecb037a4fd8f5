"""GPU parity of the block-clique ("groups") sweep kernel, the fast path of BASELINE cfg5
(scheduling QUBO: one clique of agents per task), through the C ABI."""
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


@pytest.fixture(params=["0", "1"], ids=["one-warp-per-word", "partitioned"])
def part(request, monkeypatch):
    """Both group kernels: one warp per 32 replicas with the whole model in shared memory, and
    the P x W grid over partitions of the groups (SG_GRP_PART forces either)."""
    monkeypatch.setenv("SG_GRP_PART", request.param)
    return request.param


def _sched(T, A, integer_h=True, seed=None):
    rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(T, A, seed=seed or T * A))
    if integer_h:
        h = np.round(h).astype(np.float32)
    group_of = (np.arange(T * A) // A).astype(np.int32)
    coupling = np.full(T, val[0], np.float32)
    return rowptr, colidx, val, h, group_of, coupling


@pytest.mark.parametrize("T,A,rule", [(6, 5, "metropolis"), (12, 8, "metropolis"), (9, 7, "glauber"),
                                      (5, 33, "heat_bath")])
def test_groups_replay_is_bit_exact(engine, oracle, part, T, A, rule):
    rowptr, colidx, val, h, group_of, coupling = _sched(T, A)
    n = T * A
    J = inst.csr_to_dense(rowptr, colidx, val, n)
    rng = np.random.default_rng(n)
    R, ns = 41, 3
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
    uni = rng.random((R, ns, n), dtype=np.float32)
    temps = np.array([60.0, 20.0, 5.0])
    engine.set_model_groups(group_of, coupling, h)
    engine.alloc_replicas(R)
    engine.set_spins(S0)
    engine.init_fields()
    _, Eo = oracle.batch_fields_energies(J, h, S0.astype(np.float32))
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
    assert np.array_equal(engine.batch_energies(best_s).cpu().numpy(), best_e.cpu().numpy())


def test_groups_equal_sparse_kernel_in_philox_mode(engine, part):
    rowptr, colidx, val, h, group_of, coupling = _sched(40, 25)
    n, R, ns = 1000, 70, 5
    rng = np.random.default_rng(4)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    outs = []
    for mode in ("csr", "groups"):
        if mode == "csr":
            engine.set_model_csr(rowptr, colidx, val, h)
        else:
            engine.set_model_groups(group_of, coupling, h)
        engine.alloc_replicas(R)
        engine.set_spins(S0)
        engine.init_fields()
        tr = engine.sweep(ns, np.array([25.0]), seed=8, sweep_base=2, site_order="random",
                          energy_trace=True).cpu().numpy()
        outs.append((engine.spins().cpu().numpy(), tr, engine.accepted().cpu().numpy(),
                     engine.best()[0].cpu().numpy(), engine.best()[1].cpu().numpy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert outs[0][2].sum() > 0


def test_partitioned_and_plain_group_kernels_agree(engine, monkeypatch):
    """Uneven group sizes (so partitions are unbalanced), ragged replica count, several launches:
    identical spins, per-sweep energies, acceptance counts and best-so-far records."""
    rng = np.random.default_rng(12)
    sizes = rng.integers(3, 40, size=57)
    group_of = np.repeat(np.arange(57), sizes).astype(np.int32)
    rng.shuffle(group_of)                          # groups are not contiguous in site order
    n = group_of.shape[0]
    coupling = rng.integers(1, 6, size=57).astype(np.float32) * 10.0
    h = rng.integers(-30, 31, size=n).astype(np.float32)
    R = 77
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("SG_GRP_PART", flag)
        engine.set_model_groups(group_of, coupling, h)
        engine.alloc_replicas(R)
        engine.set_spins(S0)
        engine.init_fields()
        traces = []
        for k, T in enumerate((30.0, 12.0, 4.0)):
            traces.append(engine.sweep(3, np.array([T]), seed=21, sweep_base=3 * k, site_order="random",
                                       energy_trace=True).cpu().numpy())
        outs.append((engine.spins().cpu().numpy(), np.concatenate(traces), engine.accepted().cpu().numpy(),
                     engine.energies().cpu().numpy(), engine.best()[0].cpu().numpy(),
                     engine.best()[1].cpu().numpy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert outs[0][2].sum() > 0


def test_groups_full_size_cfg5(engine, part):
    """500 tasks x 100 agents, float fields: no field array exists, so nothing drifts; the carried
    energy stays within 1e-5 relative of an exact recomputation."""
    import torch
    rowptr, colidx, val, h, group_of, coupling = _sched(500, 100, integer_h=False, seed=5005)
    n, R = 50000, 128
    engine.set_model_groups(group_of, coupling, h)
    engine.alloc_replicas(R)
    g = torch.Generator(device="cuda").manual_seed(3)
    engine.set_spins((torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8))
    engine.init_fields()
    e0 = engine.energies().double()
    engine.sweep(5, np.array([40.0]), seed=11, site_order="random")
    e1 = engine.energies().double()
    e_ref = engine.batch_energies(engine.spins()).double()
    assert ((e1 - e_ref).abs() / e_ref.abs()).max().item() < 1e-5
    assert (e_ref < e0).all()
    # against numpy on one replica: E = -1/2 sum_g c (S_g^2 - A) - h.s
    s = engine.spins()[0].cpu().numpy().astype(np.float64)
    S = s.reshape(500, 100).sum(1)
    e_np = -0.5 * 50.0 * float((S * S - 100).sum()) - float(h.astype(np.float64) @ s)
    assert abs(e_ref[0].item() - e_np) <= 1e-6 * abs(e_np)
    best_e, best_s = engine.best()
    eb = engine.batch_energies(best_s).double()
    assert ((eb - best_e.double()).abs() <= 1e-5 * eb.abs()).all()


def test_host_api_detects_clique_structure():
    import torch
    import spin_glass_anneal_rl_b200 as sg
    rowptr, colidx, val, h, _, _ = _sched(500, 100, integer_h=False, seed=5005)
    n = h.shape[0]
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    model = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=True))
    model.couplings = torch.sparse_coo_tensor(np.stack([rows, colidx]), torch.from_numpy(val), (n, n))
    model.external_fields = torch.from_numpy(h)
    cfg = sg.GPUAnnealerConfig(n_sweeps=20, initial_temp=200.0, final_temp=5.0, random_seed=3,
                               schedule_params={"alpha": 0.8}, n_replicas=64, record_interval=5)
    res = sg.GPUAnnealer(cfg).anneal(model)
    assert model._sg_engine[1].kind == "groups"
    s = res.best_configuration.numpy().astype(np.float64)
    S = s.reshape(500, 100).sum(1)
    e_np = -0.5 * 50.0 * float((S * S - 100).sum()) - float(h.astype(np.float64) @ s)
    assert abs(res.best_energy - e_np) <= 1e-5 * abs(e_np)
    assert res.energy_history[-1] < res.energy_history[0]
