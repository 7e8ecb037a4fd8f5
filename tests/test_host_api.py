"""CPU-side checks: host logic of the reference-shaped API, and the C-ABI library's symbols.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_names, load_golden

import spin_glass_anneal_rl_b200 as sg
from spin_glass_anneal_rl_b200 import _lib
from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import (AdaptiveSchedule,
                                                                       ScheduleType,
                                                                       TemperatureScheduler)


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sg_b200.h")).read()
    declared = set(re.findall(r"\b(sg_[a-z_0-9]+)\s*\(", header))
    declared -= {"sg_engine"}
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.build())
    for name in sorted(declared):
        assert hasattr(lib, name), f"libsg_b200.so lacks {name}"
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))
    assert _lib.load().sg_abi_version() == _lib.SG_ABI_VERSION == 5


def test_param_structs_match_header_layout(tmp_path):
    """ctypes mirrors of the parameter structs have the size the C compiler gives them."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sg_b200.h"\nint main(void){printf("%zu %zu\\n",'
                   'sizeof(sg_sweep_params), sizeof(sg_exchange_params)); printf("%zu\\n", '
                   'sizeof(sg_wolff_params)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert ctypes.sizeof(_lib.SweepParams) == a
    assert ctypes.sizeof(_lib.ExchangeParams) == b
    assert ctypes.sizeof(_lib.WolffParams) == c


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spin_glass_anneal_rl_b200.engine import Engine
    with pytest.raises(Exception):
        Engine(0)
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=8, use_sparse=False))
    with pytest.raises(Exception):
        sg.GPUAnnealer(sg.GPUAnnealerConfig(n_sweeps=2)).anneal(m)


def test_wolff_rule_is_routed_and_guarded():
    """UpdateRule.WOLFF is a known rule of the sweep path; it needs a model the engine holds dense."""
    from spin_glass_anneal_rl_b200.annealing import _backend
    from spin_glass_anneal_rl_b200.core.spin_dynamics import UpdateRule
    assert _backend.rule_name(UpdateRule.WOLFF) == "wolff" and _lib.SG_RULE["wolff"] == 3
    with pytest.raises(NotImplementedError):
        _backend.rule_name("swendsen_wang")

    class _Eng:
        kind = "csr"
    with pytest.raises(NotImplementedError, match="dense"):
        _backend.require_dense_for_wolff(_Eng())
    _Eng.kind = "dense"
    _backend.require_dense_for_wolff(_Eng())
    header = open(os.path.join(ROOT, "include", "sg_b200.h")).read()
    assert "#define SG_RULE_WOLFF 3" in header and "sg_sweep_wolff" in header


def test_spin_dynamics_history_analysis():
    """SpinDynamics.get_autocorrelation_time / thermal_equilibrium_check (reference
    core/spin_dynamics.py:361-421) on histories with known answers; no device involved."""
    from spin_glass_anneal_rl_b200.core.spin_dynamics import SpinDynamics, UpdateRule
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=8, use_sparse=False))
    dyn = SpinDynamics(m, temperature=1.0, update_rule=UpdateRule.METROPOLIS, random_seed=1)
    assert dyn.get_autocorrelation_time() == float("inf")          # fewer than 10 records
    with pytest.raises(ValueError):
        dyn.get_autocorrelation_time("susceptibility")
    rs = np.random.RandomState(0)
    # AR(1) with coefficient rho: the autocorrelation rho^k first falls below 1/e at ceil(1/-ln rho)
    rho, x = 0.9, [0.0]
    for _ in range(20000):
        x.append(rho * x[-1] + rs.standard_normal())
    dyn.energy_history = x
    tau = dyn.get_autocorrelation_time("energy")
    assert abs(tau - np.ceil(-1.0 / np.log(rho))) <= 2
    dyn.magnetization_history = list(rs.standard_normal(500))       # white noise: below 1/e at lag 1
    assert dyn.get_autocorrelation_time("magnetization") == 1.0
    # equilibrium check: same distribution in both windows -> True; a drift -> False
    assert dyn.thermal_equilibrium_check(window_size=50000) is False     # not enough records
    dyn.energy_history = list(rs.standard_normal(400))
    assert dyn.thermal_equilibrium_check(window_size=200) is True
    dyn.energy_history = list(rs.standard_normal(200)) + list(5.0 + rs.standard_normal(200))
    assert dyn.thermal_equilibrium_check(window_size=200) is False


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spin_glass_anneal_rl_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle", "").replace("The oracle", "") \
                    or f == "__never__", f"{f} mentions the oracle"


# ------------------------------------------------------------------ schedules
@pytest.mark.parametrize("kind,params", [
    ("linear", {}), ("exponential", {}), ("geometric", {"alpha": 0.93}),
    ("logarithmic", {"c": 2.0}), ("power_law", {"k": 0.7}), ("fast", {}), ("boltzmann", {})])
def test_schedules_equal_oracle_restatement(oracle, kind, params):
    sched = TemperatureScheduler.create_schedule(ScheduleType(kind), 4.0, 0.2, 60, **params)
    want = [oracle.schedule_temperature(kind, s, 4.0, 0.2, 60, **params) for s in range(80)]
    got = [sched.get_temperature(s) for s in range(80)]
    assert np.array_equal(np.array(got, np.float64), np.array(want, np.float64))
    assert np.array_equal(sched.precompute(60), np.array(want[:60], np.float64))


def test_adaptive_schedule_equals_oracle(oracle):
    rng = np.random.default_rng(0)
    rates = rng.random(50)
    sched = TemperatureScheduler.create_schedule(ScheduleType.ADAPTIVE, 4.0, 0.2, 50, alpha=0.9,
                                                 adaptation_window=5, target_acceptance=0.3)
    ref = oracle.AdaptiveState(4.0, 0.2, alpha=0.9, adaptation_window=5, target_acceptance=0.3)
    assert isinstance(sched, AdaptiveSchedule) and not sched.stateless
    for s, r in enumerate(rates):
        assert sched.update(s, acceptance_rate=float(r)) == ref.update(s, float(r))


@pytest.mark.parametrize("name", [n for n in golden_names("sa_sched_")])
def test_schedule_reproduces_reference_temperature_history(name):
    g = load_golden(name)
    c = g["config"]
    if c["schedule"] == "adaptive":
        pytest.skip("needs run-time acceptance feedback (covered on the GPU)")
    sched = TemperatureScheduler.create_schedule(ScheduleType(c["schedule"]), c["T0"], c["Tf"],
                                                 c["n_sweeps"], **c["params"])
    t = sched.precompute(c["n_sweeps"])
    want = g["temperature_history"][1:]
    got = t[::c["record_interval"]][:len(want)]
    assert np.allclose(got, want, rtol=1e-15, atol=0)


# ------------------------------------------------------------------ result container
@pytest.mark.parametrize("name", ["sa_pm1_n48", "sa_cfg1_float_n100", "sa_converge_n12"])
def test_result_postprocessing_matches_reference(name):
    g = load_golden(name)
    res = sg.AnnealingResult(best_configuration=torch.from_numpy(g["best_configuration"]).float(),
                             best_energy=float(g["best_energy"]),
                             energy_history=g["energy_history"].tolist(),
                             temperature_history=g["temperature_history"].tolist(),
                             acceptance_rate_history=g["acceptance_rate_history"].tolist(),
                             total_time=0.1, n_sweeps=int(g["n_sweeps_done"]))
    assert np.isclose(res.energy_std, float(g["energy_std"]), rtol=1e-12)
    assert (-1 if res.convergence_sweep is None else res.convergence_sweep) == int(g["convergence_sweep"])
    assert res.final_temperature == g["temperature_history"][-1]
    assert res.final_acceptance_rate == g["acceptance_rate_history"][-1]


def test_result_validation_and_roundtrip(tmp_path):
    ok = dict(best_configuration=torch.ones(4), best_energy=-1.0, energy_history=[0.0, -1.0],
              temperature_history=[1.0, 0.5], acceptance_rate_history=[0.0, 0.4], total_time=0.5,
              n_sweeps=2)
    res = sg.AnnealingResult(**ok)
    with pytest.raises(TypeError):
        sg.AnnealingResult(**{**ok, "best_configuration": [1, 1]})
    with pytest.raises(ValueError):
        sg.AnnealingResult(**{**ok, "best_energy": float("nan")})
    with pytest.raises(ValueError):
        sg.AnnealingResult(**{**ok, "n_sweeps": 0})
    with pytest.raises(ValueError):
        sg.AnnealingResult(**{**ok, "total_time": -1.0})
    path = str(tmp_path / "r.npz")
    res.save(path)
    back = sg.AnnealingResult.load(path)
    assert back.best_energy == res.best_energy and back.energy_history == res.energy_history
    assert torch.equal(back.best_configuration, res.best_configuration)
    assert back.convergence_sweep is None and back.random_seed is None


# ------------------------------------------------------------------ model container
def test_config_defaults_match_reference():
    c = sg.GPUAnnealerConfig()
    assert (c.n_sweeps, c.initial_temp, c.final_temp, c.record_interval) == (1000, 10.0, 0.01, 10)
    assert c.schedule_type is ScheduleType.GEOMETRIC and c.schedule_params == {"alpha": 0.95}
    assert (c.block_size, c.shared_memory_size, c.energy_tolerance, c.random_seed) == (256, 49152, 1e-8, None)
    p = sg.ParallelTemperingConfig()
    assert (p.n_replicas, p.n_sweeps, p.temp_min, p.temp_max) == (8, 1000, 0.1, 10.0)
    assert (p.temp_distribution, p.exchange_interval, p.exchange_method) == ("geometric", 10, "nearest_neighbor")
    m = sg.IsingModelConfig(n_spins=3)
    assert (m.coupling_strength, m.external_field_strength, m.use_sparse, m.device) == (1.0, 0.5, True, "cpu")


@pytest.mark.parametrize("name", golden_names("kat_"))
def test_model_energy_and_local_field_kats(name):
    """IsingModel (dense and sparse) reproduces energies / local fields / flip dE recorded
    from the reference model."""
    g = load_golden(name)
    J, h, S = torch.from_numpy(g["J"]), torch.from_numpy(g["h"]), g["S"].astype(np.float32)
    for sparse in (False, True):
        m = sg.IsingModel(sg.IsingModelConfig(n_spins=J.shape[0], use_sparse=sparse))
        m.set_couplings_from_matrix(J)
        m.set_external_fields(h)
        for b in range(S.shape[0]):
            m.set_spins(torch.from_numpy(S[b]))
            assert np.isclose(m.compute_energy(), g["E"][b], rtol=1e-5, atol=1e-5)
            for i in range(0, J.shape[0], max(1, J.shape[0] // 8)):
                assert np.isclose(m.get_local_field(i), g["F"][b, i], rtol=1e-5, atol=1e-5)
                e0 = m.compute_energy()
                d = m.flip_spin(i)
                assert np.isclose(d, g["dE"][b, i], rtol=1e-5, atol=1e-5)
                assert np.isclose(m.compute_energy() - e0, d, rtol=1e-4, atol=1e-4)
                m.flip_spin(i)


def test_model_api_surface():
    torch.manual_seed(3)
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=6, use_sparse=True))
    assert m.couplings.is_sparse and set(m.spins.tolist()) <= {-1.0, 1.0}
    m.set_coupling(0, 5, -2.0)
    assert m.get_coupling(5, 0) == -2.0
    with pytest.raises(ValueError):
        m.set_coupling(0, 6, 1.0)
    m.spins = torch.ones(6)
    m._invalidate_cache()
    assert m.get_magnetization() == 1.0 and m.compute_energy() == 2.0
    c = m.copy()
    c.set_coupling(1, 2, 1.0)
    assert m.get_coupling(1, 2) == 0.0
    back = sg.IsingModel.from_dict(m.to_dict())
    assert torch.equal(back.dense_couplings(), m.dense_couplings()) and torch.equal(back.spins, m.spins)
    with pytest.raises(ValueError):
        sg.IsingModel(sg.IsingModelConfig(n_spins=0))


def test_pt_ladder_matches_reference():
    for name in golden_names("pt_"):
        g = load_golden(name)
        c = g["config"]
        pt = sg.ParallelTempering(sg.ParallelTemperingConfig(
            n_replicas=c["n_replicas"], temp_min=c["tmin"], temp_max=c["tmax"],
            temp_distribution=c["dist"]))
        assert np.allclose(pt.temperatures, g["temperatures"], rtol=1e-15, atol=0)
        assert pt.anneal.__func__ is pt.run.__func__
    with pytest.raises(ValueError):
        sg.ParallelTempering(sg.ParallelTemperingConfig(temp_distribution="nope"))


def test_install_as_spin_glass_rl():
    import sys
    for k in [k for k in sys.modules if k.startswith("spin_glass_rl")]:
        del sys.modules[k]
    sg.install_as_spin_glass_rl()
    from spin_glass_rl.core.ising_model import IsingModel as A
    from spin_glass_rl.annealing.gpu_annealer import GPUAnnealer as B
    from spin_glass_rl.annealing.parallel_tempering import ParallelTempering as C
    from spin_glass_rl.core.spin_dynamics import UpdateRule as D
    assert A is sg.IsingModel and B is sg.GPUAnnealer and C is sg.ParallelTempering
    assert D.METROPOLIS.value == "metropolis" and D.WOLFF.value == "wolff"
    from spin_glass_rl.annealing.multi_gpu import MultiGPUAnnealer, MultiGPUConfig
    assert hasattr(MultiGPUAnnealer, "anneal_replica_exchange") and MultiGPUConfig().strategy == "data_parallel"
    for k in [k for k in sys.modules if k.startswith("spin_glass_rl")]:
        del sys.modules[k]


# ------------------------------------------------------------------ structure detection (host logic)
def _coo(n, rowptr, colidx, val):
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    return rows, colidx.astype(np.int64), val


def test_lattice_and_clique_structure_detection():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import instances as inst
    from spin_glass_anneal_rl_b200.annealing._backend import _clique_groups, _lattice_bonds

    # 2D +-J lattice, open and periodic
    for periodic in (False, True):
        Jx, Jy = inst.ea_lattice_bonds(12, seed=3, periodic=periodic)
        rowptr, colidx, val, h = inst.lattice_csr(Jx, Jy)
        rows, cols, vals = _coo(144, rowptr, colidx, val)
        got = _lattice_bonds(rows, cols, vals, h, 144)
        assert got is not None and np.array_equal(got[0], Jx) and np.array_equal(got[1], Jy)
        assert _clique_groups(rows, cols, vals, 144) is None
    # not a lattice: a field, a non-unit coupling, an asymmetric entry, a long-range bond
    Jx, Jy = inst.ea_lattice_bonds(12, seed=3)
    rowptr, colidx, val, h = inst.lattice_csr(Jx, Jy)
    rows, cols, vals = _coo(144, rowptr, colidx, val)
    assert _lattice_bonds(rows, cols, vals, h + 1.0, 144) is None
    v2 = vals.copy(); v2[0] = 2.0
    assert _lattice_bonds(rows, cols, v2, h, 144) is None
    assert _lattice_bonds(rows[1:], cols[1:], vals[1:], h, 144) is None
    c2 = cols.copy(); c2[0] = 77
    assert _lattice_bonds(rows, c2, vals, h, 144) is None

    # block cliques (scheduler) and things that are not
    rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(7, 5, seed=2))
    rows, cols, vals = _coo(35, rowptr, colidx, val)
    group_of, coupling = _clique_groups(rows, cols, vals, 35)
    assert np.array_equal(group_of, np.arange(35) // 5) and np.all(coupling == 50.0)
    assert _lattice_bonds(rows, cols, vals, np.zeros(35, np.float32), 35) is None
    v3 = vals.copy(); v3[3] = 49.0
    assert _clique_groups(rows, cols, v3, 35) is None               # not constant inside a group
    assert _clique_groups(rows[:-1], cols[:-1], vals[:-1], 35) is None   # a missing pair
    # two overlapping clique families (TSP rows and columns) are not disjoint groups
    J, _ = inst.tsp_ising(inst.random_tsp(5, 1))
    r, c = np.nonzero(J)
    assert _clique_groups(r, c, J[r, c], 25) is None


def test_engine_cache_signature_survives_id_reuse():
    """ids / data pointers of freed coupling tensors are handed out again, so the
    engine cache must hold the tensors it was built from and compare by identity + version."""
    from spin_glass_anneal_rl_b200.annealing._backend import _Signature
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=16, use_sparse=False))
    m.set_couplings_from_matrix(torch.randn(16, 16))
    sig = _Signature(m)
    assert sig.matches(m)
    for _ in range(20):     # free / reallocate same-shaped couplings: the old id comes back
        m.set_couplings_from_matrix(torch.randn(16, 16))
        assert not sig.matches(m)
    m2 = sg.IsingModel(sg.IsingModelConfig(n_spins=16, use_sparse=False))
    s2 = _Signature(m2)
    m2.couplings[0, 1] = 3.0          # in-place edit bumps the version
    assert not s2.matches(m2)
    m3 = sg.IsingModel(sg.IsingModelConfig(n_spins=16, use_sparse=True))
    s3 = _Signature(m3)
    m3.set_coupling(0, 1, 1.0)        # queued write -> new COO tensor on the next read
    assert not s3.matches(m3)
