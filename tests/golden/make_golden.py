#!/usr/bin/env python
"""Record golden traces from the UNMODIFIED reference (device='cpu', dense J).

Run in the build container only (needs /root/reference):

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The reference's own tests hold no golden vectors
for the sweep path, so these traces are the pin: the
oracle (oracle/) must reproduce every one of them from the seed alone, and the
CUDA path is then compared with the oracle and with these files.

Each fixture stores the instance (J, h, initial spins, config), the reference's
outputs, how many raw mt19937 words the run consumed, and the first attempts'
(site, uniform, p) draws for a direct check of the stream model.
"""
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)

from spin_glass_rl.core.ising_model import IsingModel, IsingModelConfig  # noqa: E402
from spin_glass_rl.core.spin_dynamics import UpdateRule  # noqa: E402
from spin_glass_rl.annealing.gpu_annealer import GPUAnnealer, GPUAnnealerConfig  # noqa: E402
from spin_glass_rl.annealing.parallel_tempering import (  # noqa: E402
    ParallelTempering, ParallelTemperingConfig)
from spin_glass_rl.annealing.temperature_scheduler import ScheduleType  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


class Recorder:
    """Counts/logs the scalar draws the sweep makes from torch's global RNG."""

    def __init__(self, keep=4096):
        self.keep = keep
        self.reset()
        self._randint, self._rand, self._exp = torch.randint, torch.rand, torch.exp

    def reset(self):
        self.n_randint = self.n_rand = self.n_vec_randint = 0
        self.sites, self.uniforms, self.probs = [], [], []

    def __enter__(self):
        rec = self

        def randint(*a, **k):
            out = rec._randint(*a, **k)
            if out.numel() == 1:
                rec.n_randint += 1
                if len(rec.sites) < rec.keep:
                    rec.sites.append(int(out.item()))
            else:
                rec.n_vec_randint += out.numel()
            return out

        def rand(*a, **k):
            out = rec._rand(*a, **k)
            rec.n_rand += out.numel()
            if len(rec.uniforms) < rec.keep:
                rec.uniforms.append(float(out.item()))
            return out

        def exp(x, *a, **k):
            out = rec._exp(x, *a, **k)
            if out.numel() == 1 and len(rec.probs) < rec.keep:
                rec.probs.append((float(x.item()), float(out.item())))
            return out

        torch.randint, torch.rand, torch.exp = randint, rand, exp
        return self

    def __exit__(self, *exc):
        torch.randint, torch.rand, torch.exp = self._randint, self._rand, self._exp

    @property
    def raw_consumed(self):
        return self.n_randint + self.n_rand + self.n_vec_randint


def load(name):
    return np.load(os.path.join(OUT, name + ".npz"))


def make_model(J, h, spins):
    m = IsingModel(IsingModelConfig(n_spins=J.shape[0], use_sparse=False))
    m.set_couplings_from_matrix(torch.from_numpy(J.copy()))
    m.set_external_fields(torch.from_numpy(h.copy()))
    m.set_spins(torch.from_numpy(spins.copy()))
    return m


def sym_pm1(n, rng):
    a = rng.integers(0, 2, size=(n, n)) * 2 - 1
    J = np.triu(a, 1)
    return (J + J.T).astype(np.float32)


def sym_gauss(n, rng, scale=1.0):
    a = rng.standard_normal((n, n)) * scale
    J = (a + a.T) / 2
    np.fill_diagonal(J, 0.0)
    return J.astype(np.float32)


def cfg1_instance():
    """N=100, torch generator seed 1001."""
    n = 100
    g = torch.Generator().manual_seed(1001)
    A = torch.randn(n, n, generator=g)
    J = (A + A.T) / 2
    J.fill_diagonal_(0.0)
    h = 0.5 * torch.randn(n, generator=g)
    s = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
    return J.numpy().astype(np.float32), h.numpy().astype(np.float32), s.numpy()


def sk_instance(n, seed=3003):
    """cfg3 generator shape (research/experimental_validation.py:112-131)."""
    rs = np.random.RandomState(seed)
    G = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n))
    J = (G + G.T) / 2
    np.fill_diagonal(J, 0.0)
    s = rs.randint(0, 2, size=n) * 2.0 - 1.0
    return J.astype(np.float32), np.zeros(n, np.float32), s.astype(np.float32)


def run_sa(name, J, h, spins0, *, seed, n_sweeps, T0, Tf, schedule="geometric", params=None,
           record_interval=10, tol=1e-8, rule="metropolis"):
    params = params if params is not None else {"alpha": 0.95}
    model = make_model(J, h, spins0)
    cfg = GPUAnnealerConfig(n_sweeps=n_sweeps, initial_temp=T0, final_temp=Tf,
                            schedule_type=ScheduleType(schedule), schedule_params=dict(params),
                            record_interval=record_interval, energy_tolerance=tol,
                            random_seed=seed)
    ann = GPUAnnealer(cfg)  # seeds torch + numpy here
    with Recorder() as rec:
        res = ann.anneal(model, UpdateRule(rule))
    cfgd = dict(kind="sa", seed=seed, n_sweeps=n_sweeps, T0=T0, Tf=Tf, schedule=schedule,
                params=params, record_interval=record_interval, tol=tol, rule=rule)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), config=json.dumps(cfgd), J=J, h=h, spins0=spins0,
        best_energy=np.float64(res.best_energy),
        best_configuration=res.best_configuration.numpy().astype(np.int8),
        energy_history=np.array(res.energy_history, np.float64),
        temperature_history=np.array(res.temperature_history, np.float64),
        acceptance_rate_history=np.array(res.acceptance_rate_history, np.float64),
        n_sweeps_done=np.int64(res.n_sweeps),
        convergence_sweep=np.int64(-1 if res.convergence_sweep is None else res.convergence_sweep),
        energy_std=np.float64(res.energy_std),
        final_spins=model.spins.numpy().astype(np.int8),
        raw_consumed=np.int64(rec.raw_consumed), n_rand=np.int64(rec.n_rand),
        head_sites=np.array(rec.sites, np.int32), head_uniforms=np.array(rec.uniforms, np.float32),
        head_probs=np.array(rec.probs, np.float64).reshape(-1, 2),
        torch_version=torch.__version__)
    print(f"{name}: N={J.shape[0]} sweeps={res.n_sweeps} best={res.best_energy:.6f} "
          f"raw={rec.raw_consumed}")


def run_pt(name, J, h, *, seed, n_replicas, n_sweeps, tmin, tmax, dist="geometric",
           exchange_interval=5, record_interval=5, rule="metropolis",
           method="nearest_neighbor"):
    n = J.shape[0]
    model = make_model(J, h, np.ones(n, np.float32))
    cfg = ParallelTemperingConfig(n_replicas=n_replicas, n_sweeps=n_sweeps, temp_min=tmin,
                                  temp_max=tmax, temp_distribution=dist,
                                  exchange_interval=exchange_interval, n_threads=1,
                                  exchange_method=method,
                                  record_interval=record_interval, random_seed=seed)
    pt = ParallelTempering(cfg)
    with Recorder() as rec:
        res = pt.run(model, UpdateRule(rule))
    cfgd = dict(kind="pt", seed=seed, n_replicas=n_replicas, n_sweeps=n_sweeps, tmin=tmin,
                tmax=tmax, dist=dist, exchange_interval=exchange_interval,
                record_interval=record_interval, rule=rule)
    if method != "nearest_neighbor":
        cfgd["method"] = method
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), config=json.dumps(cfgd), J=J, h=h,
        best_energy=np.float64(res.best_energy),
        best_configuration=res.best_configuration.numpy().astype(np.int8),
        energy_history=np.array(res.energy_history, np.float64),
        temperature_history=np.array(res.temperature_history, np.float64),
        acceptance_rate_history=np.array(res.acceptance_rate_history, np.float64),
        temperatures=np.array(pt.temperatures, np.float64),
        energy_histories=np.array(pt.energy_histories, np.float64),
        exchange_attempts=pt.exchange_attempts, exchange_accepts=pt.exchange_accepts,
        final_spins=np.stack([r.spins.numpy() for r in pt.replicas]).astype(np.int8),
        raw_consumed=np.int64(rec.raw_consumed), torch_version=torch.__version__)
    print(f"{name}: N={n} R={n_replicas} best={res.best_energy:.6f} raw={rec.raw_consumed} "
          f"acc={pt.exchange_accepts.sum():.0f}/{pt.exchange_attempts.sum():.0f}")


def run_kat(name, J, h, batch, seed):
    """Known-answer energies / local fields from the reference model itself."""
    rs = np.random.RandomState(seed)
    n = J.shape[0]
    S = (rs.randint(0, 2, size=(batch, n)) * 2 - 1).astype(np.float32)
    E = np.zeros(batch)
    F = np.zeros((batch, n))
    dE = np.zeros((batch, n))
    for b in range(batch):
        m = make_model(J, h, S[b])
        E[b] = m.compute_energy()
        for i in range(n):
            F[b, i] = m.get_local_field(i)
        for i in range(0, n, max(1, n // 8)):
            mm = make_model(J, h, S[b])
            dE[b, i] = mm.flip_spin(i)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), J=J, h=h, S=S.astype(np.int8), E=E, F=F,
                        dE=dE, torch_version=torch.__version__)
    print(f"{name}: batch={batch} N={n}")


def run_ec_kat(name, J, h, batch, seed):
    """Known answers from the reference's EnergyComputer (core/energy_computer.py) itself."""
    from spin_glass_rl.core.energy_computer import ComputeMode, EnergyComputer
    rs = np.random.RandomState(seed)
    n = J.shape[0]
    S = (rs.randint(0, 2, size=(batch, n)) * 2 - 1).astype(np.float32)
    m = make_model(J, h, S[0])
    out = {}
    for mode in ComputeMode:
        ec = EnergyComputer(m, mode)
        out["total_" + mode.value] = np.float64(ec.compute_total_energy())
    ec = EnergyComputer(m, ComputeMode.FULL)
    st = ec.compute_energy_stats()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), J=J, h=h, S=S.astype(np.int8),
        dE=np.array([ec.compute_energy_change(i) for i in range(n)], np.float64),
        stats=np.array([st.total_energy, st.interaction_energy, st.field_energy], np.float64),
        per_spin=st.per_spin_energy.numpy().astype(np.float64),
        gradient=ec.compute_energy_gradient().numpy().astype(np.float64),
        batch=ec.compute_batch_energies(torch.from_numpy(S)).numpy().astype(np.float64),
        total_other=np.float64(ec.compute_total_energy(torch.from_numpy(S[1]))),
        torch_version=torch.__version__, **out)
    print(f"{name}: batch={batch} N={n} E={out['total_full']:.4f}")


def main():
    rng = np.random.default_rng(20261018)

    # --- simulated annealing, integer couplings (bit-exact targets)
    n = 48
    J = sym_pm1(n, rng)
    h = rng.integers(-1, 2, size=n).astype(np.float32)
    s0 = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    run_sa("sa_pm1_n48", J, h, s0, seed=11, n_sweeps=80, T0=5.0, Tf=0.05,
           params={"alpha": 0.93}, record_interval=5)

    Jc, hc, sc = cfg1_instance()
    run_sa("sa_cfg1_float_n100", Jc, hc, sc, seed=1234, n_sweeps=400, T0=5.0, Tf=0.01,
           params={"alpha": 0.95}, record_interval=10)
    Ji = sym_pm1(100, rng)
    run_sa("sa_cfg1_int_n100", Ji, np.zeros(100, np.float32), sc, seed=4321, n_sweeps=200,
           T0=5.0, Tf=0.01, params={"alpha": 0.95}, record_interval=10)

    # --- every schedule type, short integer runs
    n = 24
    J = sym_pm1(n, rng)
    h = np.zeros(n, np.float32)
    s0 = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    for sched, params in [("linear", {}), ("exponential", {}), ("logarithmic", {"c": 2.0}),
                          ("power_law", {"k": 0.7}), ("fast", {}), ("boltzmann", {}),
                          ("adaptive", {"alpha": 0.9, "adaptation_window": 5,
                                        "target_acceptance": 0.3})]:
        run_sa(f"sa_sched_{sched}_n24", J, h, s0, seed=77, n_sweeps=40, T0=4.0, Tf=0.2,
               schedule=sched, params=params, record_interval=3)

    # --- Glauber / heat bath
    n = 32
    J = sym_pm1(n, rng)
    h = rng.integers(-2, 3, size=n).astype(np.float32)
    s0 = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    run_sa("sa_glauber_int_n32", J, h, s0, seed=5, n_sweeps=50, T0=3.0, Tf=0.3,
           params={"alpha": 0.95}, record_interval=4, rule="glauber")
    run_sa("sa_heatbath_int_n32", J, h, s0, seed=6, n_sweeps=50, T0=3.0, Tf=0.3,
           params={"alpha": 0.95}, record_interval=4, rule="heat_bath")
    Jg = sym_gauss(32, rng, 0.5)
    hg = (0.3 * rng.standard_normal(32)).astype(np.float32)
    run_sa("sa_glauber_float_n32", Jg, hg, s0, seed=8, n_sweeps=40, T0=2.0, Tf=0.2,
           params={"alpha": 0.95}, record_interval=4, rule="glauber")

    # --- early stop through _check_convergence (record every sweep, frozen system)
    n = 12
    J = sym_pm1(n, rng)
    s0 = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    run_sa("sa_converge_n12", J, np.zeros(n, np.float32), s0, seed=3, n_sweeps=400, T0=0.4,
           Tf=0.01, params={"alpha": 0.9}, record_interval=1)

    # --- SK float (cfg3 generator, down-scaled) and a wider integer case
    Js, hs, ss = sk_instance(256)
    run_sa("sa_sk_float_n256", Js, hs, ss, seed=42, n_sweeps=30, T0=2.0, Tf=0.05,
           params={"alpha": 0.85}, record_interval=5)
    Jw = sym_pm1(1024, rng)
    sw = (rng.integers(0, 2, size=1024) * 2 - 1).astype(np.float32)
    run_sa("sa_pm1_n1024", Jw, np.zeros(1024, np.float32), sw, seed=9, n_sweeps=4, T0=20.0,
           Tf=5.0, params={"alpha": 0.7}, record_interval=1)

    # --- parallel tempering (n_threads=1: deterministic)
    J = sym_pm1(40, rng)
    run_pt("pt_int_n40_r6", J, np.zeros(40, np.float32), seed=21, n_replicas=6, n_sweeps=60,
           tmin=0.5, tmax=6.0)
    Jg = sym_gauss(32, rng, 0.7)
    hg = (0.2 * rng.standard_normal(32)).astype(np.float32)
    run_pt("pt_float_n32_r4", Jg, hg, seed=22, n_replicas=4, n_sweeps=50, tmin=0.3, tmax=4.0,
           dist="linear", exchange_interval=3, record_interval=2)
    run_pt("pt_int_n40_r5_exp", J, np.zeros(40, np.float32), seed=23, n_replicas=5, n_sweeps=40,
           tmin=0.4, tmax=5.0, dist="exponential", exchange_interval=4, record_interval=4)

    # --- the all_pairs exchange (CPU branch of _all_pairs_exchange, parallel_tempering.py:228-232)
    run_pt("pt_int_n40_r6_allpairs", J, np.zeros(40, np.float32), seed=24, n_replicas=6,
           n_sweeps=60, tmin=0.5, tmax=6.0, exchange_interval=3, record_interval=3,
           method="all_pairs")

    # --- known-answer energies / local fields
    run_kat("kat_energy_int_n64", sym_pm1(64, rng), rng.integers(-1, 2, size=64).astype(np.float32),
            6, 100)
    run_kat("kat_energy_float_n100", Jc, hc, 4, 101)
    # --- the same through the reference's EnergyComputer
    run_ec_kat("ec_int_n64", load("kat_energy_int_n64")["J"], load("kat_energy_int_n64")["h"], 6, 102)
    run_ec_kat("ec_float_n100", Jc, hc, 4, 103)


if __name__ == "__main__":
    main()
