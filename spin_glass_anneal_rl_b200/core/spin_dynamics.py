"""UpdateRule and a SpinDynamics facade over the GPU sweep.

``UpdateRule`` matches the reference enum (reference core/spin_dynamics.py:11-16).
``SpinDynamics`` keeps the reference's single-model interface (``sweep``,
``set_temperature``, ``get_acceptance_rate``, ``n_accepted/n_rejected``,
``energy_history``, ``run_dynamics``; :19-152, :325-359) but every sweep is one launch of
the CUDA sweep kernel on one replica; the per-attempt Python loop of the reference
(:73-94) does not exist here.  WOLFF (:193-262) runs the cluster kernel (csrc/sg_wolff.cu):
n cluster updates per sweep, every cluster site counted as accepted, nothing as rejected.
"""
from __future__ import annotations

from enum import Enum
from typing import Optional

import numpy as np
import torch


class UpdateRule(Enum):
    METROPOLIS = "metropolis"
    GLAUBER = "glauber"
    HEAT_BATH = "heat_bath"
    WOLFF = "wolff"


class SpinDynamics:
    def __init__(self, model, temperature: float = 1.0,
                 update_rule: UpdateRule = UpdateRule.METROPOLIS, random_seed: Optional[int] = None):
        self.model = model
        self.temperature = temperature
        self.update_rule = update_rule
        if random_seed is not None:
            torch.manual_seed(random_seed)
            np.random.seed(random_seed)
        self._seed = int(random_seed) if random_seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        self._sweeps_done = 0
        self.n_accepted = 0
        self.n_rejected = 0
        self.energy_history = []
        self.magnetization_history = []

    def set_temperature(self, temperature: float) -> None:
        self.temperature = max(temperature, 1e-10)

    def single_spin_update(self, site: Optional[int] = None):
        """The reference's one-attempt entry point (core/spin_dynamics.py:61-71).  The device path has
        no per-attempt call -- a kernel launch per spin is the pattern this package replaces -- so this
        says so instead of quietly doing the attempt on the host."""
        raise NotImplementedError("single_spin_update: one attempt per call is not offered on the GPU path; "
                                  "use sweep() (n attempts per launch)")

    def sweep(self, n_sweeps: int = 1) -> float:
        """``n_sweeps`` Monte Carlo sweeps (n attempts each) on the GPU; returns the energy."""
        from ..annealing._backend import engine_for, require_dense_for_wolff, rule_name
        eng = engine_for(self.model)
        wolff = self.update_rule == UpdateRule.WOLFF
        if wolff:
            require_dense_for_wolff(eng)
        n = self.model.n_spins
        eng.alloc_replicas(1) if eng.n_replicas != 1 else None
        eng.set_spins(self.model.spins.reshape(1, n))
        eng.init_fields()
        acc0 = int(eng.accepted().item())
        trace = eng.sweep(n_sweeps, np.array([max(self.temperature, 1e-10)]),
                          rule=rule_name(self.update_rule), site_order="random", seed=self._seed,
                          sweep_base=self._sweeps_done, energy_trace=True, track_best=False)
        self._sweeps_done += n_sweeps
        spins = eng.spins()[0].to(torch.float32).cpu()
        acc = int(eng.accepted().item()) - acc0
        self.model.spins = spins.to(self.model.device)
        self.model._invalidate_cache()
        self.n_accepted += acc
        self.n_rejected += 0 if wolff else n_sweeps * n - acc
        energies = trace[:, 0].cpu().tolist()
        self.energy_history.extend(energies)
        self.magnetization_history.append(float(spins.sum().item()))
        return float(energies[-1])

    def run_dynamics(self, n_sweeps: int, record_interval: int = 1) -> dict:
        initial = self.model.compute_energy()
        self.sweep(n_sweeps)
        return {"initial_energy": initial, "final_energy": self.model.compute_energy(),
                "energy_history": self.energy_history.copy(),
                "acceptance_rate": self.get_acceptance_rate(), "n_sweeps": n_sweeps,
                "temperature": self.temperature}

    def get_acceptance_rate(self) -> float:
        total = self.n_accepted + self.n_rejected
        return self.n_accepted / total if total else 0.0

    @property
    def accepted_flips(self) -> int:
        return self.n_accepted

    @accepted_flips.setter
    def accepted_flips(self, value: int) -> None:
        self.n_accepted = value

    @property
    def total_flips(self) -> int:
        return self.n_accepted + self.n_rejected

    @total_flips.setter
    def total_flips(self, value: int) -> None:
        if value < self.n_accepted:
            raise ValueError("Total flips cannot be less than accepted flips")
        self.n_rejected = value - self.n_accepted

    def reset_statistics(self) -> None:
        self.n_accepted = 0
        self.n_rejected = 0
        self.energy_history = []

    # ---- analysis of the recorded histories (reference core/spin_dynamics.py:361-421; host side)
    def get_autocorrelation_time(self, observable: str = "energy") -> float:
        """First lag (in recorded sweeps) at which the normalised autocorrelation of the history
        falls below 1/e; inf for fewer than 10 records, the history length if it never does."""
        series = {"energy": self.energy_history, "magnetization": self.magnetization_history}.get(observable)
        if series is None:
            raise ValueError(f"Unknown observable: {observable}")
        x = np.asarray(series, dtype=np.float64)
        if x.size < 10:
            return float("inf")
        d = x - x.mean()
        acf = np.correlate(d, d, mode="full")[x.size - 1:]
        acf = acf / acf[0]
        below = np.flatnonzero(acf < 1.0 / np.e)
        return float(below[0]) if below.size else float(acf.size)

    def thermal_equilibrium_check(self, window_size: int = 100) -> bool:
        """Two-sample t-test between the last two windows of the energy history: True when their
        means do not differ at the 5 % level (False while fewer than two windows are recorded)."""
        if len(self.energy_history) < 2 * window_size:
            return False
        recent = np.asarray(self.energy_history[-window_size:], dtype=np.float64)
        older = np.asarray(self.energy_history[-2 * window_size:-window_size], dtype=np.float64)
        from scipy import stats
        return bool(stats.ttest_ind(recent, older).pvalue > 0.05)

    def __repr__(self) -> str:
        return (f"SpinDynamics(temperature={self.temperature:.4f}, "
                f"update_rule={self.update_rule.value}, "
                f"acceptance_rate={self.get_acceptance_rate():.4f})")
