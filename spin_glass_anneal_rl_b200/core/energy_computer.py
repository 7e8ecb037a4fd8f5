"""EnergyComputer on the B200 field / energy kernel (K2).

Drop-in for the reference's ``EnergyComputer(model, mode)`` (reference core/energy_computer.py:
29-330): same methods, arguments and return types.  Every quantity the reference computes with
per-site Python loops over ``torch.dot`` -- total energy (:50-69, :151-196), the energy change of
a flip (:71-87), the per-spin decomposition (:89-117, :217-231), the gradient (:119-140) and
batch energies (:142-158, a loop over configurations upstream) -- comes from ONE pass of the
tensor-core field kernel here: ``F = S J^T + h`` for the whole batch (sg_batch_energies), then
elementwise arithmetic.  Integer couplings give exact results; float couplings agree with the
reference's float32 dot products to ~1e-6 relative.  The three ``ComputeMode`` values are kept
(callers select them); they share the one device path, INCREMENTAL additionally keeps the
reference's cached value semantics (:160-167, :298-301).  No CPU fallback.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from enum import Enum
from typing import Optional

import numpy as np
import torch


class ComputeMode(Enum):
    FULL = "full"
    INCREMENTAL = "incremental"
    VECTORIZED = "vectorized"


@dataclass
class EnergyStats:
    total_energy: float
    interaction_energy: float
    field_energy: float
    per_spin_energy: Optional[torch.Tensor] = None
    energy_distribution: Optional[torch.Tensor] = None


class EnergyComputer:
    def __init__(self, model, mode: ComputeMode = ComputeMode.FULL):
        self.model = model
        self.mode = mode
        self._cached_energy = None
        self._cache_valid = False
        self._spin_contributions = None
        self._interaction_matrix = None
        self._field_vector = None
        self._matrix_valid = False

    # ------------------------------------------------------------------ device pass
    def _fields_energies(self, spin_configs: torch.Tensor):
        """(E [B], F [B, n]) on the engine's device for configurations [B, n] in {-1, +1}."""
        from ..annealing._backend import engine_for
        eng = engine_for(self.model, getattr(self.model, "device_index", 0))
        s8 = torch.where(spin_configs.to(eng.device) >= 0, 1, -1).to(torch.int8).reshape(-1, self.model.n_spins)
        e, f = eng.batch_energies(s8, want_fields=True)
        return e, f, s8

    def _spins(self, spins):
        return self.model.spins if spins is None else spins

    # ------------------------------------------------------------------ reference API
    def compute_total_energy(self, spins: Optional[torch.Tensor] = None) -> float:
        spins = self._spins(spins)
        if self.mode == ComputeMode.INCREMENTAL:
            if not self._cache_valid or self._cached_energy is None:
                self._cached_energy = float(self._fields_energies(spins)[0][0].item())
                self._cache_valid = True
            return self._cached_energy
        return float(self._fields_energies(spins)[0][0].item())

    def compute_energy_change(self, flip_site: int) -> float:
        """dE = 2 s_i (sum_j J_ij s_j + h_i) if spin ``flip_site`` were flipped (:71-87)."""
        if not (0 <= flip_site < self.model.n_spins):
            raise IndexError(f"flip_site {flip_site} out of range")
        _, f, s8 = self._fields_energies(self.model.spins)
        return float(2.0 * s8[0, flip_site].item() * f[0, flip_site].item())

    def compute_energy_changes(self) -> torch.Tensor:
        """All n single-flip energy changes of the current configuration in one pass (what
        ``compute_energy_change`` returns for every site)."""
        _, f, s8 = self._fields_energies(self.model.spins)
        return (2.0 * s8[0].to(torch.float32) * f[0]).to(self.model.spins.device)

    def compute_energy_stats(self, spins: Optional[torch.Tensor] = None) -> EnergyStats:
        spins = self._spins(spins)
        _, f, s8 = self._fields_energies(spins)
        s = s8[0].to(torch.float64)
        F = f[0].to(torch.float64)
        h = self.model.external_fields.to(device=F.device, dtype=torch.float64)
        interaction = float((-0.5 * (s * (F - h)).sum()).item())      # -1/2 s^T J s      (:184-196)
        field = float((-(h * s).sum()).item())                         # -h^T s            (:198-200)
        # per spin (:217-231): -1/2 s_i (local field incl. h_i)  -  h_i s_i
        per_spin = (-0.5 * s * F - h * s).to(torch.float32).to(self.model.spins.device)
        return EnergyStats(total_energy=interaction + field, interaction_energy=interaction,
                           field_energy=field, per_spin_energy=per_spin)

    def compute_energy_gradient(self, spins: Optional[torch.Tensor] = None) -> torch.Tensor:
        """dE/ds_i = -(sum_j J_ij s_j + h_i) (:119-140)."""
        spins = self._spins(spins)
        _, f, _ = self._fields_energies(spins)
        return (-f[0]).to(torch.float32).to(spins.device)

    def compute_batch_energies(self, spin_configs: torch.Tensor) -> torch.Tensor:
        """Energies of a batch [B, n] (:142-158): one K2 launch instead of B x n dot products."""
        e, _, _ = self._fields_energies(spin_configs)
        return e.to(self.model.spins.device)

    # ------------------------------------------------------------------ cache / mode bookkeeping
    def invalidate_cache(self) -> None:
        self._cache_valid = False
        self._matrix_valid = False
        self._cached_energy = None
        self._spin_contributions = None

    def update_incremental_cache(self, flip_site: int, delta_energy: float) -> None:
        if self._cache_valid and self._cached_energy is not None:
            self._cached_energy += delta_energy

    def set_mode(self, mode: ComputeMode) -> None:
        self.mode = mode
        if mode != ComputeMode.INCREMENTAL:
            self._cache_valid = False

    def benchmark_modes(self, n_trials: int = 100) -> dict:
        results = {}
        for mode in ComputeMode:
            self.set_mode(mode)
            times = []
            for _ in range(n_trials):
                t0 = time.time()
                _ = self.compute_total_energy()
                times.append(time.time() - t0)
            results[mode.value] = {"mean_time": np.mean(times), "std_time": np.std(times),
                                   "min_time": np.min(times), "max_time": np.max(times)}
        return results

    def __repr__(self) -> str:
        return (f"EnergyComputer(mode={self.mode.value}, n_spins={self.model.n_spins}, "
                f"cached={self._cache_valid})")
