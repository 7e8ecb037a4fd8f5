"""AnnealingResult: the return container of anneal() / run().

Field for field the reference's dataclass (reference annealing/result.py:9-77) with the
same validation and derived statistics, because callers read ``best_configuration``
(a float32 +-1 CPU tensor: ``(spins + 1) // 2`` in problems/routing.py:330-389),
``best_energy``, ``convergence_sweep`` and ``total_time`` (problems/base.py:136-144).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch


@dataclass
class AnnealingResult:
    best_configuration: torch.Tensor
    best_energy: float
    energy_history: List[float]
    temperature_history: List[float]
    acceptance_rate_history: List[float]
    total_time: float
    n_sweeps: int
    convergence_sweep: Optional[int] = None
    final_temperature: float = 0.0
    final_acceptance_rate: float = 0.0
    energy_std: float = 0.0
    algorithm: str = "simulated_annealing"
    device: str = "cpu"
    random_seed: Optional[int] = None

    def __post_init__(self):
        if not isinstance(self.best_configuration, torch.Tensor):
            raise TypeError("best_configuration must be a torch.Tensor")
        if not isinstance(self.best_energy, (int, float)):
            raise TypeError("best_energy must be a numeric value")
        if np.isnan(self.best_energy) or np.isinf(self.best_energy):
            raise ValueError("best_energy contains invalid values (NaN or Inf)")
        if self.total_time < 0:
            raise ValueError("total_time must be non-negative")
        if self.n_sweeps <= 0:
            raise ValueError("n_sweeps must be positive")
        if self.energy_history:
            e = np.asarray(self.energy_history, dtype=np.float64)
            if not np.all(np.isfinite(e)):
                raise ValueError("energy_history contains invalid values (NaN or Inf)")
            self.energy_std = float(np.std(e))
            if len(e) > 10:  # first window whose spread drops below 1 % of |best energy|
                w = min(50, len(e) // 4)
                for i in range(w, len(e)):
                    if np.std(e[i - w:i]) < 0.01 * abs(self.best_energy):
                        self.convergence_sweep = i - w
                        break
        if self.temperature_history:
            self.final_temperature = self.temperature_history[-1]
        if self.acceptance_rate_history:
            self.final_acceptance_rate = self.acceptance_rate_history[-1]

    _FIELDS = ("best_energy", "energy_history", "temperature_history", "acceptance_rate_history",
               "total_time", "n_sweeps", "convergence_sweep", "final_temperature",
               "final_acceptance_rate", "energy_std", "algorithm", "device", "random_seed")

    def get_summary(self) -> Dict:
        keys = ("best_energy", "total_time", "n_sweeps", "convergence_sweep", "final_temperature",
                "final_acceptance_rate", "energy_std", "algorithm", "device")
        return {k: getattr(self, k) for k in keys}

    def save(self, filepath: str) -> None:
        data = {k: getattr(self, k) for k in self._FIELDS}
        np.savez_compressed(filepath, best_configuration=self.best_configuration.cpu().numpy(), **data)

    @classmethod
    def load(cls, filepath: str) -> "AnnealingResult":
        d = np.load(filepath, allow_pickle=True)

        def opt_int(v):
            v = v.item() if hasattr(v, "item") else v
            return None if v is None else int(v)

        return cls(
            best_configuration=torch.from_numpy(d["best_configuration"]),
            best_energy=float(d["best_energy"]), energy_history=d["energy_history"].tolist(),
            temperature_history=d["temperature_history"].tolist(),
            acceptance_rate_history=d["acceptance_rate_history"].tolist(),
            total_time=float(d["total_time"]), n_sweeps=int(d["n_sweeps"]),
            convergence_sweep=opt_int(d["convergence_sweep"]),
            final_temperature=float(d["final_temperature"]),
            final_acceptance_rate=float(d["final_acceptance_rate"]),
            energy_std=float(d["energy_std"]), algorithm=str(d["algorithm"]),
            device=str(d["device"]), random_seed=opt_int(d["random_seed"]))

    def __repr__(self) -> str:
        return (f"AnnealingResult(best_energy={self.best_energy:.6f}, n_sweeps={self.n_sweeps}, "
                f"time={self.total_time:.3f}s, converged_at={self.convergence_sweep})")
