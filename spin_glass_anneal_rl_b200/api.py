"""Convenience entry points around the annealers.

* ``anneal(model, n_replicas, n_sweeps, beta_schedule)`` -- the call shape the reference's
  README advertises (reference README.md:75-83: ``scheduler.anneal(ising_model,
  n_replicas=1000, n_sweeps=10000, beta_schedule='geometric')`` returning an object with
  ``.best_configuration``); the class it names does not exist upstream.
* ``batch_energies`` / ``batch_local_fields`` -- BatchProcessor.process_batch_energies and
  VectorizedOperations.vectorized_local_fields
  (reference optimization/high_performance_computing.py:98-165, 357-372) on the GPU.
* ``VectorizedOperations`` / ``BatchProcessor`` -- the reference's class names for the same.
* ``install_as_spin_glass_rl`` -- sys.modules aliases for the reference's import paths.
"""
from __future__ import annotations

import sys
import types
from typing import Optional, Sequence, Union

import numpy as np
import torch


def anneal(ising_model, n_replicas: int = 1, n_sweeps: int = 1000,
           beta_schedule: Union[str, Sequence[float]] = "geometric", *, initial_temp: float = 10.0,
           final_temp: float = 0.01, random_seed: Optional[int] = None, update_rule=None,
           **schedule_params):
    """Anneal ``n_replicas`` replicas for ``n_sweeps`` sweeps; returns an AnnealingResult.

    ``beta_schedule`` is a schedule name ("geometric", "linear", ...) or an explicit
    sequence of inverse temperatures, one per sweep."""
    from .annealing.gpu_annealer import GPUAnnealer, GPUAnnealerConfig
    from .annealing.temperature_scheduler import ScheduleType
    from .core.spin_dynamics import UpdateRule
    params = dict(schedule_params)
    if isinstance(beta_schedule, str):
        kind = ScheduleType(beta_schedule)
        if kind is ScheduleType.GEOMETRIC and "alpha" not in params:
            # reach final_temp at the last sweep instead of the reference's fixed 0.95
            params["alpha"] = float((final_temp / initial_temp) ** (1.0 / max(1, n_sweeps - 1)))
    else:
        betas = np.asarray(beta_schedule, dtype=np.float64)
        if betas.shape != (n_sweeps,):
            raise ValueError("an explicit beta_schedule needs one value per sweep")
        temps = 1.0 / np.maximum(betas, 1e-300)
        kind = ScheduleType.CUSTOM
        params["custom_func"] = lambda s: float(temps[min(s, n_sweeps - 1)])
        initial_temp, final_temp = float(temps[0]), float(min(temps.min(), final_temp))
    cfg = GPUAnnealerConfig(n_sweeps=n_sweeps, initial_temp=initial_temp, final_temp=final_temp,
                            schedule_type=kind, schedule_params=params, random_seed=random_seed,
                            n_replicas=n_replicas, record_interval=max(1, n_sweeps // 100))
    return GPUAnnealer(cfg).anneal(ising_model, update_rule or UpdateRule.METROPOLIS)


def _engine_for_tensors(couplings: torch.Tensor, external_fields: torch.Tensor):
    from .engine import Engine
    eng = Engine(0)
    J = couplings.to_dense() if couplings.is_sparse else couplings
    eng.set_model(J.to(torch.float32), external_fields.to(torch.float32))
    return eng


def batch_energies(spin_configurations: torch.Tensor, couplings: torch.Tensor,
                   external_fields: torch.Tensor) -> torch.Tensor:
    """E[b] = -1/2 s_b^T J s_b - h^T s_b for a batch of configurations [B, n] (or [n])."""
    single = spin_configurations.dim() == 1
    S = spin_configurations.reshape(1, -1) if single else spin_configurations
    e = _engine_for_tensors(couplings, external_fields).batch_energies(S.sign().to(torch.int8))
    return e[0] if single else e


def batch_local_fields(spin_configurations: torch.Tensor, couplings: torch.Tensor,
                       external_fields: torch.Tensor) -> torch.Tensor:
    """F[b, i] = sum_j J_ij s_bj + h_i for a batch of configurations [B, n]."""
    _, f = _engine_for_tensors(couplings, external_fields).batch_energies(
        spin_configurations.sign().to(torch.int8), want_fields=True)
    return f


class VectorizedOperations:
    """Reference optimization/high_performance_computing.py:338-386, same static methods.
    The field GEMM runs on the int8 tensor-core kernel (K2-TC); the two index operations are
    element-wise gathers on the caller's tensors."""

    @staticmethod
    def vectorized_spin_flips(spin_configurations: torch.Tensor, flip_indices: torch.Tensor) -> torch.Tensor:
        """Copy of ``spin_configurations`` [B, n] with the spins at ``flip_indices`` [B, k] negated
        (an index listed twice is flipped once, as with the reference's indexed ``*=``)."""
        result = spin_configurations.clone()
        rows = torch.arange(result.shape[0], device=result.device).unsqueeze(1)
        result[rows, flip_indices.to(result.device)] *= -1
        return result

    @staticmethod
    def vectorized_local_fields(spin_configurations: torch.Tensor, couplings: torch.Tensor,
                                external_fields: torch.Tensor) -> torch.Tensor:
        """F[b, i] = sum_j J_ij s_bj + h_i for every configuration of the batch."""
        return batch_local_fields(spin_configurations, couplings, external_fields).to(spin_configurations.dtype)

    @staticmethod
    def vectorized_energy_differences(spin_configurations: torch.Tensor, local_fields: torch.Tensor,
                                      flip_indices: torch.Tensor) -> torch.Tensor:
        """dE[b, k] = 2 s_{b, i} F_{b, i} with i = flip_indices[b, k]."""
        idx = flip_indices.to(local_fields.device)
        s = torch.gather(spin_configurations.to(local_fields.device), 1, idx)
        return 2.0 * s * torch.gather(local_fields, 1, idx)


class BatchProcessor:
    """Energy half of the reference's BatchProcessor
    (optimization/high_performance_computing.py:86-165): ``process_batch_energies``."""

    def __init__(self, config=None):
        self.config = config
        self.batch_size = getattr(config, "batch_size", None)

    def process_batch_energies(self, spin_configurations: torch.Tensor, couplings: torch.Tensor,
                               external_fields: torch.Tensor) -> torch.Tensor:
        return batch_energies(spin_configurations, couplings, external_fields)


def install_as_spin_glass_rl(force: bool = False) -> None:
    """Alias this package's modules to the reference's import paths for the hot path."""
    from . import annealing, core
    from .annealing import (batch_processor, cuda_kernels, gpu_annealer, multi_gpu, parallel_tempering, result,
                            temperature_scheduler)
    from .core import energy_computer, ising_model, spin_dynamics
    from .utils import exceptions
    if "spin_glass_rl" in sys.modules and not force:
        raise RuntimeError("spin_glass_rl is already imported; pass force=True to shadow it")
    root = types.ModuleType("spin_glass_rl")
    root.__path__ = []  # mark as package
    utils = types.ModuleType("spin_glass_rl.utils")
    utils.__path__ = []
    mods = {
        "spin_glass_rl": root, "spin_glass_rl.core": core, "spin_glass_rl.annealing": annealing,
        "spin_glass_rl.utils": utils, "spin_glass_rl.utils.exceptions": exceptions,
        "spin_glass_rl.core.ising_model": ising_model,
        "spin_glass_rl.core.spin_dynamics": spin_dynamics,
        "spin_glass_rl.core.energy_computer": energy_computer,
        "spin_glass_rl.annealing.gpu_annealer": gpu_annealer,
        "spin_glass_rl.annealing.cuda_kernels": cuda_kernels,
        "spin_glass_rl.annealing.batch_processor": batch_processor,
        "spin_glass_rl.annealing.parallel_tempering": parallel_tempering,
        "spin_glass_rl.annealing.multi_gpu": multi_gpu,
        "spin_glass_rl.annealing.temperature_scheduler": temperature_scheduler,
        "spin_glass_rl.annealing.result": result,
    }
    sys.modules.update(mods)
    root.core, root.annealing, root.utils = core, annealing, utils
    utils.exceptions = exceptions
    for name in ("IsingModel", "IsingModelConfig", "SpinDynamics", "UpdateRule"):
        setattr(root, name, getattr(core, name))
    for name in ("GPUAnnealer", "GPUAnnealerConfig", "ParallelTempering", "ParallelTemperingConfig",
                 "AnnealingResult", "ScheduleType", "TemperatureScheduler"):
        setattr(root, name, getattr(annealing, name))
