"""Many models per launch.

Mirror of the reference's ``BatchProcessor`` (reference annealing/batch_processor.py:22-43,
180-288, 423-484, 533-556): ``BatchProcessor(config, annealer_config, device)
.process_models_batch(models, callback) -> List[AnnealingResult]``.  Upstream this is one
``GPUAnnealer.anneal`` per model on a pool of four host threads.  Here models of the same size
with n <= 224 are stacked into ONE engine (``sg_set_model_dense_batch``): their replicas are
model-major and every sweep launch of the small-model kernel covers all models at once, so the
per-model cost of a short anneal (the RL environment's pattern, reference
rl_integration/environment.py:318-336) is a slice of one launch instead of a chain of launches and
host round trips of its own.  Other models (larger, sparse-only, ADAPTIVE schedule) go through
``GPUAnnealer.anneal`` one by one.

Differences from a loop of ``anneal`` calls, both consequences of sharing launches: all models run
the full ``n_sweeps`` (no per-model early stop on convergence), and all use the same site order.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch

from ._backend import as_pm1_float, random_spins, rule_name
from .gpu_annealer import GPUAnnealer, GPUAnnealerConfig
from .result import AnnealingResult
from .temperature_scheduler import TemperatureScheduler
from ..core.spin_dynamics import UpdateRule
from ..utils.exceptions import DeviceError

STACK_LIMIT = 224   # largest n of the small-model kernel (csrc/sg_sweep_small.cu)


@dataclass
class BatchConfig:
    """Reference annealing/batch_processor.py:22-43 (same fields, same validation)."""
    batch_size: int = 32
    max_memory_usage: float = 0.8
    prefetch_batches: int = 2
    use_mixed_precision: bool = False
    enable_gradient_checkpointing: bool = True
    memory_optimization_level: int = 1
    streaming_mode: bool = False
    checkpoint_interval: int = 100

    def __post_init__(self):
        if self.batch_size <= 0:
            raise ValueError("Batch size must be positive")
        if not 0 < self.max_memory_usage <= 1:
            raise ValueError("Max memory usage must be between 0 and 1")
        if self.memory_optimization_level not in [0, 1, 2]:
            raise ValueError("Memory optimization level must be 0, 1, or 2")


def plan_stacks(sizes: List[int], stackable: List[bool], batch_size: int) -> List[List[int]]:
    """Group model indices into launches: models of equal size that can be stacked go together,
    at most ``batch_size`` per group, in first-appearance order; everything else runs alone."""
    groups: List[List[int]] = []
    open_group: Dict[int, List[int]] = {}
    for i, (n, ok) in enumerate(zip(sizes, stackable)):
        if not ok:
            groups.append([i])
            continue
        g = open_group.get(n)
        if g is None or len(g) >= batch_size:
            g = []
            open_group[n] = g
            groups.append(g)
        g.append(i)
    return groups


class BatchProcessor:
    def __init__(self, config: BatchConfig, annealer_config: GPUAnnealerConfig,
                 device: Optional[torch.device] = None):
        self.config = config
        self.annealer_config = annealer_config
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.annealer = GPUAnnealer(annealer_config)
        self.processed_batches = 0
        self.total_processing_time = 0.0
        self.checkpoint_data: Dict[str, Any] = {}
        self._engine = None
        self._launch_seed = 0

    # ------------------------------------------------------------------ public API
    def process_models_batch(self, models: List, callback: Optional[Callable[[List[AnnealingResult]], None]] = None,
                             update_rule: UpdateRule = UpdateRule.METROPOLIS) -> List[AnnealingResult]:
        """Anneal every model with ``annealer_config``; results in the order of ``models``."""
        t0 = time.time()
        cfg = self.annealer_config
        schedule = TemperatureScheduler.create_schedule(
            cfg.schedule_type, cfg.initial_temp, cfg.final_temp, cfg.n_sweeps, **cfg.schedule_params)
        sizes = [int(m.n_spins) for m in models]
        ok = [schedule.stateless and n <= STACK_LIMIT and n >= 2 for n in sizes]
        results: List[Optional[AnnealingResult]] = [None] * len(models)
        for group in plan_stacks(sizes, ok, self.config.batch_size):
            if len(group) == 1:
                results[group[0]] = self.annealer.anneal(models[group[0]], update_rule)
            else:
                for i, res in zip(group, self._anneal_stack([models[i] for i in group], update_rule)):
                    results[i] = res
        if callback:
            callback(results)
        self.total_processing_time += time.time() - t0
        self.processed_batches += 1
        return results

    def process_models_stream(self, model_generator: Callable[[], List], total_batches: int,
                              callback: Optional[Callable[[List[AnnealingResult]], None]] = None) -> None:
        """Reference :290-345: ``total_batches`` calls of ``model_generator`` -> process -> callback."""
        for _ in range(total_batches):
            models = model_generator()
            if not models:
                break
            self.process_models_batch(models, callback)

    def get_processing_stats(self) -> Dict[str, Any]:
        avg = self.total_processing_time / self.processed_batches if self.processed_batches else 0.0
        return {"processed_batches": self.processed_batches,
                "total_processing_time": self.total_processing_time,
                "avg_batch_time": avg, "device": str(self.device)}

    def reset(self) -> None:
        self.processed_batches = 0
        self.total_processing_time = 0.0
        self.checkpoint_data.clear()

    # ------------------------------------------------------------------ the stacked anneal
    def _anneal_stack(self, models: List, update_rule: UpdateRule) -> List[AnnealingResult]:
        """GPUAnnealer.anneal's loop (reference annealing/gpu_annealer.py:127-183) for M stacked
        models: same schedule, same record cadence, best tracking per sweep in the kernel."""
        from ..engine import Engine
        cfg = self.annealer_config
        start = time.time()
        try:
            if self._engine is None:
                self._engine = Engine(cfg.device_index)
            eng = self._engine
        except (RuntimeError, OSError) as exc:
            raise DeviceError(f"B200 annealing engine unavailable: {exc}") from exc
        M, n = len(models), int(models[0].n_spins)
        R = max(1, int(cfg.n_replicas))
        J = torch.stack([(m.couplings.to_dense() if m.couplings.is_sparse else m.couplings).cpu()
                         for m in models]).to(torch.float32)
        h = torch.stack([m.external_fields.cpu() for m in models]).to(torch.float32)
        eng.set_models(J, h)
        eng.alloc_replicas(M * R)       # buffers are reused when the stack has the previous shape
        seed = cfg.random_seed if cfg.random_seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        self._launch_seed += 1
        philox_seed = (int(seed) << 20) ^ (0x80000 + self._launch_seed)
        gen = torch.Generator(device=eng.device)
        gen.manual_seed(int(seed) + 104729 * self._launch_seed)
        spins0 = random_spins(M * R, n, eng.device, gen)
        first = torch.arange(M, device=eng.device) * R          # replica 0 of every model ...
        start_spins = torch.stack([m.spins.cpu() for m in models]).sign().to(torch.int8)
        spins0[first] = start_spins.to(eng.device)                    # ... starts from model.spins
        eng.set_spins(spins0)
        eng.init_fields()

        schedule = TemperatureScheduler.create_schedule(
            cfg.schedule_type, cfg.initial_temp, cfg.final_temp, cfg.n_sweeps, **cfg.schedule_params)
        temps_all = np.maximum(schedule.precompute(cfg.n_sweeps), 1e-10)
        interval = max(1, int(cfg.record_interval))
        e0 = eng.energies()[first].cpu().numpy()
        energy_hist = [[float(e0[k])] for k in range(M)]
        temp_hist = [cfg.initial_temp]
        acc_hist = [[0.0] for _ in range(M)]
        rule = rule_name(update_rule)
        if rule == "wolff":
            raise NotImplementedError("stacked small models run the single-spin rules; anneal a model with "
                                      "UpdateRule.WOLFF through GPUAnnealer")
        done = 0
        while done < cfg.n_sweeps:
            last = done if done % interval == 0 else min(cfg.n_sweeps - 1, (done // interval + 1) * interval)
            chunk_t = temps_all[done:last + 1]
            k = len(chunk_t)
            trace = eng.sweep(k, chunk_t, temps_sweep_stride=1, rule=rule, site_order=cfg.site_order,
                              seed=philox_seed, sweep_base=done, energy_trace=True, track_best=True)
            eng.refresh_fields()
            done += k
            if last % interval == 0:
                cur = trace[-1, first].cpu().numpy()
                acc = eng.accepted()[first].cpu().numpy()
                temp_hist.append(float(chunk_t[-1]))
                for m_i in range(M):
                    energy_hist[m_i].append(float(cur[m_i]))
                    acc_hist[m_i].append(float(acc[m_i]) / (done * n))
        _, best_s = eng.best()
        best_e = eng.batch_energies(best_s).reshape(M, R)      # exact energies of the best records
        r_best = torch.argmin(best_e, dim=1)
        pick = first + r_best
        best_cfg = as_pm1_float(best_s[pick])                  # [M, n] float32 +-1 on the host
        best_val = best_e[torch.arange(M, device=eng.device), r_best].cpu().numpy()
        final = eng.spins()[first].to(torch.float32).cpu()
        total_time = time.time() - start
        out = []
        for m_i, m in enumerate(models):
            m.spins = final[m_i].clone().to(m.device)
            if hasattr(m, "_invalidate_cache"):
                m._invalidate_cache()
            out.append(AnnealingResult(
                best_configuration=best_cfg[m_i].clone(), best_energy=float(best_val[m_i]),
                energy_history=energy_hist[m_i], temperature_history=list(temp_hist),
                acceptance_rate_history=acc_hist[m_i], total_time=total_time, n_sweeps=done,
                algorithm="simulated_annealing", device=str(eng.device), random_seed=cfg.random_seed))
        self.annealer.total_flips += int(eng.accepted().sum().item())
        self.annealer.total_time += total_time
        return out
