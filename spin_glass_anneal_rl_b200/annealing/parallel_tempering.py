"""ParallelTempering: replica exchange on the B200 sweep engine.

Drop-in for the reference's ``ParallelTempering(ParallelTemperingConfig).run(model,
update_rule) -> AnnealingResult`` (reference annealing/parallel_tempering.py:16-144):

* same config fields and defaults (:16-36) plus ``n_ladders`` (independent ladders run
  side by side; R = n_ladders x n_replicas replicas) and ``device_index``;
* the ladder is generated exactly as in the reference (:146-173); rung 0 is the HOTTEST;
* every outer iteration sweeps all replicas once (:191-203 -> one kernel launch for all
  sweeps up to the next exchange / record point), exchanges are attempted when
  ``sweep % exchange_interval == 0 and sweep > 0`` between adjacent rungs starting at a
  random parity (:113-114, :214-220) with p = min(1, exp((b_j - b_i)(E_j - E_i))) (:244-246);
* configurations stay in place, TEMPERATURES move (the reference swaps the spin tensors,
  :252-258 -- the same Markov chain): ``energy_histories[k]`` is still the energy of
  whatever configuration sits on rung k, and ``energy_history`` in the result is rung 0's
  (the reference returns ``energy_histories[0]``, :134);
* ``anneal`` is an alias of ``run`` so that ``ProblemTemplate.solve_with_annealer(pt)``
  (reference problems/base.py:133) works.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from ..core.spin_dynamics import UpdateRule
from ._backend import as_pm1_float, engine_for, random_spins, rule_name, site_order_for
from .result import AnnealingResult


@dataclass
class ParallelTemperingConfig:
    n_replicas: int = 8
    n_sweeps: int = 1000
    temp_min: float = 0.1
    temp_max: float = 10.0
    temp_distribution: str = "geometric"
    exchange_interval: int = 10
    exchange_method: str = "nearest_neighbor"
    n_threads: Optional[int] = None
    record_interval: int = 10
    random_seed: Optional[int] = None
    # --- additions
    n_ladders: int = 1
    site_order: str = "random"
    device_index: int = 0


class ParallelTempering:
    def __init__(self, config: ParallelTemperingConfig):
        self.config = config
        if config.random_seed is not None:
            torch.manual_seed(config.random_seed)
            np.random.seed(config.random_seed)
        self.temperatures = self._generate_temperature_ladder()
        self.replicas: List = []   # reference attribute (list of models); state lives on the GPU
        self.dynamics: List = []
        self.exchange_attempts = np.zeros((config.n_replicas - 1,))
        self.exchange_accepts = np.zeros((config.n_replicas - 1,))
        self.energy_histories: List[List[float]] = [[] for _ in range(config.n_replicas)]
        self.temp_histories: List[List[float]] = [[] for _ in range(config.n_replicas)]
        self.n_threads = config.n_threads or min(config.n_replicas, 8)
        self.use_cuda = torch.cuda.is_available()
        self.device = torch.device("cuda", config.device_index) if self.use_cuda else torch.device("cpu")
        if self.use_cuda:   # reference parallel_tempering.py:77-80
            from .cuda_kernels import CUDAKernelManager
            self.cuda_kernels = CUDAKernelManager(self.device)
        else:
            self.cuda_kernels = None
        self._final_spins = None

    def _generate_temperature_ladder(self) -> List[float]:
        c = self.config
        if c.temp_distribution == "geometric":
            ratio = c.temp_min / c.temp_max
            return [c.temp_max * (ratio ** (i / (c.n_replicas - 1))) for i in range(c.n_replicas)]
        if c.temp_distribution == "linear":
            return np.linspace(c.temp_max, c.temp_min, c.n_replicas).tolist()
        if c.temp_distribution == "exponential":
            return np.logspace(np.log10(c.temp_max), np.log10(c.temp_min), c.n_replicas).tolist()
        raise ValueError(f"Unknown temperature distribution: {c.temp_distribution}")

    # ------------------------------------------------------------------ run
    def run(self, model, update_rule: UpdateRule = UpdateRule.METROPOLIS) -> AnnealingResult:
        c = self.config
        start = time.time()
        if c.exchange_method not in ("nearest_neighbor", "all_pairs"):
            raise ValueError(f"Unknown exchange method: {c.exchange_method}")
        rule = rule_name(update_rule)
        eng = engine_for(model, c.device_index)
        n, K, L = model.n_spins, c.n_replicas, max(1, int(c.n_ladders))
        R = K * L
        seed = c.random_seed if c.random_seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        host_rng = np.random.RandomState(seed & 0xFFFFFFFF)
        gen = torch.Generator(device=eng.device)
        gen.manual_seed(int(seed) + 104729)

        if eng.n_replicas != R:
            eng.alloc_replicas(R)
        eng.set_spins(random_spins(R, n, eng.device, gen))  # replicas start random (:175-189)
        eng.init_fields()
        eng.set_ladder([max(float(t), 1e-10) for t in self.temperatures])

        self.energy_histories = [[] for _ in range(K)]
        self.temp_histories = [[] for _ in range(K)]
        recorded = []  # device tensors [K] of rung energies (ladder 0), fetched once at the end
        sweep = 0
        xround = 0
        ex_iv, rec_iv = max(1, c.exchange_interval), max(1, c.record_interval)
        while sweep < c.n_sweeps:
            # fuse the sweeps up to the next exchange or record point into one launch
            nxt = sweep
            while True:
                if (nxt % ex_iv == 0 and nxt > 0) or nxt % rec_iv == 0 or nxt == c.n_sweeps - 1:
                    break
                nxt += 1
            k = nxt - sweep + 1
            eng.sweep(k, None, rule=rule, site_order=site_order_for(eng, c.site_order), seed=int(seed),
                      sweep_base=sweep, track_best=True)
            eng.refresh_fields()  # exact fields / energies before they feed an exchange decision
            sweep = nxt
            if sweep % ex_iv == 0 and sweep > 0:
                eng.exchange(int(host_rng.randint(0, 2)), seed=int(seed) ^ 0x5DEECE66D, round=xround)
                xround += 1
            if sweep % rec_iv == 0:
                rep_at = eng.ladder_state()[0][:K].long()
                recorded.append(eng.energies()[rep_at])
            sweep += 1

        rep_at, rep_T, att, acc = eng.ladder_state()
        self.exchange_attempts = att.sum(dim=0).double().cpu().numpy()[:max(K - 1, 0)]
        self.exchange_accepts = acc.sum(dim=0).double().cpu().numpy()[:max(K - 1, 0)]
        if recorded:
            hist = torch.stack(recorded).double().cpu().numpy()  # [n_records, K]
            for r in range(K):
                self.energy_histories[r] = hist[:, r].tolist()
                self.temp_histories[r] = [self.temperatures[r]] * hist.shape[0]
        _, best_s = eng.best()
        best_e = eng.batch_energies(best_s)   # exact energies of the best configurations
        r_best = int(torch.argmin(best_e).item())
        accepted = eng.accepted().double()
        rates = (accepted[rep_at[:K].long()] / float(c.n_sweeps * n)).cpu().tolist()
        self._final_spins = eng.spins()
        self._rung_replica = rep_at.cpu().numpy()
        total_time = time.time() - start
        return AnnealingResult(
            best_configuration=as_pm1_float(best_s[r_best]), best_energy=float(best_e[r_best].item()),
            energy_history=list(self.energy_histories[0]), temperature_history=list(self.temp_histories[0]),
            acceptance_rate_history=rates, total_time=total_time, n_sweeps=c.n_sweeps,
            algorithm="parallel_tempering", device=str(self.device), random_seed=c.random_seed)

    anneal = run  # lets solve_with_annealer(ParallelTempering(...)) work (reference base.py:133)

    # ------------------------------------------------------------------ statistics
    def get_exchange_rates(self) -> np.ndarray:
        rates = np.zeros_like(self.exchange_accepts)
        nz = self.exchange_attempts > 0
        rates[nz] = self.exchange_accepts[nz] / self.exchange_attempts[nz]
        return rates

    def get_statistics(self) -> dict:
        return {
            "temperatures": list(self.temperatures),
            "exchange_rates": self.get_exchange_rates().tolist(),
            "exchange_attempts": self.exchange_attempts.tolist(),
            "exchange_accepts": self.exchange_accepts.tolist(),
            "mean_energies": [float(np.mean(h)) if h else float("nan") for h in self.energy_histories],
        }

    def __repr__(self) -> str:
        return (f"ParallelTempering(n_replicas={self.config.n_replicas}, "
                f"T=[{self.config.temp_min}, {self.config.temp_max}], "
                f"n_sweeps={self.config.n_sweeps})")
