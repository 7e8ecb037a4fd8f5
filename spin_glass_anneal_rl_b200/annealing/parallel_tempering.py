"""ParallelTempering: replica exchange on the B200 sweep engine.

Drop-in for the reference's ``ParallelTempering(ParallelTemperingConfig).run(model,
update_rule) -> AnnealingResult`` (reference annealing/parallel_tempering.py:16-144):

* same config fields and defaults (:16-36) plus ``n_ladders`` (independent ladders run
  side by side; R = n_ladders x n_replicas replicas), ``device_index``, ``rng_mode`` /
  ``replay`` (the reference's recorded random stream, injected) and ``shard`` (ladders that
  span GPUs);
* ``exchange_method``: "nearest_neighbor" (:214-220) and "all_pairs" as the reference's CPU branch
  runs it (:228-232: every pair i < j attempted with probability 0.1, in order);
* ``acceptance_rate_history`` is per TEMPERATURE as in the reference (:137), although the
  configurations do not move: every launch credits a replica's accepted flips to the rung it sat
  on;
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
from ._backend import (as_pm1_float, engine_for, mix_seed, random_spins, require_dense_for_wolff, rule_name,
                       site_order_for)
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
    kernel: str = "auto"
    # "replay": every random number comes from a recorded run of the reference, by TEMPERATURE
    # SLOT (the reference keeps slot k at temperature k and moves configurations):
    #   replay = {"spins0": [K, n], "sites": int [n_sweeps, K, n], "uniforms": f32 [n_sweeps, K, n],
    #             "exchange_draws": one list of numpy draws per exchange round, in the order the
    #             reference made them (nearest_neighbor: the start parity, then one uniform per
    #             pair; all_pairs: selection draw per pair + acceptance draw per selected pair)}
    # One ladder; the best configuration is then taken at the record points over the slots, as
    # the reference does (:117-125), instead of after every sweep.
    rng_mode: str = "philox"
    replay: Optional[dict] = None
    # ladders that span GPUs (one process per GPU): this rank holds the global replicas
    # [shard.start, shard.start + shard.count) of n_ladders x n_replicas; every exchange round
    # all-gathers the energies (multi_gpu.gather_energies) and every rank applies the same
    # decisions.  None = all replicas on this GPU.
    shard: Optional[object] = None


class ParallelTempering:
    def __init__(self, config: ParallelTemperingConfig):
        self.config = config
        if config.random_seed is not None:
            torch.manual_seed(config.random_seed)
            np.random.seed(int(config.random_seed) & 0xFFFFFFFF)
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
        if c.rng_mode not in ("philox", "replay"):
            raise ValueError(f"Unknown rng_mode: {c.rng_mode}")
        rule = rule_name(update_rule)
        eng = engine_for(model, c.device_index)
        if rule == "wolff":
            require_dense_for_wolff(eng)
        n, K, L = model.n_spins, c.n_replicas, max(1, int(c.n_ladders))
        Rg = K * L                               # replicas over all ranks
        sh = c.shard
        lo, R = (int(sh.start), int(sh.count)) if sh is not None else (0, Rg)
        replay = c.replay if c.rng_mode == "replay" else None
        if c.rng_mode == "replay":
            if replay is None or L != 1 or sh is not None:
                raise ValueError("rng_mode='replay' needs the replay streams, one ladder and no sharding")
        seed = c.random_seed if c.random_seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        host_rng = np.random.RandomState(seed & 0xFFFFFFFF)   # same draws on every rank
        key = mix_seed(seed, 0x50540001)
        xkey = mix_seed(seed, 0x50540002)

        if eng.n_replicas != R:
            eng.alloc_replicas(R)
        if replay is not None:
            spins0 = torch.as_tensor(np.asarray(replay["spins0"]), device=eng.device).to(torch.int8)
            r_sites = torch.as_tensor(np.asarray(replay["sites"]), dtype=torch.int32, device=eng.device)
            if rule == "wolff":
                # uniforms[s][k]: the list the cluster updates of temperature slot k consumed in sweep s
                r_uni = [[np.asarray(u, np.float32).reshape(-1) for u in per_slot] for per_slot in replay["uniforms"]]
            else:
                r_uni = torch.as_tensor(np.nan_to_num(np.asarray(replay["uniforms"], np.float32), nan=0.5),
                                        dtype=torch.float32, device=eng.device)
            draws = [list(d) for d in replay["exchange_draws"]]
        else:
            gen = torch.Generator(device=eng.device)
            gen.manual_seed(int(mix_seed(seed, 0x50540003) & 0x7FFFFFFFFFFFFFFF))
            # replicas start random (:175-189); a shard takes its rows of the global draw, so a
            # sharded run starts from the configurations the single-GPU run would
            spins0 = random_spins(Rg, n, eng.device, gen)[lo:lo + R]
        eng.set_spins(spins0.contiguous())
        eng.init_fields()
        eng.set_ladder([max(float(t), 1e-10) for t in self.temperatures], n_global=Rg, replica_offset=lo)

        def all_energies():
            """Energies of every replica by global id (the collective C1 when sharded)."""
            e = eng.energies()
            if sh is None:
                return e
            from .multi_gpu import gather_energies
            return gather_energies(e, Rg)

        def rung_of_local():
            rep_at = eng.ladder_state()[0].long()
            rung = torch.empty(Rg, dtype=torch.long, device=eng.device)
            rung[rep_at] = torch.arange(Rg, device=eng.device) % K
            return rep_at, rung[lo:lo + R]

        self.energy_histories = [[] for _ in range(K)]
        self.temp_histories = [[] for _ in range(K)]
        recorded = []  # device tensors [K] of rung energies (ladder 0), fetched once at the end
        rung_acc = torch.zeros(K, dtype=torch.long, device=eng.device)   # accepted flips per rung
        acc_prev = eng.accepted().clone()
        rp_best_e = torch.full((), float("inf"), dtype=torch.float32, device=eng.device)
        rp_best_s = torch.zeros(n, dtype=torch.int8, device=eng.device)
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
            rep_at, rung_loc = rung_of_local()
            if replay is not None and rule == "wolff":
                # start sites and uniform lists of the slot (= rung) -> the replica sitting on it; a
                # replica's list for the launch is the concatenation over the launch's sweeps
                sl = r_sites[sweep:sweep + k]                       # [k, K, n]
                sites_rep = sl[:, rung_loc, :].permute(1, 0, 2).contiguous()   # [R, k, n]
                rungs = rung_loc.cpu().tolist()
                lists = [np.concatenate([r_uni[s][rg] for s in range(sweep, sweep + k)]) for rg in rungs]
                width = max(1, max(len(u) for u in lists))
                uni_rep = np.full((R, width), 2.0, np.float32)
                for i, u in enumerate(lists):
                    uni_rep[i, :len(u)] = u
                eng.sweep_wolff(k, None, sites=sites_rep, sites_replica_stride=k * n, sites_sweep_stride=n,
                                uniforms=uni_rep, track_best=True)
            elif replay is not None:
                # slot (= rung) streams -> the replica that sits on the rung in this segment
                sl = r_sites[sweep:sweep + k]                       # [k, K, n]
                ul = r_uni[sweep:sweep + k]
                sites_rep = sl[:, rung_loc, :].permute(1, 0, 2).contiguous()   # [R, k, n]
                uni_rep = ul[:, rung_loc, :].permute(1, 0, 2).contiguous()
                eng.sweep(k, None, rule=rule, sites=sites_rep, sites_block_stride=k * n,
                          sites_sweep_stride=n, uniforms=uni_rep, replicas_per_block=1,
                          track_best=True, kernel="simt" if c.kernel == "auto" else c.kernel)
            else:
                eng.sweep(k, None, rule=rule, site_order=site_order_for(eng, c.site_order), seed=key,
                          sweep_base=sweep, track_best=True, kernel=c.kernel, replica_base=lo)
            eng.refresh_fields()  # exact fields / energies before they feed an exchange decision
            acc_now = eng.accepted()
            rung_acc.index_add_(0, rung_loc, (acc_now - acc_prev))
            acc_prev = acc_now.clone()
            sweep = nxt
            if sweep % ex_iv == 0 and sweep > 0:
                e_all = all_energies() if sh is not None else None
                if replay is not None:
                    d = draws[xround]
                    if c.exchange_method == "nearest_neighbor":
                        eng.exchange(int(d[0]), uniforms=np.asarray(d[1:] + [0.0] * (K // 2), np.float64)[:max(K // 2, 1)])
                    else:
                        u = np.zeros(K * (K - 1), np.float64)
                        u[:len(d)] = d
                        eng.exchange(0, uniforms=u, method="all_pairs")
                elif c.exchange_method == "nearest_neighbor":
                    eng.exchange(int(host_rng.randint(0, 2)), seed=xkey, round=xround, energies_all=e_all)
                else:
                    eng.exchange(0, seed=xkey, round=xround, method="all_pairs", energies_all=e_all)
                xround += 1
            if sweep % rec_iv == 0:
                rep_at = eng.ladder_state()[0][:K].long()
                e_now = all_energies()
                rung_e = e_now[rep_at]
                recorded.append(rung_e)
                if replay is not None:      # the reference's best tracking (:121-125)
                    idx = torch.argmin(rung_e)          # first minimum, like min() over the slots
                    better = rung_e[idx] < rp_best_e
                    rp_best_e = torch.where(better, rung_e[idx], rp_best_e)
                    rp_best_s = torch.where(better, eng.spins()[rep_at[idx]], rp_best_s)
            sweep += 1

        rep_at, rep_T, att, acc = eng.ladder_state()
        self.exchange_attempts = att.sum(dim=0).double().cpu().numpy()[:max(K - 1, 0)]
        self.exchange_accepts = acc.sum(dim=0).double().cpu().numpy()[:max(K - 1, 0)]
        if recorded:
            hist = torch.stack(recorded).double().cpu().numpy()  # [n_records, K]
            for r in range(K):
                self.energy_histories[r] = hist[:, r].tolist()
                self.temp_histories[r] = [self.temperatures[r]] * hist.shape[0]
        if sh is not None:
            from .multi_gpu import sum_over_ranks
            rung_acc = sum_over_ranks(rung_acc)
        # acceptance rate per TEMPERATURE (the reference's dynamics[k] stays with slot k, :137)
        rates = (rung_acc.double() / float(c.n_sweeps * n * L)).cpu().tolist()
        if rule == "wolff":   # cluster sites all count as accepted, nothing as rejected (:254)
            rates = [1.0 if r > 0 else 0.0 for r in rates]
        self._final_spins = eng.spins()
        self._rung_replica = rep_at.cpu().numpy()
        if replay is not None:
            best_cfg, best_val = rp_best_s, float(rp_best_e.item())
        else:
            _, best_s = eng.best()
            best_e = eng.batch_energies(best_s)   # exact energies of the best configurations
            r_best = int(torch.argmin(best_e).item())
            best_cfg, best_val = best_s[r_best], float(best_e[r_best].item())
            if sh is not None:
                from .multi_gpu import global_argmin
                best_val, best_cfg, _ = global_argmin(best_e, best_s, sh)
        total_time = time.time() - start
        return AnnealingResult(
            best_configuration=as_pm1_float(best_cfg), best_energy=best_val,
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
