"""GPUAnnealer: simulated annealing on the B200 sweep engine.

Drop-in for the reference's ``GPUAnnealer(GPUAnnealerConfig).anneal(model, update_rule)
-> AnnealingResult`` (reference annealing/gpu_annealer.py:30-183):

* the config keeps every reference field with the same defaults (:30-59) and appends
  ``n_replicas``, ``site_order``, ``replicas_per_block``, ``device_index``, ``rng_mode`` /
  ``replay`` (injected reference stream), ``kernel``, ``coupling_dtype``;
* the loop keeps the reference's bookkeeping: temperature from the schedule per sweep
  (:141-142, clamp of SpinDynamics.set_temperature), best energy / configuration compared
  after EVERY sweep (:151-153, done inside the kernel), histories appended when
  ``sweep % record_interval == 0`` (:156-159), early stop through the same
  relative-std test (:254-269), ``model.spins`` left at the final configuration, the best
  configuration returned as a float32 +-1 CPU tensor (:171);
* with ``n_replicas > 1`` all replicas follow the same schedule from independent random
  starts (replica 0 starts from ``model.spins``); histories follow replica 0, the best
  energy / configuration are taken over all replicas.

The per-attempt work -- SpinDynamics.sweep() -- runs in the CUDA sweep kernel, with the
sweeps between two record points fused into one launch.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from ..core.spin_dynamics import UpdateRule
from ._backend import (as_pm1_float, engine_for, mix_seed, random_spins, require_dense_for_wolff, rule_name,
                       site_order_for)
from .result import AnnealingResult
from .temperature_scheduler import ScheduleType, TemperatureScheduler


@dataclass
class GPUAnnealerConfig:
    n_sweeps: int = 1000
    initial_temp: float = 10.0
    final_temp: float = 0.01
    schedule_type: ScheduleType = ScheduleType.GEOMETRIC
    schedule_params: Dict = None
    block_size: int = 256
    shared_memory_size: int = 48 * 1024
    record_interval: int = 10
    energy_tolerance: float = 1e-8
    random_seed: Optional[int] = None
    enable_adaptive_optimization: bool = True
    enable_caching: bool = True
    enable_performance_profiling: bool = True
    adaptive_config: Optional[object] = None
    compute_config: Optional[object] = None
    # --- additions (defaults keep the reference behaviour)
    n_replicas: int = 1
    site_order: str = "random"       # "random" | "sequential" | "random_per_block"
    replicas_per_block: int = 0      # 0 = let the engine choose
    device_index: int = 0
    # "philox": in-kernel counter RNG (production).  "replay": the (site, uniform) of every attempt
    # is injected from a recorded stream of the reference -- ``replay`` = {"sites": int [n_sweeps,
    # n], "uniforms": float32 [n_sweeps, n] (or [n_replicas, n_sweeps, n])}; anneal() then follows
    # the reference's trajectory (bit for bit on integer couplings).  For UpdateRule.WOLFF "sites"
    # are the start sites of the n cluster updates of every sweep and "uniforms" is the flat list
    # of the uniforms the run consumes, in order ([m], or [n_replicas, m])
    rng_mode: str = "philox"
    replay: Optional[Dict] = None
    kernel: str = "auto"             # "auto" | "simt" | "tc" | "small" (SG_KERNEL_*)
    # how the dense couplings are held for the tensor-core sweep: "auto"/"fp32" = three bf16
    # planes (every fp32 coupling exactly), "bf16" = one plane, "int8" = one plane, integer
    # couplings required (exact)
    coupling_dtype: str = "auto"

    def __post_init__(self):
        if self.schedule_params is None:
            self.schedule_params = {"alpha": 0.95}


class GPUAnnealer:
    def __init__(self, config: GPUAnnealerConfig):
        self.config = config
        if config.random_seed is not None:
            torch.manual_seed(config.random_seed)
            np.random.seed(int(config.random_seed) & 0xFFFFFFFF)
        self.use_cuda = torch.cuda.is_available()
        self.device = torch.device("cuda", config.device_index) if self.use_cuda else torch.device("cpu")
        if self.use_cuda:
            print(f"Using GPU: {torch.cuda.get_device_name(config.device_index)}")
        else:
            print("CUDA not available: anneal() needs a B200 (there is no CPU path)")
        # reference attributes (gpu_annealer.py:84-90): the per-call operator interface, kept for
        # callers that use it directly; anneal() itself runs the batched launches below
        if self.use_cuda:
            from .cuda_kernels import CUDAKernelManager, GPUMemoryOptimizer
            self.cuda_kernels = CUDAKernelManager(self.device)
            self.memory_optimizer = GPUMemoryOptimizer(self.device)
        else:
            self.cuda_kernels = None
            self.memory_optimizer = None
        self.total_flips = 0
        self.total_time = 0.0
        self._launch_seed = 0

    # ------------------------------------------------------------------ anneal
    def anneal(self, model, update_rule: UpdateRule = UpdateRule.METROPOLIS) -> AnnealingResult:
        cfg = self.config
        start = time.time()
        rule = rule_name(update_rule)
        eng = engine_for(model, cfg.device_index)
        wolff = rule == "wolff"
        if wolff:
            require_dense_for_wolff(eng)
        n, R = model.n_spins, max(1, int(cfg.n_replicas))
        seed = cfg.random_seed if cfg.random_seed is not None else int(torch.initial_seed() & 0x7FFFFFFF)
        self._launch_seed += 1
        philox_seed = mix_seed(seed, self._launch_seed)

        if cfg.rng_mode not in ("philox", "replay"):
            raise ValueError(f"Unknown rng_mode: {cfg.rng_mode}")
        planes = {"auto": 0, "fp32": 3, "bf16": 1, "int8": 1}.get(cfg.coupling_dtype)
        if planes is None:
            raise ValueError(f"Unknown coupling_dtype: {cfg.coupling_dtype}")
        if cfg.coupling_dtype == "int8":
            Jm = model.couplings.to_dense() if model.couplings.is_sparse else model.couplings
            if not bool(torch.all(Jm == Jm.round())) or float(Jm.abs().max()) > 256:
                raise ValueError("coupling_dtype='int8' needs integer couplings with |J| <= 256")
        replay_sites = replay_uni = None
        if cfg.rng_mode == "replay":
            if not cfg.replay or "sites" not in cfg.replay or "uniforms" not in cfg.replay:
                raise ValueError("rng_mode='replay' needs replay={'sites': ..., 'uniforms': ...}")
            replay_sites = torch.as_tensor(np.asarray(cfg.replay["sites"]), dtype=torch.int32,
                                           device=eng.device).reshape(-1, n)
        wolff_cursor = None
        if cfg.rng_mode == "replay" and wolff:
            replay_uni = torch.as_tensor(np.asarray(cfg.replay["uniforms"], np.float32), dtype=torch.float32,
                                         device=eng.device)
            replay_uni = replay_uni.reshape(1, -1) if replay_uni.dim() == 1 else replay_uni
            if replay_uni.shape[0] != R:
                # (a cluster run consumes as many uniforms as ITS clusters ask for: replicas that start
                # elsewhere cannot share the recorded list of one run)
                raise ValueError("replay stream: UpdateRule.WOLFF needs one uniform list per replica, "
                                 "uniforms [m] for n_replicas = 1 or [n_replicas, m]")
            replay_uni = replay_uni.contiguous()
            wolff_cursor = torch.zeros(R, dtype=torch.int64, device=eng.device)
        elif cfg.rng_mode == "replay":
            replay_uni = torch.as_tensor(np.nan_to_num(np.asarray(cfg.replay["uniforms"], np.float32), nan=0.5),
                                         dtype=torch.float32, device=eng.device)
            replay_uni = replay_uni.reshape(-1, replay_sites.shape[0], n)
            if replay_uni.shape[0] == 1 and R > 1:
                replay_uni = replay_uni.expand(R, -1, -1)
            if replay_uni.shape[0] != R:   # (a recorded run that stopped early holds fewer sweeps)
                raise ValueError("replay stream: need sites [n_sweeps, n] and uniforms [(R,) n_sweeps, n]")

        def launch(k, temps, first):
            """k sweeps starting at absolute sweep `first`."""
            if wolff:   # cluster updates: exact energies after every sweep, no kernel options
                common = dict(temps_sweep_stride=1, sweep_base=first, energy_trace=True, track_best=True)
                if replay_sites is not None:
                    return eng.sweep_wolff(k, temps, sites=replay_sites[first:first + k].contiguous(),
                                           uniforms=replay_uni, cursor=wolff_cursor, **common)
                return eng.sweep_wolff(k, temps, site_order=cfg.site_order, seed=philox_seed, **common)
            common = dict(temps_sweep_stride=1, rule=rule, sweep_base=first, energy_trace=True,
                          track_best=True, kernel=cfg.kernel, coupling_planes=planes)
            if replay_sites is not None:
                return eng.sweep(k, temps, sites=replay_sites[first:first + k].contiguous(),
                                 uniforms=replay_uni[:, first:first + k, :].contiguous(), **common)
            return eng.sweep(k, temps, site_order=site_order_for(eng, cfg.site_order), seed=philox_seed,
                             replicas_per_block=cfg.replicas_per_block, **common)

        gen = torch.Generator(device=eng.device)
        gen.manual_seed(int(seed) + 7919 * self._launch_seed)
        spins0 = random_spins(R, n, eng.device, gen)
        spins0[0] = model.spins.to(eng.device).sign().to(torch.int8)
        if eng.n_replicas != R:
            eng.alloc_replicas(R)
        eng.set_spins(spins0)
        eng.init_fields()  # local fields, energies, best := initial (reference :130-131)
        acc_base = eng.accepted()[0].item()

        schedule = TemperatureScheduler.create_schedule(
            cfg.schedule_type, cfg.initial_temp, cfg.final_temp, cfg.n_sweeps, **cfg.schedule_params)
        if cfg.n_sweeps <= 0:
            raise ValueError("n_sweeps must be positive")

        e0 = float(eng.energies()[0].item())
        energy_history: List[float] = [e0]
        temperature_history: List[float] = [cfg.initial_temp]
        acceptance_rate_history: List[float] = [0.0]
        interval = max(1, int(cfg.record_interval))
        temps_all = None
        if schedule.stateless:
            temps_all = np.maximum(schedule.precompute(cfg.n_sweeps), 1e-10)
            schedule.temperature_history.extend(temps_all.tolist())

        def acceptance_rate(done):
            """cumulative n_accepted / (n_accepted + n_rejected) of replica 0, as the reference counts"""
            acc = eng.accepted()[0].item() - acc_base
            if wolff:   # every cluster site counts as accepted, nothing as rejected (:254)
                return 1.0 if acc > 0 else 0.0
            return acc / (done * n) if done else 0.0

        sweep = 0
        done = 0  # sweeps executed
        stopped = False
        if temps_all is None and not wolff and not os.environ.get("SG_ADAPTIVE_HOST"):
            # ADAPTIVE on the device: the geometric base schedule is precomputed, the feedback
            # step (running acceptance rate of replica 0 -> temperature of the next sweep) is a
            # one-thread kernel between two sweep launches, so the host only synchronises at the
            # record points, like for every other schedule
            sc = schedule.config
            base = torch.tensor([float(schedule.base_schedule.get_temperature(s)) for s in range(cfg.n_sweeps)],
                                dtype=torch.float64, device=eng.device)
            temps_dev = torch.zeros(cfg.n_sweeps, dtype=torch.float64, device=eng.device)
            state = torch.zeros(int(sc.adaptation_window) + 1, dtype=torch.float64, device=eng.device)
            while done < cfg.n_sweeps and not stopped:
                eng.adaptive_temperature(done, base, state, temps_dev, accepted_base=int(acc_base),
                                         window=int(sc.adaptation_window),
                                         target_acceptance=float(sc.target_acceptance),
                                         adaptation_rate=float(sc.adaptation_rate),
                                         final_temp=float(sc.final_temp))
                trace = launch(1, temps_dev[done:done + 1], done)
                eng.refresh_fields()
                sweep = done
                done += 1
                if sweep % interval == 0:
                    cur_e = float(eng.energies()[0].item())   # exact: fields were just refreshed
                    acc = eng.accepted()[0].item() - acc_base
                    energy_history.append(cur_e)
                    temperature_history.append(float(temps_dev[sweep].item()))
                    acceptance_rate_history.append(acc / (done * n))
                    if self._check_convergence(energy_history):
                        print(f"Converged at sweep {sweep}")
                        stopped = True
            schedule.temperature_history.extend(temps_dev[:done].cpu().tolist())
            schedule.current_temp = schedule.temperature_history[-1]
        while done < cfg.n_sweeps and not stopped:
            # run up to and including the next record sweep (sweep % interval == 0)
            if temps_all is not None:
                last = done if done % interval == 0 else min(cfg.n_sweeps - 1, (done // interval + 1) * interval)
                chunk_t = temps_all[done:last + 1]
            else:  # ADAPTIVE: one sweep at a time, fed with the cumulative acceptance rate
                chunk_t = np.array([max(schedule.update(done, acceptance_rate=acceptance_rate(done)), 1e-10)])
                last = done
            k = len(chunk_t)
            trace = launch(k, chunk_t, done)
            # incremental field updates drift for non-integer couplings; the reference recomputes
            # every field from scratch, so refresh them exactly between launches (one K2 pass)
            eng.refresh_fields()
            done += k
            sweep = last
            if sweep % interval == 0:
                cur_e = float(eng.energies()[0].item())   # exact: fields were just refreshed
                energy_history.append(cur_e)
                temperature_history.append(float(chunk_t[-1]))
                acceptance_rate_history.append(acceptance_rate(done))
                if self._check_convergence(energy_history):
                    print(f"Converged at sweep {sweep}")
                    stopped = True

        _, best_s = eng.best()
        best_e = eng.batch_energies(best_s)   # exact energies of the best configurations
        r_best = int(torch.argmin(best_e).item())
        final_spins = eng.spins()
        model.spins = final_spins[0].to(torch.float32).to(model.device)
        model._invalidate_cache() if hasattr(model, "_invalidate_cache") else None
        total_time = time.time() - start
        self.total_flips += int((eng.accepted().sum().item()))
        self.total_time += total_time
        return AnnealingResult(
            best_configuration=as_pm1_float(best_s[r_best]), best_energy=float(best_e[r_best].item()),
            energy_history=energy_history, temperature_history=temperature_history,
            acceptance_rate_history=acceptance_rate_history, total_time=total_time, n_sweeps=done,
            algorithm="simulated_annealing", device=str(self.device), random_seed=cfg.random_seed)

    # ------------------------------------------------------------------ helpers kept from the reference API
    def _check_convergence(self, energy_history: List[float]) -> bool:
        if len(energy_history) < 50:
            return False
        recent = energy_history[-20:]
        std, mean = np.std(recent), np.mean(recent)
        if abs(mean) > 0:
            return (std / abs(mean)) < self.config.energy_tolerance
        return std < self.config.energy_tolerance

    def _move_model_to_gpu(self, model):
        """The reference copied the model to the device per call (:185-197); here the engine
        keeps J and h resident, so this only makes sure they are uploaded."""
        engine_for(model, self.config.device_index)
        return model

    def benchmark(self, model_sizes: List[int], n_trials: int = 3) -> Dict:
        from ..core.ising_model import IsingModel, IsingModelConfig
        results = {}
        for size in model_sizes:
            times, energies, sps = [], [], []
            for _ in range(n_trials):
                model = IsingModel(IsingModelConfig(n_spins=size, use_sparse=False))
                a = torch.rand(size, size) * 2 - 1
                mask = (torch.rand(size, size) < min(1.0, 2.0 / size)).float()
                J = torch.triu(a * mask, 1)
                model.set_couplings_from_matrix(J + J.T)
                res = self.anneal(model)
                times.append(res.total_time)
                energies.append(res.best_energy)
                sps.append(res.n_sweeps / max(res.total_time, 1e-9))
            results[size] = {"mean_time": float(np.mean(times)), "std_time": float(np.std(times)),
                             "mean_energy": float(np.mean(energies)), "std_energy": float(np.std(energies)),
                             "mean_sps": float(np.mean(sps)), "std_sps": float(np.std(sps))}
        return results

    def get_memory_usage(self) -> Dict:
        if not self.use_cuda:
            return {"device": "cpu", "memory_allocated": 0, "memory_reserved": 0}
        alloc, res = torch.cuda.memory_allocated(self.device), torch.cuda.memory_reserved(self.device)
        return {"device": str(self.device), "memory_allocated": alloc, "memory_reserved": res,
                "memory_allocated_mb": alloc / 1024 ** 2, "memory_reserved_mb": res / 1024 ** 2}

    def __repr__(self) -> str:
        return (f"GPUAnnealer(device={self.device}, n_sweeps={self.config.n_sweeps}, "
                f"schedule={self.config.schedule_type.value})")
