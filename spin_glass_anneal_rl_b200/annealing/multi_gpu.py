"""Replica sharding across the GPUs of one node (one process per GPU, torch.distributed).

Replaces the reference's ``annealing/multi_gpu.py`` (thread pool over Python objects,
:110-307; its "communication_backend" string is never used, :26,41).  Replicas are
independent between exchanges and an exchange touches only (E, beta) scalars of one
ladder, so whole ladders are placed on one GPU where possible and the sweep path needs NO
data-path collective.  Collectives: C2, the final argmin -- an all_gather of
(best energy, rank, local replica) followed by a broadcast of the winning configuration
from its owner (N bytes) -- and C1, for ladders that span GPUs
(``shard_replicas_split``): one all-gather of the per-replica energies per exchange round
(``gather_energies``), after which every rank applies the same decisions from the same
counter RNG (``sg_exchange`` with ``energies_all``; temperatures move, configurations stay;
replaces ``anneal_replica_exchange``, reference annealing/multi_gpu.py:234-307).  Works with the nccl backend on GPUs and with gloo on CPU
tensors (used by the CPU tests of the host logic).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class ReplicaShard:
    rank: int
    world: int
    start: int      # first global replica of this rank
    count: int      # replicas held by this rank
    n_rungs: int    # ladder length (1 = plain multi-start annealing)

    @property
    def n_ladders(self) -> int:
        return self.count // self.n_rungs


def shard_replicas(n_replicas: int, world: int, rank: int, n_rungs: int = 1) -> ReplicaShard:
    """Contiguous block of replicas for ``rank``; ladders are never split across ranks.

    ``n_replicas`` must be a multiple of ``n_rungs``; ladders are dealt as evenly as
    possible (the first ``n_ladders % world`` ranks get one more)."""
    if n_rungs < 1 or n_replicas % n_rungs != 0:
        raise ValueError("n_replicas must be a multiple of the ladder length")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    ladders = n_replicas // n_rungs
    base, extra = divmod(ladders, world)
    mine = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return ReplicaShard(rank, world, first * n_rungs, mine * n_rungs, n_rungs)


def global_argmin(best_energy: torch.Tensor, best_spins: torch.Tensor,
                  shard: Optional[ReplicaShard] = None,
                  group=None) -> Tuple[float, torch.Tensor, int]:
    """Best configuration over all ranks.

    best_energy [R_local] and best_spins [R_local, n] live on this rank's device (or on the
    CPU under gloo).  Returns (energy, configuration [n] on the caller's device, global
    replica id); every rank gets the same answer.  Without an initialised process group this
    is the local argmin."""
    local_idx = int(torch.argmin(best_energy).item()) if best_energy.numel() else -1
    local_e = float(best_energy[local_idx].item()) if local_idx >= 0 else float("inf")
    if not (dist.is_available() and dist.is_initialized()):
        gid = (shard.start if shard else 0) + local_idx
        return local_e, best_spins[local_idx].clone(), gid
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = best_energy.device
    mine = torch.tensor([local_e, float(rank), float(local_idx),
                         float((shard.start if shard else 0) + local_idx)],
                        dtype=torch.float64, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    table = torch.stack(gathered)                     # [world, 4]
    # ties are broken by the lowest rank so that every rank picks the same owner
    owner = int(torch.argmin(table[:, 0]).item())
    energy = float(table[owner, 0].item())
    gid = int(table[owner, 3].item())
    n = best_spins.shape[1]
    winner = best_spins[local_idx].clone() if rank == owner else \
        torch.empty(n, dtype=best_spins.dtype, device=dev)
    dist.broadcast(winner, src=dist.get_global_rank(group, owner) if group is not None else owner,
                   group=group)
    return energy, winner, gid


def rank_seed(seed: Optional[int], rank: int) -> int:
    """Independent Philox keys per rank (replica ids are local to a rank).  With ``seed=None`` the
    base is torch's initial seed -- which the usual ``torch.manual_seed(k)`` on every rank makes
    identical across ranks -- so the rank is always mixed in (splitmix64)."""
    from ._backend import mix_seed
    base = int(torch.initial_seed()) if seed is None else int(seed)
    return int(mix_seed(base, 0x52414E4B, rank) & 0x7FFFFFFFFFFFFFFF)


def shard_replicas_split(n_replicas: int, world: int, rank: int, n_rungs: int = 1) -> ReplicaShard:
    """Equal contiguous blocks that MAY cut a ladder (fewer ladders than GPUs, or one long
    ladder): exchanges then need the all-gathered energy table (``gather_energies``)."""
    if n_rungs < 1 or n_replicas % n_rungs != 0:
        raise ValueError("n_replicas must be a multiple of the ladder length")
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if n_replicas % world != 0:
        raise ValueError("split ladders need n_replicas divisible by the number of ranks")
    per = n_replicas // world
    return ReplicaShard(rank, world, rank * per, per, n_rungs)


def gather_energies(local_energy: torch.Tensor, n_global: int, group=None) -> torch.Tensor:
    """Collective C1: all-gather of the per-replica energies (4 B per replica; the rung of every
    replica is already known to every rank, since all ranks apply the same exchange decisions).
    local_energy [R_local] float32 on this rank's device -> [n_global] by global replica id.
    NCCL on GPUs (NVLink), gloo on CPU tensors.  Without a process group: the input itself."""
    if not (dist.is_available() and dist.is_initialized()):
        if local_energy.numel() != n_global:
            raise ValueError("no process group: the local energies must be the whole table")
        return local_energy
    world = dist.get_world_size(group)
    if local_energy.numel() * world != n_global:
        raise ValueError("gather_energies needs equal shards (n_global = world x R_local)")
    out = torch.empty(n_global, dtype=local_energy.dtype, device=local_energy.device)
    dist.all_gather_into_tensor(out, local_energy.contiguous(), group=group)
    return out


def sum_over_ranks(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized():
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


@dataclass
class MultiGPUConfig:
    """Mirror of the reference's MultiGPUConfig (annealing/multi_gpu.py:20-43) for the fields that
    mean something here.  ``strategy`` "data_parallel" = independent multi-start replicas,
    "replica_exchange" = whole temperature ladders per GPU; ``communication_backend`` is what
    torch.distributed was initialised with by the launcher (torchrun), "nccl" on GPUs."""
    n_replicas: int = 8192            # over all GPUs
    strategy: str = "data_parallel"
    communication_backend: str = "nccl"
    n_rungs: int = 64                 # ladder length for replica_exchange

    def __post_init__(self):
        if self.strategy not in ("data_parallel", "replica_exchange"):
            raise ValueError(f"Unknown strategy: {self.strategy} (model_parallel would split the "
                             "couplings across GPUs, which changes the model; it is not offered)")
        if self.communication_backend not in ("nccl", "gloo", "mpi"):
            raise ValueError(f"Unknown communication backend: {self.communication_backend}")


class MultiGPUAnnealer:
    """One process per GPU (torchrun): every rank anneals its shard of the replicas on its own
    B200 with no data-path collective; the best configuration over all ranks is found with one
    all_gather + one broadcast at the end.  Without an initialised process group it is a
    single-GPU annealer.  ``anneal(model)`` returns the same AnnealingResult on every rank."""

    def __init__(self, config: MultiGPUConfig, annealer_config=None):
        self.config = config
        self.annealer_config = annealer_config
        init = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size() if init else 1
        self.rank = dist.get_rank() if init else 0

    def shard(self) -> ReplicaShard:
        rungs = self.config.n_rungs if self.config.strategy == "replica_exchange" else 1
        if rungs > 1 and self.world > 1 and (self.config.n_replicas // rungs) % self.world != 0:
            return shard_replicas_split(self.config.n_replicas, self.world, self.rank, rungs)
        return shard_replicas(self.config.n_replicas, self.world, self.rank, rungs)

    # the reference's per-strategy entry points (annealing/multi_gpu.py:121-307) on this design
    def anneal_data_parallel(self, model, update_rule=None):
        """Independent multi-start replicas on every GPU, best of all ranks (reference :121-176 runs
        one full anneal per GPU in a thread)."""
        return self._with_strategy("data_parallel", model, update_rule)

    def anneal_replica_exchange(self, model, update_rule=None):
        """Parallel tempering over all GPUs: whole ladders per GPU, or ladders that span GPUs with the
        energies all-gathered every exchange round (reference :234-307 holds one temperature per GPU)."""
        return self._with_strategy("replica_exchange", model, update_rule)

    def anneal_model_parallel(self, model, update_rule=None):
        """Not offered: the reference (:178-232) cuts the couplings into block-diagonal pieces, which
        anneals a different model."""
        raise NotImplementedError("model_parallel would split the couplings across GPUs, which changes "
                                  "the model; use data_parallel or replica_exchange")

    def _with_strategy(self, strategy, model, update_rule):
        import copy
        saved = self.config
        self.config = copy.copy(saved)
        self.config.strategy = strategy
        try:
            return self.anneal(model, update_rule)
        finally:
            self.config = saved

    def cleanup(self) -> None:
        """Nothing to release: the process group belongs to the launcher (torchrun)."""

    def anneal(self, model, update_rule=None):
        import copy
        from ..core.spin_dynamics import UpdateRule
        from .gpu_annealer import GPUAnnealer, GPUAnnealerConfig
        from .parallel_tempering import ParallelTempering, ParallelTemperingConfig
        from .result import AnnealingResult
        rule = update_rule or UpdateRule.METROPOLIS
        sh = self.shard()
        local_device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if self.config.strategy == "replica_exchange":
            cfg = copy.copy(self.annealer_config) if self.annealer_config else ParallelTemperingConfig()
            ladders = self.config.n_replicas // self.config.n_rungs
            cfg.device_index = local_device
            if self.world > 1 and ladders % self.world != 0:
                # ladders span GPUs: one global replica set, same seed everywhere, energies
                # all-gathered at every exchange round (C1); the result is already global
                cfg.n_replicas, cfg.n_ladders = self.config.n_rungs, ladders
                cfg.shard = shard_replicas_split(self.config.n_replicas, self.world, self.rank,
                                                 self.config.n_rungs)
                if cfg.random_seed is None:
                    t = torch.tensor([torch.initial_seed() & 0x7FFFFFFF], dtype=torch.int64,
                                     device=torch.device("cuda", local_device)
                                     if torch.cuda.is_available() else "cpu")
                    dist.broadcast(t, src=0)
                    cfg.random_seed = int(t.item())
                return ParallelTempering(cfg).run(model, rule)
            cfg.n_replicas, cfg.n_ladders = self.config.n_rungs, max(1, sh.n_ladders)
            cfg.random_seed = rank_seed(cfg.random_seed, self.rank)
            res = ParallelTempering(cfg).run(model, rule)
        else:
            cfg = copy.copy(self.annealer_config) if self.annealer_config else GPUAnnealerConfig()
            cfg.n_replicas = max(1, sh.count)
            cfg.random_seed = rank_seed(cfg.random_seed, self.rank)
            cfg.device_index = local_device
            res = GPUAnnealer(cfg).anneal(model, rule)
        dev = torch.device("cuda", local_device) if torch.cuda.is_available() else torch.device("cpu")
        e, s, gid = global_argmin(torch.tensor([res.best_energy], dtype=torch.float64, device=dev),
                                  res.best_configuration.to(dev).reshape(1, -1), sh)
        return AnnealingResult(best_configuration=s.cpu(), best_energy=float(e),
                               energy_history=res.energy_history,
                               temperature_history=res.temperature_history,
                               acceptance_rate_history=res.acceptance_rate_history,
                               total_time=res.total_time, n_sweeps=res.n_sweeps,
                               algorithm=res.algorithm + f"+{self.world}gpu", device=str(dev),
                               random_seed=res.random_seed)
