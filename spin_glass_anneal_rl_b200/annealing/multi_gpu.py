"""Replica sharding across the GPUs of one node (one process per GPU, torch.distributed).

Replaces the reference's ``annealing/multi_gpu.py`` (thread pool over Python objects,
:110-307; its "communication_backend" string is never used, :26,41).  Replicas are
independent between exchanges and an exchange touches only (E, beta) scalars of one
ladder, so whole ladders are placed on one GPU and the sweep path needs NO data-path
collective.  The only collective is the final argmin: an all_gather of
(best energy, rank, local replica) followed by a broadcast of the winning configuration
from its owner (N bytes).  Works with the nccl backend on GPUs and with gloo on CPU
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


def rank_seed(seed: Optional[int], rank: int) -> Optional[int]:
    """Independent Philox keys per rank (replica ids are local to a rank)."""
    return None if seed is None else int(seed) * 1_000_003 + rank
