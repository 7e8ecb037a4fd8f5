"""Host logic of the multi-GPU path on CPU: world_size-2 gloo group (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spin_glass_anneal_rl_b200.annealing.multi_gpu import (global_argmin, rank_seed,
                                                           shard_replicas)


def test_shards_are_disjoint_and_keep_ladders_whole():
    for n_rep, world, rungs in [(8192, 8, 64), (8192, 3, 64), (10, 4, 1), (64, 8, 64), (48, 5, 8)]:
        seen = []
        for r in range(world):
            sh = shard_replicas(n_rep, world, r, rungs)
            assert sh.count % rungs == 0 and sh.start % rungs == 0
            seen += list(range(sh.start, sh.start + sh.count))
        assert seen == list(range(n_rep))
        counts = [shard_replicas(n_rep, world, r, rungs).n_ladders for r in range(world)]
        assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_replicas(10, 2, 0, 4)
    with pytest.raises(ValueError):
        shard_replicas(8, 2, 2, 1)
    assert rank_seed(None, 0) != rank_seed(None, 1) and rank_seed(7, 0) != rank_seed(7, 1)
    assert rank_seed(2 ** 62, 0) != rank_seed(2 ** 62 + 2 ** 44, 0)   # large seeds do not alias


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, total = 12, 10
        sh = shard_replicas(total, world, rank, 1)
        g = torch.Generator().manual_seed(5)
        all_e = torch.randn(total, generator=g)
        all_s = (torch.randint(0, 2, (total, n), generator=g) * 2 - 1).to(torch.int8)
        e, cfg, gid = global_argmin(all_e[sh.start:sh.start + sh.count].clone(),
                                    all_s[sh.start:sh.start + sh.count].clone(), sh)
        want = int(torch.argmin(all_e))
        ok = (gid == want) and abs(e - float(all_e[want])) < 1e-12 and torch.equal(cfg, all_s[want])
        # a tie between ranks must resolve to the same owner everywhere
        tie_e = torch.tensor([1.0, -3.0]) if rank == 0 else torch.tensor([-3.0, 0.0])
        tie_s = torch.full((2, n), rank, dtype=torch.int8)
        e2, cfg2, _ = global_argmin(tie_e, tie_s, shard_replicas(4, world, rank, 1))
        ok = ok and e2 == -3.0 and int(cfg2[0]) == 0
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_global_argmin_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_global_argmin_without_process_group():
    e = torch.tensor([3.0, -1.0, 2.0])
    s = torch.arange(12, dtype=torch.int8).reshape(3, 4)
    be, cfg, gid = global_argmin(e, s, shard_replicas(6, 2, 1, 1))
    assert be == -1.0 and torch.equal(cfg, s[1]) and gid == 3 + 1


def test_multi_gpu_config_validation():
    from spin_glass_anneal_rl_b200.annealing.multi_gpu import MultiGPUAnnealer, MultiGPUConfig
    with pytest.raises(ValueError):
        MultiGPUConfig(strategy="model_parallel")
    with pytest.raises(ValueError):
        MultiGPUConfig(communication_backend="smoke-signals")
    a = MultiGPUAnnealer(MultiGPUConfig(n_replicas=128, strategy="replica_exchange", n_rungs=16))
    sh = a.shard()
    assert (a.world, a.rank, sh.count, sh.n_ladders) == (1, 0, 128, 8)
    # the reference's per-strategy entry points run anneal() with that strategy and restore the config
    seen = []
    a.anneal = lambda model, rule=None: seen.append(a.config.strategy) or "result"
    assert a.anneal_data_parallel(object()) == "result" and a.anneal_replica_exchange(object()) == "result"
    assert seen == ["data_parallel", "replica_exchange"] and a.config.strategy == "replica_exchange"
    with pytest.raises(NotImplementedError):
        a.anneal_model_parallel(object())
    a.cleanup()


def _worker_c1(rank, world, port, out):
    """Collective C1 on CPU tensors: every rank ends up with the same energy table, so the same
    exchange decisions (restated here from reference parallel_tempering.py:234-258) give the same
    rung -> replica map on every rank."""
    from spin_glass_anneal_rl_b200.annealing.multi_gpu import (gather_energies, shard_replicas_split,
                                                               sum_over_ranks)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K, Lad = 8, 3                       # three 8-rung ladders = 24 replicas over 2 ranks:
        total = K * Lad                     # the middle ladder is cut in half
        sh = shard_replicas_split(total, world, rank, K)
        g = torch.Generator().manual_seed(11)
        all_e = torch.randn(total, generator=g)
        table = gather_energies(all_e[sh.start:sh.start + sh.count].clone(), total)
        ok = torch.equal(table, all_e) and sh.count == total // world
        temps = torch.logspace(1, -1, K, dtype=torch.float64)
        rep_at = torch.arange(total).reshape(Lad, K).clone()
        gu = torch.Generator().manual_seed(99)         # "shared counter RNG": same on every rank
        for parity in (0, 1, 0):
            for lad in range(Lad):
                for k in range(parity, K - 1, 2):
                    a, b = int(rep_at[lad, k]), int(rep_at[lad, k + 1])
                    p = min(1.0, float(torch.exp((1 / temps[k + 1] - 1 / temps[k]) *
                                                 (table[b].double() - table[a].double()))))
                    if float(torch.rand((), generator=gu, dtype=torch.float64)) < p:
                        rep_at[lad, k], rep_at[lad, k + 1] = b, a
        maps = [torch.empty_like(rep_at) for _ in range(world)]
        dist.all_gather(maps, rep_at)
        ok = ok and all(torch.equal(m, rep_at) for m in maps)
        s = sum_over_ranks(torch.tensor([rank + 1, 10]))
        ok = ok and s.tolist() == [3, 20]
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_split_ladder_energy_allgather_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_c1, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_split_shards():
    from spin_glass_anneal_rl_b200.annealing.multi_gpu import shard_replicas_split
    seen = []
    for r in range(8):
        sh = shard_replicas_split(64, 8, r, 64)      # ONE 64-rung ladder over 8 GPUs
        assert sh.count == 8
        seen += list(range(sh.start, sh.start + sh.count))
    assert seen == list(range(64))
    with pytest.raises(ValueError):
        shard_replicas_split(10, 4, 0, 5)
