"""Two GPUs, one process each (NCCL): ladders that span GPUs (collective C1) must evolve exactly
like the same replica set on one GPU -- same Philox keys (global replica ids), same exchange
decisions from the all-gathered energy table.  Skipped on a single-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

from conftest import has_cuda

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _instance(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(-2, 3, size=(n, n))
    J = np.triu(a, 1)
    return (J + J.T).astype(np.float32), rng.integers(-1, 2, size=n).astype(np.float32)


def _run(cfg_kwargs, J, h, shard, device):
    import torch
    import spin_glass_anneal_rl_b200 as sg
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=J.shape[0], use_sparse=False))
    m.set_couplings_from_matrix(torch.from_numpy(J))
    m.set_external_fields(torch.from_numpy(h))
    cfg = sg.ParallelTemperingConfig(device_index=device, shard=shard, **cfg_kwargs)
    pt = sg.ParallelTempering(cfg)
    res = pt.run(m)
    return dict(spins=pt._final_spins.cpu().numpy(), rung=pt._rung_replica.copy(),
                att=pt.exchange_attempts.copy(), acc=pt.exchange_accepts.copy(),
                best_e=res.best_energy, best_s=res.best_configuration.numpy().copy(),
                hist=np.array(pt.energy_histories), rates=np.array(res.acceptance_rate_history))


def _worker(rank, world, port, cases, out):
    import torch
    import torch.distributed as dist
    from spin_glass_anneal_rl_b200.annealing.multi_gpu import shard_replicas_split
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        for ci, (n, kw) in enumerate(cases):
            J, h = _instance(n, 100 + ci)
            total = kw["n_replicas"] * kw["n_ladders"]
            sh = shard_replicas_split(total, world, rank, kw["n_replicas"])
            r = _run(kw, J, h, sh, rank)
            out[(ci, rank)] = r
    finally:
        dist.destroy_process_group()


CASES = [
    (256, dict(n_replicas=16, n_ladders=1, n_sweeps=41, temp_min=0.5, temp_max=5.0, exchange_interval=5,
               record_interval=10, random_seed=11)),
    (1024, dict(n_replicas=8, n_ladders=9, n_sweeps=23, temp_min=0.4, temp_max=4.0, exchange_interval=3,
                record_interval=4, random_seed=12, exchange_method="all_pairs")),
]


def test_split_ladders_on_two_gpus_equal_one_gpu():
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), CASES, out), nprocs=world, join=True)
    for ci, (n, kw) in enumerate(CASES):
        J, h = _instance(n, 100 + ci)
        ref = _run(kw, J, h, None, 0)
        total = kw["n_replicas"] * kw["n_ladders"]
        per = total // world
        got = np.concatenate([out[(ci, r)]["spins"] for r in range(world)])
        assert np.array_equal(got, ref["spins"]), "configurations differ from the single-GPU run"
        for r in range(world):
            o = out[(ci, r)]
            assert o["spins"].shape[0] == per
            assert np.array_equal(o["rung"], ref["rung"]), "rung -> replica map differs"
            assert np.array_equal(o["att"], ref["att"]) and np.array_equal(o["acc"], ref["acc"])
            assert np.array_equal(o["hist"], ref["hist"])
            assert o["best_e"] == ref["best_e"] and np.array_equal(o["best_s"], ref["best_s"])
            assert np.allclose(o["rates"], ref["rates"], rtol=0, atol=1e-12)
        assert ref["acc"].sum() > 0
