"""The reference's operator interface (CUDAKernelManager, annealing/cuda_kernels.py:228-436).

CPU part: the oracle's restatements of the three loops reproduce the outputs recorded from the
reference itself (tests/golden/op_*.npz, made by tests/golden/make_operator_golden.py).
GPU part: the CUDA-backed CUDAKernelManager of this package gives the same results through the
same calls -- bit-exact on integer couplings; float couplings within the stated tolerances."""
import glob
import os

import numpy as np
import pytest

from conftest import has_cuda

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def _names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, prefix + "*.npz")))


# ------------------------------------------------------------------ oracle pinned to the reference
@pytest.mark.parametrize("name", _names("op_metropolis_"))
def test_oracle_metropolis_operator_matches_reference(oracle, name):
    g = _load(name)
    s, acc, changes, placed = oracle.operator_metropolis_update(
        g["J"], g["h"], g["spins0"], float(g["temperature"]), int(g["n_updates"]),
        uniform_stream=g["stream"])
    assert np.array_equal(s, g["spins"])
    assert acc == int(g["accepted"])
    assert int((placed != 0.5).sum()) <= int(g["stream_used"])
    if "int" in name:
        assert np.array_equal(changes, g["energy_changes"])
    else:
        assert np.allclose(changes, g["energy_changes"], rtol=0, atol=2e-5)
    # the positional form of the same stream gives the same trajectory
    s2, acc2, changes2, _ = oracle.operator_metropolis_update(
        g["J"], g["h"], g["spins0"], float(g["temperature"]), int(g["n_updates"]), uniforms=placed)
    assert np.array_equal(s2, s) and acc2 == acc and np.array_equal(changes2, changes)


@pytest.mark.parametrize("name", _names("op_energy_"))
def test_oracle_energy_operator_matches_reference(oracle, name):
    g = _load(name)
    e = oracle.operator_energy(g["J"], g["h"], g["spins"])
    assert e == pytest.approx(float(g["energy"]), rel=2e-6, abs=1e-5)


@pytest.mark.parametrize("name", _names("op_exchange_"))
def test_oracle_exchange_operator_matches_reference(oracle, name):
    g = _load(name)
    S, E, acc = oracle.operator_exchange(g["spins0"], g["energies0"], g["temperatures"], g["uniforms"])
    assert acc == int(g["accepted"])
    assert np.array_equal(S, g["spins"]) and np.array_equal(E, g["energies"])


def test_operator_manager_without_cuda_raises():
    """No CPU fallback: on a CPU device every operator call fails loudly."""
    import torch
    from spin_glass_anneal_rl_b200.annealing.cuda_kernels import CUDAKernelManager, GPUMemoryOptimizer
    from spin_glass_anneal_rl_b200.utils.exceptions import DeviceError
    mgr = CUDAKernelManager(torch.device("cpu"))
    assert mgr.compiled_kernels == {}
    s, J, h = torch.ones(4), torch.zeros(4, 4), torch.zeros(4)
    with pytest.raises(DeviceError):
        mgr.metropolis_update_optimized(s, J, h, 1.0)
    with pytest.raises(DeviceError):
        mgr.compute_energy_optimized(s, J, h)
    with pytest.raises(DeviceError):
        mgr.parallel_tempering_exchange_optimized(torch.ones(2, 4), torch.zeros(2), torch.ones(2))
    opt = GPUMemoryOptimizer(torch.device("cpu"))
    assert opt.get_optimal_batch_size(100, available_memory=1 << 30) == 64
    assert opt.get_optimal_batch_size(20000, available_memory=1 << 30) == 1
    assert opt.optimize_coupling_matrix_storage(torch.eye(8)).is_sparse
    assert not opt.optimize_coupling_matrix_storage(torch.ones(8, 8)).is_sparse


# ------------------------------------------------------------------ CUDA operators
@pytest.fixture(scope="module")
def manager():
    if not has_cuda():
        pytest.skip("needs a CUDA device")
    import torch
    from spin_glass_anneal_rl_b200.annealing.cuda_kernels import CUDAKernelManager
    return CUDAKernelManager(torch.device("cuda", 0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names("op_metropolis_"))
def test_cuda_metropolis_operator_replays_reference(manager, oracle, name):
    import torch
    g = _load(name)
    T, nu = float(g["temperature"]), int(g["n_updates"])
    _, _, _, placed = oracle.operator_metropolis_update(g["J"], g["h"], g["spins0"], T, nu,
                                                        uniform_stream=g["stream"])
    spins = torch.from_numpy(g["spins0"].copy()).cuda()
    out, acc, changes = manager.metropolis_update_optimized(
        spins, torch.from_numpy(g["J"]).cuda(), torch.from_numpy(g["h"]).cuda(), T, nu,
        uniforms=torch.from_numpy(placed))
    assert out is spins                                   # updated in place, like the reference
    assert np.array_equal(out.cpu().numpy(), g["spins"])
    assert acc == int(g["accepted"])
    if "int" in name:
        assert np.array_equal(changes.cpu().numpy(), g["energy_changes"])
    else:
        # float couplings: the kernel accumulates the local field in a different order
        assert np.allclose(changes.cpu().numpy(), g["energy_changes"], rtol=0, atol=5e-5)


@pytest.mark.gpu
def test_cuda_metropolis_operator_philox_statistics(manager):
    """Default RNG: sequential-order Metropolis at fixed T reaches the same mean energy as the
    oracle's loop (one replica, many passes; tolerance 5 sigma of the oracle's block means)."""
    import torch
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    n, T = 48, 2.0
    a = rng.integers(-1, 2, size=(n, n)).astype(np.float32)
    J = np.triu(a, 1)
    J = (J + J.T).astype(np.float32)
    h = np.zeros(n, np.float32)
    Jd, hd = torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda()
    s = torch.from_numpy((rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)).cuda()
    manager.metropolis_update_optimized(s, Jd, hd, T, 200)
    e_gpu = []
    for _ in range(40):
        _, acc, _ = manager.metropolis_update_optimized(s, Jd, hd, T, 25)
        e_gpu.append(manager.compute_energy_optimized(s, Jd, hd))
        assert 0 < acc < 25 * n
    so = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    so, _, _, _ = orc.operator_metropolis_update(J, h, so, T, 50, uniform_stream=rng.random(50 * n))
    e_cpu = []
    for _ in range(40):
        so, _, _, _ = orc.operator_metropolis_update(J, h, so, T, 25, uniform_stream=rng.random(25 * n))
        e_cpu.append(orc.operator_energy(J, h, so))
    se = np.sqrt(np.var(e_cpu) / len(e_cpu) + np.var(e_gpu) / len(e_gpu))
    assert abs(np.mean(e_gpu) - np.mean(e_cpu)) < 5 * se + 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names("op_energy_"))
def test_cuda_energy_operator_matches_reference(manager, name):
    import torch
    g = _load(name)
    e = manager.compute_energy_optimized(torch.from_numpy(g["spins"]).cuda(), torch.from_numpy(g["J"]).cuda(),
                                         torch.from_numpy(g["h"]).cuda())
    if "int" in name:
        assert e == float(g["energy"])
    else:
        assert e == pytest.approx(float(g["energy"]), rel=2e-6, abs=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names("op_exchange_"))
@pytest.mark.parametrize("dtype", ["float32", "int8"])
def test_cuda_exchange_operator_replays_reference(manager, name, dtype):
    import torch
    g = _load(name)
    S = torch.from_numpy(g["spins0"].copy()).to(getattr(torch, dtype)).cuda()
    E = torch.from_numpy(g["energies0"].copy()).cuda()
    acc = manager.parallel_tempering_exchange_optimized(
        S, E, torch.from_numpy(g["temperatures"]).cuda(), uniforms=torch.from_numpy(g["uniforms"]))
    assert acc == int(g["accepted"])
    assert np.array_equal(S.cpu().numpy().astype(np.float32), g["spins"])
    assert np.array_equal(E.cpu().numpy(), g["energies"])


@pytest.mark.gpu
def test_cuda_exchange_operator_large_and_strided(manager, oracle):
    """2048 replicas x 4096 spins, rows inside a wider matrix (row stride > row length), Philox
    draws: the result must be a permutation of the rows consistent with the swapped energies."""
    import torch
    R, n = 2048, 4096
    gen = torch.Generator(device="cuda").manual_seed(3)
    big = (torch.randint(0, 2, (R, n + 64), device="cuda", generator=gen) * 2 - 1).float()
    S = big[:, :n]
    tag = torch.arange(R, device="cuda", dtype=torch.float32)
    S[:, 0] = tag                                   # row identity
    E0 = torch.randn(R, device="cuda", generator=gen) * 5
    E = E0.clone()
    T = torch.from_numpy(np.geomspace(10.0, 0.1, R).astype(np.float32)).cuda()
    pad_before = big[:, n:].clone()
    acc = manager.parallel_tempering_exchange_optimized(S, E, T)
    assert 0 < acc < R
    src = S[:, 0].long()
    assert torch.equal(torch.sort(src).values, torch.arange(R, device="cuda"))
    assert torch.equal(E, E0[src])                   # energies travelled with their rows
    assert torch.equal(big[:, n:], pad_before)       # bytes outside the rows untouched
    moved = int((src != torch.arange(R, device="cuda")).sum())
    assert moved > 0


@pytest.mark.gpu
def test_cuda_operator_cache_is_not_fooled_by_reused_addresses(manager, oracle):
    """a loop over same-sized instances frees J and allocates the next one at
    the same address with version 0.  Every call must still answer for the model it was given."""
    import torch
    rng = np.random.default_rng(17)
    n = 48
    s = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    h = torch.zeros(n, device="cuda")
    seen = set()
    for _ in range(12):
        a = rng.integers(-3, 4, size=(n, n))
        J = np.triu(a, 1)
        J = (J + J.T).astype(np.float32)
        Jd = torch.from_numpy(J).cuda()
        seen.add(Jd.data_ptr())
        e = manager.compute_energy_optimized(torch.from_numpy(s).cuda(), Jd, h)
        assert e == oracle.energy(J, np.zeros(n, np.float32), s)
        del Jd
    assert len(seen) < 12, "the allocator never reused an address: the test did not test anything"
