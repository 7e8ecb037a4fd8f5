#!/usr/bin/env python
"""Record the outputs of the reference's operator entry points (CUDAKernelManager,
annealing/cuda_kernels.py:228-436) on CPU, where they run the Python loops every upstream
user gets.  Build container only (needs /root/reference):

    OMP_NUM_THREADS=1 python tests/golden/make_operator_golden.py

torch.rand is patched to hand out a known stream, so the fixtures hold every uniform the
reference consumed, in order; tests/test_operator_api.py replays them through the oracle
(pin) and through the CUDA library."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from spin_glass_rl.annealing.cuda_kernels import CUDAKernelManager  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


class Stream:
    def __init__(self, values):
        self.values, self.pos, self._rand = values, 0, torch.rand

    def __enter__(self):
        def rand(*a, **k):
            v = float(self.values[self.pos])
            self.pos += 1
            return torch.full((1,), v, dtype=torch.float32)
        torch.rand = rand
        return self

    def __exit__(self, *exc):
        torch.rand = self._rand


def instance(rng, n, integer, diag=False):
    if integer:
        a = rng.integers(-2, 3, size=(n, n)).astype(np.float32)
        h = rng.integers(-2, 3, size=n).astype(np.float32)
    else:
        a = rng.normal(size=(n, n)).astype(np.float32)
        h = (0.5 * rng.normal(size=n)).astype(np.float32)
    J = np.triu(a, 1)
    J = (J + J.T).astype(np.float32)
    if diag:
        J[np.arange(n), np.arange(n)] = rng.integers(-2, 3, size=n).astype(np.float32)
    s = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    return J, h, s


def main():
    mgr = CUDAKernelManager(torch.device("cpu"))
    rng = np.random.default_rng(20261018)
    for name, n, integer, diag, T, nu in (("op_metropolis_int_n24", 24, True, False, 1.5, 3),
                                          ("op_metropolis_int_diag_n17", 17, True, True, 2.0, 2),
                                          ("op_metropolis_float_n20", 20, False, False, 0.75, 2)):
        J, h, s0 = instance(rng, n, integer, diag)
        stream = rng.random(n * nu).astype(np.float32)
        with Stream(stream) as st:
            spins, acc, changes = mgr.metropolis_update_optimized(
                torch.from_numpy(s0.copy()), torch.from_numpy(J), torch.from_numpy(h), T, nu)
            used = st.pos
        np.savez(os.path.join(OUT, name + ".npz"), J=J, h=h, spins0=s0, temperature=T, n_updates=nu,
                 stream=stream, stream_used=used, spins=spins.numpy(), accepted=acc,
                 energy_changes=changes.numpy(), torch_version=torch.__version__)
        print(name, "accepted", acc, "uniforms used", used)
    for name, n, integer, diag in (("op_energy_int_n32", 32, True, True), ("op_energy_float_n40", 40, False, False)):
        J, h, s0 = instance(rng, n, integer, diag)
        e = mgr.compute_energy_optimized(torch.from_numpy(s0), torch.from_numpy(J), torch.from_numpy(h))
        np.savez(os.path.join(OUT, name + ".npz"), J=J, h=h, spins=s0, energy=e)
        print(name, e)
    for name, R, n in (("op_exchange_r6", 6, 10), ("op_exchange_r33", 33, 7)):
        S = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.float32)
        E = rng.normal(scale=3.0, size=R).astype(np.float32)
        T = (np.geomspace(5.0, 0.2, R) if R < 10 else rng.uniform(0.2, 5.0, size=R)).astype(np.float32)
        u = rng.random(R - 1).astype(np.float32)
        St, Et = torch.from_numpy(S.copy()), torch.from_numpy(E.copy())
        with Stream(u) as st:
            acc = mgr.parallel_tempering_exchange_optimized(St, Et, torch.from_numpy(T))
            assert st.pos == R - 1
        np.savez(os.path.join(OUT, name + ".npz"), spins0=S, energies0=E, temperatures=T, uniforms=u,
                 spins=St.numpy(), energies=Et.numpy(), accepted=acc)
        print(name, "accepted", acc)


if __name__ == "__main__":
    main()
