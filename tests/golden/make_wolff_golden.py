#!/usr/bin/env python
"""Record golden traces of UpdateRule.WOLFF from the UNMODIFIED reference (device='cpu', dense J).

Run in the build container only (needs /root/reference):

    OMP_NUM_THREADS=1 python tests/golden/make_wolff_golden.py

Writes tests/golden/wolff_*.npz, same fields as the sa_* fixtures of make_golden.py (the run goes
through GPUAnnealer.anneal(model, UpdateRule.WOLFF), i.e. SpinDynamics.sweep() calling
_wolff_cluster_dense n times per sweep, core/spin_dynamics.py:193-262).  The oracle must reproduce
each of them from the seed alone (tests/test_oracle_golden.py); the CUDA kernel is then replayed
against the oracle's trace of the same run.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import run_pt, run_sa, sym_gauss  # noqa: E402  (imports the reference)


def lattice_couplings(L, rng, p_neg=0.8):
    """L x L periodic lattice as a dense matrix: bonds -1 with probability p_neg (the sign the
    reference's cluster growth follows), +1 otherwise."""
    n = L * L
    J = np.zeros((n, n), np.float32)
    for x in range(L):
        for y in range(L):
            i = x * L + y
            for j in (((x + 1) % L) * L + y, x * L + (y + 1) % L):
                v = -1.0 if rng.random() < p_neg else 1.0
                J[i, j] = J[j, i] = v
    return J


def main():
    rng = np.random.default_rng(20261019)

    # integer couplings on a 6 x 6 lattice, mostly of the sign that grows clusters (bit-exact target)
    J = lattice_couplings(6, rng)
    n = J.shape[0]
    h = rng.integers(-1, 2, size=n).astype(np.float32)
    s0 = (rng.integers(0, 2, size=n) * 2 - 1).astype(np.float32)
    run_sa("wolff_lattice_int_n36", J, h, s0, seed=31, n_sweeps=12, T0=4.0, Tf=0.5,
           params={"alpha": 0.85}, record_interval=2, rule="wolff")

    # dense +-1 couplings: every site is a neighbour of every other one, large clusters
    a = rng.integers(0, 2, size=(20, 20)) * 2 - 1
    Jd = np.triu(a, 1)
    Jd = (Jd + Jd.T).astype(np.float32)
    sd = (rng.integers(0, 2, size=20) * 2 - 1).astype(np.float32)
    run_sa("wolff_dense_int_n20", Jd, np.zeros(20, np.float32), sd, seed=32, n_sweeps=10, T0=6.0,
           Tf=1.0, params={"alpha": 0.8}, record_interval=1, rule="wolff")

    # Gaussian couplings (float probabilities 1 - exp(2 J / T))
    Jg = sym_gauss(24, rng, 0.8)
    hg = (0.3 * rng.standard_normal(24)).astype(np.float32)
    sg = (rng.integers(0, 2, size=24) * 2 - 1).astype(np.float32)
    run_sa("wolff_gauss_float_n24", Jg, hg, sg, seed=33, n_sweeps=10, T0=3.0, Tf=0.3,
           params={"alpha": 0.8}, record_interval=3, rule="wolff")

    # asymmetric couplings: the growth reads the ROW of the dequeued site (couplings[current, nb])
    Ja = (rng.integers(-2, 3, size=(16, 16))).astype(np.float32)
    np.fill_diagonal(Ja, 0.0)
    sa = (rng.integers(0, 2, size=16) * 2 - 1).astype(np.float32)
    run_sa("wolff_asym_int_n16", Ja, np.zeros(16, np.float32), sa, seed=34, n_sweeps=8, T0=5.0,
           Tf=1.0, params={"alpha": 0.8}, record_interval=1, rule="wolff")

    # parallel tempering with the cluster move (every replica's SpinDynamics runs _wolff_update);
    # the "ptw_" prefix keeps it apart from the single-spin pt_* fixtures
    Jl = lattice_couplings(5, rng, p_neg=0.7)
    run_pt("ptw_lattice_int_n25_r4", Jl, np.zeros(25, np.float32), seed=35, n_replicas=4, n_sweeps=14,
           tmin=1.5, tmax=6.0, exchange_interval=3, record_interval=2, rule="wolff")


if __name__ == "__main__":
    main()
