"""The vectorised instance generators (tools/instances.py) reproduce the encodings recorded from
the reference's own problem classes (tests/golden/inst_*.npz, made by make_instance_golden.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import instances as inst  # noqa: E402


@pytest.mark.parametrize("name", ["inst_tsp5", "inst_tsp6"])
def test_tsp_encoding_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    J, h = inst.tsp_ising(g["xy"])
    assert np.array_equal(h, g["h"])
    assert np.allclose(J, g["J"], rtol=0, atol=1e-4)
    n = g["xy"].shape[0]
    assert np.array_equal(inst.random_tsp(n, int(g["seed"])), g["xy"])


@pytest.mark.parametrize("name", ["inst_sched_3x4", "inst_sched_6x5"])
def test_scheduling_encoding_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    rowptr, colidx, val, h = inst.scheduling_ising(g["duration"], g["due_date"], g["cost_rate"])
    n = h.shape[0]
    assert np.array_equal(inst.csr_to_dense(rowptr, colidx, val, n), g["J"])
    assert np.allclose(h, g["h"], rtol=1e-6, atol=0)
    T, A = g["duration"].shape[0], g["cost_rate"].shape[0]
    dur, due, rate = inst.random_scheduling(T, A, int(g["seed"]))
    assert np.allclose(dur, g["duration"]) and np.allclose(due, g["due_date"]) and np.allclose(rate, g["cost_rate"])


def test_full_size_shapes():
    rowptr, colidx, val, h = inst.scheduling_ising(*inst.random_scheduling(500, 100))
    assert h.shape[0] == 50000 and colidx.shape[0] == 50000 * 99 and rowptr[-1] == colidx.shape[0]
    rowptr, colidx, val, h = inst.ea_lattice(256)
    assert h.shape[0] == 65536 and rowptr[-1] == 2 * 2 * 256 * 255
    deg = np.diff(rowptr)
    assert deg.min() == 2 and deg.max() == 4           # open boundaries
    J = inst.csr_to_dense(*inst.ea_lattice(8)[:3], 64)
    assert np.array_equal(J, J.T) and set(np.unique(J)) <= {-1.0, 0.0, 1.0}
