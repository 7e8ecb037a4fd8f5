"""Synthetic instances of the BASELINE.json configurations, generated vectorised.

The reference's own encoders call IsingModel.set_coupling once per coupling, which for sparse
models is a dense round trip of the whole matrix (core/ising_model.py:94-99): encoding TSP-64
or the 50k-spin scheduler that way takes hours.  These functions re-derive the SAME J and h in
closed form; tests/test_instances.py pins them against encodings recorded from the reference on
small instances (tests/golden/inst_*.npz).  Instance generators only -- not part of the product.
"""
import numpy as np


def sk(n=4096, seed=3003):
    """cfg3: Sherrington-Kirkpatrick, research/experimental_validation.py:112-131."""
    rs = np.random.RandomState(seed)
    G = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((G + G.T) / 2).astype(np.float32)
    np.fill_diagonal(J, 0.0)
    return J, np.zeros(n, np.float32)


def random_dense(n=100, seed=1001):
    """cfg1: A ~ N(0,1), J = (A + A^T)/2, zero diagonal, h = 0.5 N(0,1) (torch generator)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, generator=g)
    J = (A + A.T) / 2
    J.fill_diagonal_(0.0)
    h = 0.5 * torch.randn(n, generator=g)
    return J.numpy().astype(np.float32), h.numpy().astype(np.float32)


def cardinality_terms(n_group, k, penalty):
    """CardinalityConstraint -> EqualityConstraint(target = 2k - n, weight = penalty / 4)
    (core/constraints.py:147-158, 73-92): field per member, coupling per pair."""
    lam, target = penalty / 4.0, 2 * k - n_group
    return lam * (1.0 - 2.0 * target), 2.0 * lam


def tsp_ising(xy, penalty=100.0):
    """cfg4: TSP position encoding, spin (city c, position p) = c * n + p
    (problems/routing.py:193-328): -d(c, c') between consecutive positions, cardinality-1
    penalties on every city row and every position column; penalties auto-scaled by
    sqrt(n / 50) above 50 cities (:236-241).  Returns dense float32 (J, h)."""
    xy = np.asarray(xy, np.float64)
    n = xy.shape[0]
    if n > 50:
        penalty = penalty * np.sqrt(n / 50.0)
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    field, coupling = cardinality_terms(n, 1, penalty)
    N = n * n
    J = np.zeros((n, n, n, n), np.float64)   # [c, p, c', p']
    idx = np.arange(n)
    # same city, different position / same position, different city
    J[idx, :, idx, :] += coupling
    J[:, idx, :, idx] += coupling
    # objective: consecutive positions, different cities
    for p in range(n):
        q = (p + 1) % n
        J[:, p, :, q] = -d
        J[:, q, :, p] = -d.T
    J = J.reshape(N, N)
    np.fill_diagonal(J, 0.0)
    # same-city couplings were overwritten on the diagonal blocks c == c' by the objective loop
    # only where d == 0 (c == c'): restore them
    for c in range(n):
        blk = J[c * n:(c + 1) * n, c * n:(c + 1) * n]
        blk[:] = coupling
        np.fill_diagonal(blk, 0.0)
    h = np.full(N, 2.0 * field)
    return J.astype(np.float32), h.astype(np.float32)


def random_tsp(n_cities=64, seed=4004, area=100.0):
    """TSPProblem.generate_random_instance draw order (problems/routing.py:95-131): per city
    x, y, then a demand for every city but the depot; then one vehicle capacity."""
    rs = np.random.RandomState(seed)
    xy = np.zeros((n_cities, 2))
    for i in range(n_cities):
        xy[i, 0] = rs.uniform(0, area)
        xy[i, 1] = rs.uniform(0, area)
        if i > 0:
            rs.uniform(1.0, 10.0)
    return xy


def scheduling_ising(duration, due_date, cost_rate, penalty=100.0):
    """cfg5: SimpleScheduler (problems/simple_scheduler.py:67-127), spin = task * n_agents + agent:
    coupling penalty/2 inside every task's clique of agents; the field is OVERWRITTEN by
    duration * cost_rate + 0.1 * max(0, duration - due_date) (:119-123).  Returns CSR
    (rowptr, colidx, val) and h."""
    duration, due_date, cost_rate = (np.asarray(a, np.float64) for a in (duration, due_date, cost_rate))
    T, A = duration.shape[0], cost_rate.shape[0]
    _, coupling = cardinality_terms(A, 1, penalty)
    n = T * A
    base = (np.arange(n) // A) * A
    cols = base[:, None] + np.arange(A)[None, :]
    keep = cols != np.arange(n)[:, None]
    colidx = cols[keep].reshape(n, A - 1).astype(np.int32)
    rowptr = (np.arange(n + 1) * (A - 1)).astype(np.int64)
    val = np.full(colidx.size, coupling, np.float32)
    due_pen = np.where(due_date > 0, np.maximum(0.0, duration - due_date) * 0.1, 0.0)
    h = (duration[:, None] * cost_rate[None, :] + due_pen[:, None]).reshape(n).astype(np.float32)
    return rowptr, colidx.reshape(-1), val, h


def random_scheduling(n_tasks=500, n_agents=100, seed=5005, max_duration=20.0, horizon=100.0):
    """SimpleScheduler.generate_random_instance draw order (:37-65)."""
    rs = np.random.RandomState(seed)
    dur, due = np.zeros(n_tasks), np.zeros(n_tasks)
    for i in range(n_tasks):
        dur[i] = rs.uniform(5.0, max_duration)
        due[i] = rs.uniform(dur[i], horizon * 0.8)
    rate = np.array([rs.uniform(0.5, 2.0) for _ in range(n_agents)])
    return dur, due, rate


def ea_lattice_bonds(L=256, seed=2002, periodic=False):
    """Bond arrays of the 2D +-J lattice: Jx[x, y] couples (x, y)-(x+1, y), Jy[x, y] couples
    (x, y)-(x, y+1), int8 in {-1, 0, +1}; open boundaries (zeros in the last row / column) as the
    reference's generator actually builds it (research/experimental_validation.py:134-180)."""
    rs = np.random.RandomState(seed)
    Jx = np.zeros((L, L), np.int8)
    Jy = np.zeros((L, L), np.int8)
    if periodic:
        Jx[:] = rs.choice([-1, 1], size=(L, L))
        Jy[:] = rs.choice([-1, 1], size=(L, L))
    else:
        Jx[:L - 1, :] = rs.choice([-1, 1], size=(L - 1, L))
        Jy[:, :L - 1] = rs.choice([-1, 1], size=(L, L - 1))
    return Jx, Jy


def lattice_csr(Jx, Jy):
    """CSR rows of the symmetric coupling matrix of a lattice given by its bond arrays."""
    L = Jx.shape[0]
    n = L * L
    xs, ys = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    a = (xs * L + ys).ravel()
    bx = (((xs + 1) % L) * L + ys).ravel()
    by = (xs * L + (ys + 1) % L).ravel()
    rows = np.concatenate([a, bx, a, by])
    cols = np.concatenate([bx, a, by, a])
    vals = np.concatenate([Jx.ravel(), Jx.ravel(), Jy.ravel(), Jy.ravel()]).astype(np.float32)
    keep = vals != 0
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr), cols.astype(np.int32), vals, np.zeros(n, np.float32)


def ea_lattice(L=256, seed=2002):
    """cfg2: 2D Edwards-Anderson +-J, open boundaries, spin = x * L + y.  Returns CSR + h = 0."""
    return lattice_csr(*ea_lattice_bonds(L, seed))


def checkerboard_sequence(L):
    """Site order of one checkerboard sweep: all (x + y) even sites in row-major order, then all
    odd ones (the order sg_sweep_lattice.cu realises in parallel)."""
    xs, ys = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    site = (xs * L + ys).ravel()
    color = ((xs + ys) & 1).ravel()
    return np.concatenate([site[color == 0], site[color == 1]]).astype(np.int32)


def csr_to_dense(rowptr, colidx, val, n):
    J = np.zeros((n, n), np.float32)
    for i in range(n):
        J[i, colidx[rowptr[i]:rowptr[i + 1]]] = val[rowptr[i]:rowptr[i + 1]]
    return J
