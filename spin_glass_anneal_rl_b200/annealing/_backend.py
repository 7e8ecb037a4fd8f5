"""Glue between the host-side model container and the GPU engine.

An ``Engine`` (C-ABI handle: J^T and h resident in HBM) is cached on the model object
and re-uploaded only when the coupling / field tensors change identity or version --
the RL environment calls anneal() many times on the same J
(reference rl_integration/environment.py:318-336).  There is no CPU path: without a
CUDA device (or without libsg_b200.so) this raises DeviceError.
"""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..utils.exceptions import DeviceError

# Dense models up to this many spins stay dense: <= 4096 on the tensor-core sweep, up to 7168 on the
# sequential-FMA kernel (register-resident fields); beyond, the model goes to the sparse (CSR) kernel.
DENSE_LIMIT = 7168

RULE_NAMES = {"metropolis": "metropolis", "glauber": "glauber", "heat_bath": "heat_bath", "wolff": "wolff"}


def rule_name(update_rule) -> str:
    name = getattr(update_rule, "value", update_rule)
    if name not in RULE_NAMES:
        raise NotImplementedError(
            f"update rule {name!r} is not available on the B200 sweep path "
            "(Metropolis, Glauber, heat bath and Wolff are)")
    return name


def require_dense_for_wolff(eng) -> None:
    """The cluster move reads coupling rows of a dense model (the reference's sparse branch,
    core/spin_dynamics.py:264-323, walks the COO entries and is not mirrored)."""
    if getattr(eng, "kind", "dense") != "dense":
        raise NotImplementedError(
            f"UpdateRule.WOLFF needs a dense model (n <= {DENSE_LIMIT}); this model runs on the "
            f"{eng.kind} kernel")


class _Signature:
    """Identity of the (J, h) an engine was built from.  It HOLDS the two tensors: ids (and data
    pointers) of freed tensors are reused, so only `is` on live objects plus the version counter
    identifies a model."""

    def __init__(self, model):
        self.J, self.h = model.couplings, model.external_fields
        self.vJ, self.vh = self.J._version, self.h._version

    def matches(self, model) -> bool:
        J, h = model.couplings, model.external_fields
        return J is self.J and h is self.h and J._version == self.vJ and h._version == self.vh


def _lattice_bonds(rows, cols, vals, h, n):
    """(Jx, Jy) if the couplings are a 2D +-J nearest-neighbour lattice with spin = x * L + y and
    h = 0 (the Edwards-Anderson instances of research/experimental_validation.py:134-180), else
    None.  Such models run on the checkerboard multi-spin-coded kernel."""
    L = int(round(np.sqrt(n)))
    if L * L != n or L < 4 or np.any(h != 0) or vals.size == 0 or np.any(np.abs(vals) != 1.0):
        return None
    up = rows < cols
    if 2 * int(up.sum()) != rows.size:
        return None
    r, c, v = rows[up], cols[up], vals[up]
    x, y, d = r // L, r % L, c - r
    Jx = np.zeros((L, L), np.int8)
    Jy = np.zeros((L, L), np.int8)
    right = (d == 1) & (y < L - 1)
    down = d == L
    wrap_r = (d == L - 1) & (y == 0)          # (x, 0) -- (x, L-1): bond of site (x, L-1)
    wrap_d = (d == n - L) & (x == 0)          # (0, y) -- (L-1, y): bond of site (L-1, y)
    if not np.all(right | down | wrap_r | wrap_d) or ((wrap_r.any() or wrap_d.any()) and L % 2):
        return None
    Jy[x[right], y[right]] = v[right]
    Jx[x[down], y[down]] = v[down]
    Jy[x[wrap_r], L - 1] = v[wrap_r]
    Jx[L - 1, y[wrap_d]] = v[wrap_d]
    # symmetric? rebuild the lower triangle and compare
    lo = ~up
    key_up = np.sort(r.astype(np.int64) * n + c)
    key_lo = np.sort(cols[lo].astype(np.int64) * n + rows[lo])
    if not np.array_equal(key_up, key_lo):
        return None
    order_u, order_l = np.argsort(r.astype(np.int64) * n + c), np.argsort(cols[lo].astype(np.int64) * n + rows[lo])
    if not np.array_equal(v[order_u], vals[lo][order_l]):
        return None
    return Jx, Jy


def _clique_groups(rows, cols, vals, n):
    """(group_of, coupling) if J_ij is one constant per group for every pair i != j of the group
    and zero between groups (the one-hot / cardinality penalty structure of
    core/constraints.py:126-158, e.g. SimpleScheduler), else None."""
    if rows.size == 0 or np.any(rows == cols):
        return None
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    deg = np.bincount(rows, minlength=n)
    rep = np.arange(n)
    np.minimum.at(rep, rows, cols)                     # smallest member of the neighbourhood
    if np.any(rep[rows] != rep[cols]):
        return None
    uniq, group_of = np.unique(rep, return_inverse=True)
    size = np.bincount(group_of, minlength=uniq.size)
    if np.any(deg != size[group_of] - 1):
        return None
    if rows.size > 1 and np.any((rows[1:] == rows[:-1]) & (cols[1:] == cols[:-1])):
        return None
    coupling = np.zeros(uniq.size, np.float32)
    coupling[group_of[rows]] = vals
    if np.any(coupling[group_of[rows]] != vals):
        return None
    if 4 * n + 64 * uniq.size > 227 * 1024:
        return None
    return group_of.astype(np.int32), coupling


def site_order_for(eng: Engine, requested: str) -> str:
    """Lattice models are swept in checkerboard order (every site once per sweep)."""
    return "checkerboard" if getattr(eng, "kind", "dense") == "lattice" else requested


def engine_for(model, device_index: int = 0) -> Engine:
    """The engine holding this model's couplings (built / refreshed on demand)."""
    if not hasattr(model, "couplings") or not hasattr(model, "spins"):
        raise AttributeError("anneal() needs an IsingModel (couplings / external_fields / spins)")
    try:
        cached = getattr(model, "_sg_engine", None)
        if cached is not None and cached[0].matches(model) and cached[1].device_index == device_index:
            return cached[1]
        sig = _Signature(model)
        eng = cached[1] if cached is not None and cached[1].device_index == device_index \
            else Engine(device_index)
        J = model.couplings
        n = int(J.shape[0])
        dense_ok = n <= DENSE_LIMIT
        if dense_ok and n > 4096 and J.is_sparse:
            # a sparse model of this size is only worth a dense upload if it is not actually sparse
            nnz = int(J._nnz()) if not J.is_coalesced() else int(J.indices().shape[1])
            dense_ok = nnz > 0.25 * n * n
        if not dense_ok:
            # sparse path (K1-CSR): the 50k-spin scheduling QUBOs and lattices the reference's
            # callers build as COO (problems/base.py:107-116) cannot be held as dense matrices
            coo = (J if J.is_sparse else J.to_sparse()).coalesce().cpu()
            rows, cols = coo.indices()[0].numpy(), coo.indices()[1].numpy()
            vals = coo.values().to(torch.float32).numpy()
            hh = model.external_fields.to(torch.float32).cpu().numpy()
            bonds = _lattice_bonds(rows, cols, vals, hh, n)
            groups = None if bonds is not None else _clique_groups(rows, cols, vals, n)
            if bonds is not None:
                eng.set_model_lattice2d(*bonds)
                eng.kind = "lattice"
            elif groups is not None:
                eng.set_model_groups(groups[0], groups[1], hh)
                eng.kind = "groups"
            else:
                order = np.lexsort((cols, rows))
                rowptr = np.zeros(n + 1, np.int64)
                np.add.at(rowptr, rows + 1, 1)
                eng.set_model_csr(np.cumsum(rowptr), cols[order].astype(np.int32), vals[order], hh)
                eng.kind = "csr"
        else:
            eng.kind = "dense"
            dense = (J.to_dense() if J.is_sparse else J).to(torch.float32)
            eng.set_model(dense, model.external_fields.to(torch.float32))
        model._sg_engine = (sig, eng)
        return eng
    except (RuntimeError, OSError) as exc:  # missing library, no GPU, CUDA failure
        if isinstance(exc, DeviceError):
            raise
        raise DeviceError(f"B200 annealing engine unavailable: {exc}") from exc


def mix_seed(*parts: int) -> int:
    """64-bit Philox key from (seed, rank, launch, ...): splitmix64 over the parts, so that large
    seeds do not alias and neighbouring (seed, launch) pairs give unrelated streams."""
    x = 0x9E3779B97F4A7C15
    for p in parts:
        x = (x ^ (int(p) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
    return x


def random_spins(n_replicas: int, n: int, device, generator: torch.Generator) -> torch.Tensor:
    return (torch.randint(0, 2, (n_replicas, n), device=device, generator=generator,
                          dtype=torch.int8) * 2 - 1).to(torch.int8)


def as_pm1_float(spins_i8: torch.Tensor) -> torch.Tensor:
    return spins_i8.to(torch.float32).cpu()
