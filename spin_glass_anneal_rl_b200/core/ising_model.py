"""IsingModel: the state/layout contract of the hot path.

Same public surface as the reference's ``spin_glass_rl.core.ising_model``
(reference core/ising_model.py:36-258): ``IsingModelConfig(n_spins,
coupling_strength, external_field_strength, use_sparse, device)`` and an
``IsingModel`` whose ``spins`` (float32, +-1), ``couplings`` (dense [n, n] or
sparse COO, both triangles stored) and ``external_fields`` are ordinary, publicly
mutable torch tensors on the host.  H = -1/2 s^T J s - h^T s.

``set_coupling`` on a sparse model is O(1): the reference densifies and re-sparsifies the whole
matrix per call (core/ising_model.py:94-99), which makes its own problem encoders
(problems/routing.py:275-294, core/constraints.py:360-388: one call per pair) take ~4 million
N^2 round trips for a 64-city TSP.  Here the writes are queued and folded into the COO tensor in
one vectorised pass the next time ``couplings`` is read, with the same last-write-wins and
zero-removal result.

The model is only the container the callers build; every sweep runs on the GPU
engine (annealing/_backend.py uploads J, h once and keeps them resident).  Unlike the
reference, the sparse representation works for every method: the reference slices a
sparse COO tensor in flip_spin (core/ising_model.py:133-135), which raises on current
torch, so every ProblemTemplate model failed on its first accepted flip.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch


@dataclass
class IsingModelConfig:
    n_spins: int
    coupling_strength: float = 1.0
    external_field_strength: float = 0.5
    use_sparse: bool = True
    device: str = "cpu"


def _random_spins(n: int, device) -> torch.Tensor:
    # same draw as the reference (core/ising_model.py:67,202) so seeded callers see the
    # same initial configuration
    return (torch.randint(0, 2, (n,), device=device) * 2 - 1).float()


class IsingModel:
    def __init__(self, config: IsingModelConfig):
        if not isinstance(config.n_spins, int) or config.n_spins <= 0:
            raise ValueError(f"n_spins must be a positive integer, got {config.n_spins!r}")
        self.config = config
        self.n_spins = config.n_spins
        self.device = torch.device(config.device)
        self.spins = _random_spins(self.n_spins, self.device)
        self._pending: Dict[int, float] = {}   # queued set_coupling writes: i * n + j -> strength
        if config.use_sparse:
            self.couplings = torch.sparse_coo_tensor(
                torch.empty((2, 0), dtype=torch.long), torch.empty(0),
                (self.n_spins, self.n_spins), device=self.device, check_invariants=False)
        else:
            self.couplings = torch.zeros((self.n_spins, self.n_spins), device=self.device)
        self.external_fields = torch.zeros(self.n_spins, device=self.device)
        self._energy_cache = None
        self._cache_valid = False

    # ------------------------------------------------------------------ couplings / fields
    @property
    def couplings(self) -> torch.Tensor:
        """Dense [n, n] or sparse COO tensor; publicly readable and assignable like the
        reference's attribute.  Reading folds the queued ``set_coupling`` writes in."""
        if self._pending:
            self._flush_pending()
        return self._couplings

    @couplings.setter
    def couplings(self, value: torch.Tensor) -> None:
        self._pending.clear()
        self._couplings = value
        self._cache_valid = False

    def _flush_pending(self) -> None:
        """Apply the queued writes to the COO tensor: entries written are replaced (last write
        wins), zero strengths are dropped (what dense[i, j] = v; to_sparse() leaves)."""
        n = self.n_spins
        keys_u = np.fromiter(self._pending.keys(), dtype=np.int64, count=len(self._pending))
        vals_u = np.fromiter(self._pending.values(), dtype=np.float32, count=len(self._pending))
        self._pending.clear()
        old = self._couplings.coalesce()
        idx = old.indices().cpu().numpy()
        keys_o = idx[0].astype(np.int64) * n + idx[1]
        vals_o = old.values().cpu().numpy().astype(np.float32)
        keep = ~np.isin(keys_o, keys_u)
        nz = vals_u != 0.0
        keys = np.concatenate([keys_o[keep], keys_u[nz]])
        vals = np.concatenate([vals_o[keep], vals_u[nz]])
        order = np.argsort(keys, kind="stable")
        keys, vals = keys[order], vals[order]
        ind = torch.from_numpy(np.stack([keys // n, keys % n]))
        self._couplings = torch.sparse_coo_tensor(ind, torch.from_numpy(vals), (n, n), device=self.device,
                                                  check_invariants=False, is_coalesced=True)

    def dense_couplings(self) -> torch.Tensor:
        """Couplings as a dense float32 [n, n] tensor (what the engine uploads)."""
        J = self.couplings
        return (J.to_dense() if J.is_sparse else J).to(torch.float32)

    def _store_couplings(self, dense: torch.Tensor) -> None:
        self.couplings = dense.to_sparse() if self.config.use_sparse else dense
        self._invalidate_cache()

    def set_coupling(self, i: int, j: int, strength: float) -> None:
        if not (0 <= i < self.n_spins and 0 <= j < self.n_spins):
            raise ValueError(f"Spin indices out of range: i={i}, j={j}, n_spins={self.n_spins}")
        if self._couplings.is_sparse:
            self._pending[i * self.n_spins + j] = float(strength)
            self._pending[j * self.n_spins + i] = float(strength)
            self._invalidate_cache()
        else:
            self.couplings[i, j] = strength
            self.couplings[j, i] = strength
            self._invalidate_cache()

    def get_coupling(self, i: int, j: int) -> float:
        if not (0 <= i < self.n_spins and 0 <= j < self.n_spins):
            raise ValueError(f"Spin indices out of range: i={i}, j={j}, n_spins={self.n_spins}")
        key = i * self.n_spins + j
        if key in self._pending:
            return self._pending[key]
        J = self._couplings
        if not J.is_sparse:
            return float(J[i, j].item())
        J = J.coalesce()
        idx = J.indices()
        sel = (idx[0] == i) & (idx[1] == j)
        return float(J.values()[sel].sum().item())

    def set_couplings_from_matrix(self, coupling_matrix: torch.Tensor) -> None:
        self._store_couplings(coupling_matrix.clone().to(self.device))

    def set_external_field(self, i: int, strength: float) -> None:
        self.external_fields[i] = strength
        self._invalidate_cache()

    def set_external_fields(self, fields: torch.Tensor) -> None:
        self.external_fields = fields.clone().to(self.device)
        self._invalidate_cache()

    # ------------------------------------------------------------------ single-configuration maths
    def _row(self, i: int) -> torch.Tensor:
        J = self.couplings
        if J.is_sparse:
            J = J.coalesce()
            idx, val = J.indices(), J.values()
            sel = idx[0] == i
            row = torch.zeros(self.n_spins, dtype=val.dtype, device=val.device)
            row.index_add_(0, idx[1][sel], val[sel])
            return row
        return J[i]

    def get_local_field(self, i: int) -> float:
        """sum_j J_ij s_j + h_i, diagonal included (reference :176-185)."""
        return torch.dot(self._row(i).float(), self.spins.float()).item() + self.external_fields[i].item()

    def flip_spin(self, i: int) -> float:
        """Flip spin i in place and return dE = 2 s_i (sum_j J_ij s_j + h_i) (reference :125-147)."""
        delta = 2.0 * self.spins[i].item() * self.get_local_field(i)
        self.spins[i] *= -1
        self._invalidate_cache()
        return delta

    def compute_energy(self) -> float:
        """-1/2 s^T J s - h^T s, cached until the next mutation (reference :149-174)."""
        if self._cache_valid and self._energy_cache is not None:
            return self._energy_cache
        s = self.spins.float()
        J = self.couplings
        Js = torch.sparse.mm(J, s.unsqueeze(1)).squeeze(1) if J.is_sparse else torch.mv(J.float(), s)
        energy = -0.5 * torch.dot(s, Js.float()).item() - torch.dot(self.external_fields.float(), s).item()
        self._energy_cache, self._cache_valid = energy, True
        return energy

    def get_magnetization(self) -> float:
        return self.spins.sum().item() / self.n_spins

    def set_spins(self, spins: torch.Tensor) -> None:
        self.spins = spins.clone().to(self.device)
        self._invalidate_cache()

    def get_spins(self) -> torch.Tensor:
        return self.spins.clone()

    def reset_to_random(self) -> None:
        self.spins = _random_spins(self.n_spins, self.device)
        self._invalidate_cache()

    def copy(self) -> "IsingModel":
        other = IsingModel(self.config)  # draws n randints exactly like the reference (:205-211)
        other.spins = self.spins.clone()
        other.couplings = self.couplings.clone()
        other.external_fields = self.external_fields.clone()
        return other

    # ------------------------------------------------------------------ serialisation
    def to_dict(self) -> Dict:
        c = self.config
        return {
            "config": {"n_spins": c.n_spins, "coupling_strength": c.coupling_strength,
                       "external_field_strength": c.external_field_strength,
                       "use_sparse": c.use_sparse, "device": c.device},
            "spins": self.spins.cpu().numpy(),
            "couplings": self.dense_couplings().cpu().numpy(),
            "external_fields": self.external_fields.cpu().numpy(),
        }

    @classmethod
    def from_dict(cls, data: Dict) -> "IsingModel":
        model = cls(IsingModelConfig(**data["config"]))
        model.spins = torch.from_numpy(data["spins"]).to(model.device)
        model._store_couplings(torch.from_numpy(data["couplings"]).to(model.device))
        model.external_fields = torch.from_numpy(data["external_fields"]).to(model.device)
        return model

    def _invalidate_cache(self) -> None:
        self._cache_valid = False

    def __repr__(self) -> str:
        return (f"IsingModel(n_spins={self.n_spins}, energy={self.compute_energy():.4f}, "
                f"magnetization={self.get_magnetization():.4f})")
