"""The reference's device-operator interface, backed by the C ABI.

``CUDAKernelManager`` is the object the reference's annealers hold as ``self.cuda_kernels``
and call on their CUDA path (reference annealing/gpu_annealer.py:85, 205-250 and
annealing/parallel_tempering.py:78, 262-295); its three entry points are declared at
annealing/cuda_kernels.py:228-369.  Upstream their kernels never launch (the module-loading
call does not exist) and the Python loops at :371-436 run instead; those loops define the
behaviour mirrored here -- same names, arguments, return values, in-place effects:

* ``metropolis_update_optimized(spins, couplings, external_fields, temperature, n_updates)``
  -> ``(spins, accepted_flips, energy_changes)``: ``n_updates`` passes over the sites in index
  order; local field without the diagonal term; accept iff dE <= 0 or u < exp(-dE/T);
  ``energy_changes[i]`` accumulates the accepted dE of site i; ``spins`` is updated in place.
* ``compute_energy_optimized(spins, couplings, external_fields)`` -> ``float``:
  -1/2 s^T J s - h^T s (diagonal included, as at :398-403).
* ``parallel_tempering_exchange_optimized(spins_arrays, energies, temperatures)`` -> ``int``:
  one ordered pass over adjacent pairs with p = exp((1/T[i+1] - 1/T[i]) (E[i] - E[i+1])),
  configurations and energies swapped in place.

Each call goes to the CUDA library (sg_sweep on the sequential-FMA kernel / sg_batch_energies /
sg_exchange_chain).  There is no PyTorch fallback: on a non-CUDA device the calls raise
``DeviceError``.  The batched annealers of this package do not go through this per-call
interface (one launch per sweep of one replica is what it offers); it exists so that code
written against the reference's operator API runs on the GPU unchanged.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Optional, Tuple

import torch

from ..utils.exceptions import DeviceError


class CUDAKernelManager:
    """Reference annealing/cuda_kernels.py:126-369."""

    _MAX_CACHED_MODELS = 4

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.compiled_kernels = {}
        self._engines: "OrderedDict[tuple, object]" = OrderedDict()
        self._sweeps_done = 0
        self._exchange_rounds = 0
        self.seed = 0x5EED5EED
        self._compile_kernels()

    def _compile_kernels(self) -> None:
        """Load (building in-tree if needed) the CUDA library; nothing to do on a CPU device."""
        if self.device.type != "cuda" or not torch.cuda.is_available():
            return
        from .. import _lib
        self._lib = _lib.load()
        self.compiled_kernels = {"metropolis_update": "sg_sweep", "compute_energy": "sg_batch_energies",
                                 "parallel_tempering": "sg_exchange_chain"}

    # ------------------------------------------------------------------ helpers
    def _require(self, name: str) -> None:
        if name not in self.compiled_kernels:
            raise DeviceError(f"CUDAKernelManager.{name}: no CUDA device / library (there is no CPU fallback)",
                              {"device": str(self.device)})

    def _device_index(self) -> int:
        return self.device.index if self.device.index is not None else torch.cuda.current_device()

    def _engine(self, couplings: torch.Tensor, external_fields: torch.Tensor, zero_diagonal: bool):
        """Engine holding (J, h).  A cache entry KEEPS the caller's tensors and matches only the
        very same objects at the same version: addresses and ids of freed tensors are handed out
        again by the allocators, so they cannot identify a model."""
        for key, (Jc, hc, vJ, vh, zd, eng) in list(self._engines.items()):
            if (Jc is couplings and hc is external_fields and vJ == couplings._version
                    and vh == external_fields._version and zd == zero_diagonal):
                self._engines.move_to_end(key)
                return eng
        from ..engine import Engine
        J = couplings.to_dense() if couplings.is_sparse else couplings
        J = J.to(device=self.device, dtype=torch.float32)
        if zero_diagonal and bool(torch.diagonal(J).ne(0).any()):
            J = J.clone()
            J.fill_diagonal_(0.0)
        eng = Engine(self._device_index())
        eng.set_model(J, external_fields.to(device=self.device, dtype=torch.float32))
        eng.alloc_replicas(1)
        self._cache_serial = getattr(self, "_cache_serial", 0) + 1
        self._engines[self._cache_serial] = (couplings, external_fields, couplings._version,
                                             external_fields._version, zero_diagonal, eng)
        while len(self._engines) > self._MAX_CACHED_MODELS:
            self._engines.popitem(last=False)[1][-1].close()
        return eng

    # ------------------------------------------------------------------ the three operators
    def metropolis_update_optimized(self, spins: torch.Tensor, couplings: torch.Tensor,
                                    external_fields: torch.Tensor, temperature: float,
                                    n_updates: int = 1, *, uniforms: Optional[torch.Tensor] = None
                                    ) -> Tuple[torch.Tensor, int, torch.Tensor]:
        """Reference :228-282 / :371-397.  ``uniforms`` (optional, [n_updates, n]) injects the
        uniform of every attempt (attempt k of pass p uses uniforms[p, k]; the reference draws one
        only when dE > 0) -- replay / parity testing; default is the library's Philox stream."""
        self._require("metropolis_update")
        n = spins.shape[0]
        eng = self._engine(couplings, external_fields, zero_diagonal=True)
        s8 = torch.where(spins.to(self.device) >= 0, 1, -1).to(torch.int8).reshape(1, n)
        eng.set_spins(s8)
        eng.init_fields()
        acc0 = int(eng.accepted()[0].item())
        changes = torch.zeros((1, n), dtype=torch.float32, device=eng.device)
        u = None
        if uniforms is not None:
            u = uniforms.to(device=eng.device, dtype=torch.float32).reshape(1, n_updates, n).contiguous()
        eng.sweep(int(n_updates), torch.tensor([float(temperature)], dtype=torch.float64),
                  rule="metropolis", site_order="sequential", seed=self.seed,
                  sweep_base=self._sweeps_done, uniforms=u, track_best=False, kernel="simt",
                  site_energy_changes=changes)
        self._sweeps_done += int(n_updates)
        accepted = int(eng.accepted()[0].item()) - acc0
        spins.copy_(eng.spins()[0].to(device=spins.device, dtype=spins.dtype))
        return spins, accepted, changes[0].to(device=spins.device, dtype=spins.dtype)

    def compute_energy_optimized(self, spins: torch.Tensor, couplings: torch.Tensor,
                                 external_fields: torch.Tensor) -> float:
        """Reference :284-324 / :398-403."""
        self._require("compute_energy")
        eng = self._engine(couplings, external_fields, zero_diagonal=False)
        s8 = torch.where(spins.to(self.device) >= 0, 1, -1).to(torch.int8).reshape(1, -1)
        return float(eng.batch_energies(s8)[0].item())

    def parallel_tempering_exchange_optimized(self, spins_arrays: torch.Tensor, energies: torch.Tensor,
                                              temperatures: torch.Tensor, *,
                                              uniforms: Optional[torch.Tensor] = None) -> int:
        """Reference :326-369 / :405-436.  ``spins_arrays`` [R, n] and ``energies`` [R] are
        modified in place; ``uniforms`` (optional, [R-1]) injects the draw of every pair."""
        self._require("parallel_tempering")
        if spins_arrays.dim() != 2 or energies.shape[0] != spins_arrays.shape[0] \
                or temperatures.shape[0] != spins_arrays.shape[0]:
            raise ValueError("expected spins_arrays [R, n], energies [R], temperatures [R]")
        if not spins_arrays.is_cuda or spins_arrays.stride(1) != 1:
            raise DeviceError("spins_arrays must be a CUDA tensor with contiguous rows",
                              {"device": str(spins_arrays.device)})
        dev = spins_arrays.device
        e32 = energies.to(device=dev, dtype=torch.float32).contiguous()
        t32 = temperatures.to(device=dev, dtype=torch.float32).contiguous()
        u32 = None if uniforms is None else uniforms.to(device=dev, dtype=torch.float32).contiguous()
        n_acc = ctypes.c_int32(0)
        R, n = spins_arrays.shape
        esz = spins_arrays.element_size()
        from .._lib import check
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = self._lib.sg_exchange_chain(
                dev.index if dev.index is not None else torch.cuda.current_device(),
                ctypes.c_void_p(spins_arrays.data_ptr()), spins_arrays.stride(0) * esz, n * esz, R,
                ctypes.c_void_p(e32.data_ptr()), ctypes.c_void_p(t32.data_ptr()),
                ctypes.c_void_p(u32.data_ptr()) if u32 is not None else None,
                self.seed, self._exchange_rounds, ctypes.byref(n_acc), stream)
        check(rc, "sg_exchange_chain")
        self._exchange_rounds += 1
        if e32.data_ptr() != energies.data_ptr():
            energies.copy_(e32.to(device=energies.device, dtype=energies.dtype))
        return int(n_acc.value)


class GPUMemoryOptimizer:
    """Sizing helper the reference's annealers keep next to the kernel manager
    (reference annealing/cuda_kernels.py:445-580).  Host logic only."""

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.memory_pool = {}

    def get_optimal_batch_size(self, n_spins: int, available_memory: Optional[int] = None) -> int:
        """Replicas (configurations) that fit next to one coupling matrix, capped at 64 like the
        reference (:457-490)."""
        if available_memory is None:
            if torch.cuda.is_available() and self.device.type == "cuda":
                total = torch.cuda.get_device_properties(self.device).total_memory
                available_memory = total - torch.cuda.memory_allocated(self.device)
            else:
                available_memory = 4 * 1024 ** 3
        memory_per_config = (n_spins + n_spins ** 2) * 4 * 2
        usable = int(available_memory * 0.8)
        return min(max(1, usable // memory_per_config), 64)

    def optimize_coupling_matrix_storage(self, couplings: torch.Tensor,
                                         sparsity_threshold: float = 0.1) -> torch.Tensor:
        """Sparse COO when more than ``sparsity_threshold`` of the entries are zero (:518-538)."""
        sparsity = 1.0 - float(torch.count_nonzero(couplings)) / max(1, couplings.numel())
        return couplings.to_sparse_coo() if sparsity > sparsity_threshold else couplings

    def clear_memory_cache(self) -> None:
        if torch.cuda.is_available() and self.device.type == "cuda":
            torch.cuda.empty_cache()
        self.memory_pool.clear()

    def get_memory_stats(self) -> dict:
        stats = {"device": str(self.device), "memory_allocated": 0, "memory_reserved": 0,
                 "max_memory_allocated": 0, "memory_stats": {}}
        if torch.cuda.is_available() and self.device.type == "cuda":
            stats["memory_allocated"] = torch.cuda.memory_allocated(self.device)
            stats["memory_reserved"] = torch.cuda.memory_reserved(self.device)
            stats["max_memory_allocated"] = torch.cuda.max_memory_allocated(self.device)
        return stats
