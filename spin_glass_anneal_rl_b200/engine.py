"""Thin object wrapper over the C ABI (include/sg_b200.h).

One ``Engine`` = one Ising model (J, h) plus R replicas resident on one GPU.
PyTorch is used only as plumbing: device tensors are passed as raw pointers
(``tensor.data_ptr()``), the launch stream is torch's current stream, and
host inputs may be numpy arrays or CPU tensors.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from ._lib import ExchangeParams, SweepParams, WolffParams, check

ArrayLike = Union[np.ndarray, torch.Tensor]


def _as_host(a: ArrayLike, dtype) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


class Engine:
    """Replica-batched annealing engine on one B200 (``device`` = CUDA ordinal)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SGError("no CUDA device: the B200 engine has no CPU fallback")
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        h = ctypes.c_void_p()
        check(self._lib.sg_create(self.device_index, ctypes.byref(h)), "sg_create")
        self._h = h
        self.n = 0
        self.n_replicas = 0
        self._keep = []  # tensors referenced by in-flight launches

    # ------------------------------------------------------------------ plumbing
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ptr(self, t: Optional[torch.Tensor]) -> ctypes.c_void_p:
        if t is None:
            return ctypes.c_void_p(0)
        assert t.is_cuda and t.is_contiguous() and t.device == self.device
        return ctypes.c_void_p(t.data_ptr())

    def _dev(self, a: ArrayLike, dtype: torch.dtype) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=self.device)

    def synchronize(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
        self._keep.clear()

    # ------------------------------------------------------------------ model / replicas
    def set_model(self, J: ArrayLike, h: ArrayLike) -> None:
        """Dense couplings (n x n float32, any symmetry) and external fields."""
        n = int(J.shape[0])
        assert tuple(J.shape) == (n, n) and int(h.shape[0]) == n
        if isinstance(J, torch.Tensor) and J.is_cuda:
            Jd, hd = self._dev(J, torch.float32), self._dev(h, torch.float32)
            check(self._lib.sg_set_model_dense(self._h, n, self._ptr(Jd), n, self._ptr(hd), 1,
                                               self.stream), "sg_set_model_dense")
            self._keep += [Jd, hd]
        else:
            Jh, hh = _as_host(J, np.float32), _as_host(h, np.float32)
            check(self._lib.sg_set_model_dense(self._h, n, Jh.ctypes.data_as(ctypes.c_void_p), n,
                                               hh.ctypes.data_as(ctypes.c_void_p), 0, self.stream),
                  "sg_set_model_dense")
        if n != self.n or getattr(self, "_csr", False) or getattr(self, "_stacked", False):
            self.n_replicas = 0
        self._csr = False
        self._stacked = False
        self.n_models = 1
        self.n = n

    def set_models(self, J: ArrayLike, h: ArrayLike) -> None:
        """Several small dense models at once: J [M, n, n], h [M, n] (n <= 224).  Replicas are
        model-major afterwards (``alloc_replicas(M * r)``: replica k anneals model k // r) and
        every launch covers all models."""
        M, n = int(J.shape[0]), int(J.shape[1])
        assert tuple(J.shape) == (M, n, n) and tuple(h.shape) == (M, n)
        if isinstance(J, torch.Tensor) and J.is_cuda:
            Jd, hd = self._dev(J, torch.float32), self._dev(h, torch.float32)
            check(self._lib.sg_set_model_dense_batch(self._h, M, n, self._ptr(Jd), self._ptr(hd), 1,
                                                     self.stream), "sg_set_model_dense_batch")
            self._keep += [Jd, hd]
        else:
            Jh, hh = _as_host(J, np.float32), _as_host(h, np.float32)
            check(self._lib.sg_set_model_dense_batch(self._h, M, n, Jh.ctypes.data_as(ctypes.c_void_p),
                                                     hh.ctypes.data_as(ctypes.c_void_p), 0, self.stream),
                  "sg_set_model_dense_batch")
        if not (getattr(self, "n_models", 1) == M and self.n == n and getattr(self, "_stacked", False)):
            self.n_replicas = 0            # same-shaped stacks keep their replica buffers
        self._csr = False
        self._stacked = True
        self.n_models = M
        self.n = n

    def set_model_csr(self, rowptr: ArrayLike, colidx: ArrayLike, val: ArrayLike, h: ArrayLike) -> None:
        """Sparse couplings as CSR rows of J (host arrays); any number of spins."""
        rp, ci = _as_host(rowptr, np.int64), _as_host(colidx, np.int32)
        v, hh = _as_host(val, np.float32), _as_host(h, np.float32)
        n = int(hh.shape[0])
        assert rp.shape[0] == n + 1 and ci.shape[0] == v.shape[0] == int(rp[-1])
        check(self._lib.sg_set_model_csr(self._h, n, int(rp[-1]), rp.ctypes.data_as(ctypes.c_void_p),
                                         ci.ctypes.data_as(ctypes.c_void_p),
                                         v.ctypes.data_as(ctypes.c_void_p),
                                         hh.ctypes.data_as(ctypes.c_void_p), self.stream),
              "sg_set_model_csr")
        self.n = n
        self.n_replicas = 0
        self._csr = True
        self._stacked = False
        self.n_models = 1

    def set_model_lattice2d(self, Jx: ArrayLike, Jy: ArrayLike) -> None:
        """2D +-J lattice: Jx[x, y] couples (x, y)-(x+1, y), Jy[x, y] couples (x, y)-(x, y+1);
        entries in {-1, 0, +1} (0 = no bond).  Sweeps use site_order="checkerboard"."""
        jx, jy = _as_host(Jx, np.int8), _as_host(Jy, np.int8)
        L = int(jx.shape[0])
        assert jx.shape == (L, L) and jy.shape == (L, L)
        check(self._lib.sg_set_model_lattice2d(self._h, L, jx.ctypes.data_as(ctypes.c_void_p),
                                               jy.ctypes.data_as(ctypes.c_void_p), self.stream),
              "sg_set_model_lattice2d")
        self.n = L * L
        self.n_replicas = 0
        self._csr = True   # not the dense layout
        self._stacked = False
        self.n_models = 1

    def set_model_groups(self, group_of: ArrayLike, coupling: ArrayLike, h: ArrayLike) -> None:
        """Block-clique couplings: J_ij = coupling[g] for i != j in the same group g."""
        go, cp, hh = _as_host(group_of, np.int32), _as_host(coupling, np.float32), _as_host(h, np.float32)
        check(self._lib.sg_set_model_groups(self._h, int(hh.shape[0]), int(cp.shape[0]),
                                            go.ctypes.data_as(ctypes.c_void_p),
                                            cp.ctypes.data_as(ctypes.c_void_p),
                                            hh.ctypes.data_as(ctypes.c_void_p), self.stream),
              "sg_set_model_groups")
        self.n = int(hh.shape[0])
        self.n_replicas = 0
        self._csr = True   # not the dense layout
        self._stacked = False
        self.n_models = 1

    def alloc_replicas(self, n_replicas: int) -> None:
        check(self._lib.sg_alloc_replicas(self._h, int(n_replicas), self.stream),
              "sg_alloc_replicas")
        self.n_replicas = int(n_replicas)

    def set_spins(self, spins: ArrayLike) -> None:
        """spins[R][n] in {-1,+1} (any integer/float dtype)."""
        assert tuple(spins.shape) == (self.n_replicas, self.n), "spins must be [R, n]"
        if isinstance(spins, torch.Tensor) and spins.is_cuda:
            s = self._dev(spins, torch.int8)
            check(self._lib.sg_set_spins(self._h, self._ptr(s), 1, self.stream), "sg_set_spins")
            self._keep.append(s)
        else:
            s = _as_host(spins, np.int8)
            check(self._lib.sg_set_spins(self._h, s.ctypes.data_as(ctypes.c_void_p), 0,
                                         self.stream), "sg_set_spins")

    def upload_spins_async(self, host_spins: torch.Tensor, slot: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Asynchronous host -> device copy of spins[R][n] (PINNED int8 CPU tensor) into staging
        buffer ``slot`` (0 / 1) on ``stream`` (default: the current stream)."""
        assert host_spins.dtype == torch.int8 and host_spins.is_pinned() and host_spins.is_contiguous()
        assert tuple(host_spins.shape) == (self.n_replicas, self.n)
        st = self.stream if stream is None else ctypes.c_void_p(stream.cuda_stream)
        check(self._lib.sg_upload_spins_async(self._h, ctypes.c_void_p(host_spins.data_ptr()), int(slot), st),
              "sg_upload_spins_async")

    def set_spins_staged(self, slot: int) -> None:
        check(self._lib.sg_set_spins_staged(self._h, int(slot), self.stream), "sg_set_spins_staged")

    def best_config(self, out_energy: Optional[torch.Tensor] = None, out_replica: Optional[torch.Tensor] = None,
                    out_spins: Optional[torch.Tensor] = None):
        """Lowest best-so-far energy, its replica and configuration.  With pinned CPU output tensors
        (float32[1], int32[1], int8[n]) the transfer is asynchronous; without, device tensors are
        returned."""
        host = out_energy is not None and not out_energy.is_cuda
        if out_energy is None:
            out_energy = torch.empty(1, dtype=torch.float32, device=self.device)
            out_replica = torch.empty(1, dtype=torch.int32, device=self.device)
            out_spins = torch.empty(self.n, dtype=torch.int8, device=self.device)
        p = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
        check(self._lib.sg_get_best_config(self._h, p(out_energy), p(out_replica), p(out_spins),
                                           0 if host else 1, self.stream), "sg_get_best_config")
        return out_energy, out_replica, out_spins

    def init_fields(self) -> None:
        check(self._lib.sg_init_fields(self._h, self.stream), "sg_init_fields")

    def refresh_fields(self) -> None:
        """Exact fields / energies from the current spins (best-so-far records untouched)."""
        check(self._lib.sg_refresh_fields(self._h, self.stream), "sg_refresh_fields")

    def reset_best(self) -> None:
        check(self._lib.sg_reset_best(self._h, self.stream), "sg_reset_best")

    # ------------------------------------------------------------------ read-back (device tensors)
    def spins(self) -> torch.Tensor:
        out = torch.empty((self.n_replicas, self.n), dtype=torch.int8, device=self.device)
        check(self._lib.sg_get_spins(self._h, self._ptr(out), 1, self.stream), "sg_get_spins")
        return out

    def energies(self) -> torch.Tensor:
        out = torch.empty(self.n_replicas, dtype=torch.float32, device=self.device)
        check(self._lib.sg_get_energies(self._h, self._ptr(out), 1, self.stream),
              "sg_get_energies")
        return out

    def fields(self) -> torch.Tensor:
        out = torch.empty((self.n_replicas, self.n), dtype=torch.float32, device=self.device)
        check(self._lib.sg_get_fields(self._h, self._ptr(out), 1, self.stream), "sg_get_fields")
        return out

    def accepted(self) -> torch.Tensor:
        out = torch.empty(self.n_replicas, dtype=torch.int64, device=self.device)
        check(self._lib.sg_get_accepted(self._h, self._ptr(out), 1, self.stream),
              "sg_get_accepted")
        return out

    def best(self):
        e = torch.empty(self.n_replicas, dtype=torch.float32, device=self.device)
        s = torch.empty((self.n_replicas, self.n), dtype=torch.int8, device=self.device)
        check(self._lib.sg_get_best(self._h, self._ptr(e), self._ptr(s), 1, self.stream),
              "sg_get_best")
        return e, s

    def best_energies(self) -> torch.Tensor:
        e = torch.empty(self.n_replicas, dtype=torch.float32, device=self.device)
        check(self._lib.sg_get_best(self._h, self._ptr(e), ctypes.c_void_p(0), 1, self.stream),
              "sg_get_best")
        return e

    # ------------------------------------------------------------------ the sweep
    def sweep(self, n_sweeps: int, temps: Optional[ArrayLike] = None, *,
              temps_sweep_stride: int = 0, temps_replica_stride: int = 0,
              rule: str = "metropolis", site_order: str = "random", seed: int = 0,
              sweep_base: int = 0, sites: Optional[ArrayLike] = None,
              sites_block_stride: int = 0, sites_sweep_stride: Optional[int] = None,
              uniforms: Optional[ArrayLike] = None, energy_trace: bool = False,
              track_best: bool = True, replicas_per_block: int = 0, kernel: str = "auto",
              coupling_planes: int = 0, replica_base: int = 0,
              site_energy_changes: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Run ``n_sweeps`` sweeps on every replica (one kernel launch).

        ``site_energy_changes``: optional device float32 [R, n] accumulator of the accepted
        energy changes per site (sequential-FMA kernel only).

        temps: float64 array addressed as temps[s*temps_sweep_stride + r*temps_replica_stride]
        (None = ladder temperatures).  ``uniforms`` switches to injected-uniform mode,
        ``sites`` to an explicit site list.  Returns the [n_sweeps, R] energy trace if asked.
        """
        if rule == "wolff":
            if replicas_per_block or kernel != "auto" or coupling_planes or site_energy_changes is not None:
                raise ValueError("the Wolff cluster move has no kernel / block-shape options")
            return self.sweep_wolff(n_sweeps, temps, temps_sweep_stride=temps_sweep_stride,
                                    temps_replica_stride=temps_replica_stride, site_order=site_order,
                                    seed=seed, sweep_base=sweep_base, sites=sites,
                                    sites_replica_stride=sites_block_stride,
                                    sites_sweep_stride=sites_sweep_stride, uniforms=uniforms,
                                    energy_trace=energy_trace, track_best=track_best,
                                    replica_base=replica_base)
        p = SweepParams()
        p.struct_size = ctypes.sizeof(SweepParams)
        p.n_sweeps = int(n_sweeps)
        p.rule = _lib.SG_RULE[rule]
        p.replicas_per_block = int(replicas_per_block)
        p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        p.sweep_base = int(sweep_base)
        p.track_best = 1 if track_best else 0
        p.kernel = _lib.SG_KERNEL[kernel]
        p.coupling_planes = int(coupling_planes)
        p.replica_base = int(replica_base)
        keep = []
        if temps is not None:
            t = self._dev(temps, torch.float64)
            keep.append(t)
            p.temps = t.data_ptr()
            p.temps_sweep_stride = int(temps_sweep_stride)
            p.temps_replica_stride = int(temps_replica_stride)
        if sites is not None:
            sd = self._dev(sites, torch.int32)
            keep.append(sd)
            p.sites = sd.data_ptr()
            p.site_mode = _lib.SG_SITES["explicit"]
            p.sites_block_stride = int(sites_block_stride)
            p.sites_sweep_stride = int(self.n if sites_sweep_stride is None else sites_sweep_stride)
        else:
            p.site_mode = _lib.SG_SITES[site_order]
        if uniforms is not None:
            u = self._dev(uniforms, torch.float32)
            assert u.numel() == self.n_replicas * n_sweeps * self.n
            keep.append(u)
            p.uniforms = u.data_ptr()
            p.rng_mode = _lib.SG_RNG_INJECTED
        else:
            p.rng_mode = _lib.SG_RNG_PHILOX
        trace = None
        if energy_trace:
            trace = torch.empty((n_sweeps, self.n_replicas), dtype=torch.float32,
                                device=self.device)
            p.energy_trace = trace.data_ptr()
        if site_energy_changes is not None:
            d = site_energy_changes
            if (d.device != self.device or d.dtype != torch.float32 or not d.is_contiguous()
                    or d.numel() != self.n_replicas * self.n):
                raise ValueError("site_energy_changes must be a contiguous device float32 [R, n] tensor")
            keep.append(d)
            p.site_energy_changes = d.data_ptr()
        check(self._lib.sg_sweep(self._h, ctypes.byref(p), self.stream), "sg_sweep")
        self._keep += keep
        if len(self._keep) > 256:
            self.synchronize()
        return trace

    def sweep_wolff(self, n_sweeps: int, temps: Optional[ArrayLike] = None, *,
                    temps_sweep_stride: int = 0, temps_replica_stride: int = 0,
                    site_order: str = "random", seed: int = 0, sweep_base: int = 0,
                    sites: Optional[ArrayLike] = None, sites_replica_stride: int = 0,
                    sites_sweep_stride: Optional[int] = None, uniforms: Optional[ArrayLike] = None,
                    cursor: Optional[torch.Tensor] = None, energy_trace: bool = False,
                    track_best: bool = True, replica_base: int = 0) -> Optional[torch.Tensor]:
        """``n_sweeps`` sweeps of the Wolff cluster move (n cluster updates per sweep and replica),
        exact energies after every sweep.

        ``uniforms`` ([R, m] float32, or [m] for one replica) switches to injected mode: replica r
        consumes its row in order, one value per candidate neighbour, starting at ``cursor[r]``
        (device int64 [R], updated in place; None = start at 0, and ``self.wolff_cursor`` holds the
        positions afterwards).  ``sites``: explicit start sites, addressed as
        sites[r*sites_replica_stride + s*sites_sweep_stride + k]."""
        p = WolffParams()
        p.struct_size = ctypes.sizeof(WolffParams)
        p.n_sweeps = int(n_sweeps)
        p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        p.sweep_base = int(sweep_base)
        p.track_best = 1 if track_best else 0
        p.replica_base = int(replica_base)
        keep = []
        if temps is not None:
            t = self._dev(temps, torch.float64)
            keep.append(t)
            p.temps = t.data_ptr()
            p.temps_sweep_stride = int(temps_sweep_stride)
            p.temps_replica_stride = int(temps_replica_stride)
        if sites is not None:
            sd = self._dev(sites, torch.int32)
            keep.append(sd)
            p.sites = sd.data_ptr()
            p.site_mode = _lib.SG_SITES["explicit"]
            p.sites_replica_stride = int(sites_replica_stride)
            p.sites_sweep_stride = int(self.n if sites_sweep_stride is None else sites_sweep_stride)
        else:
            p.site_mode = _lib.SG_SITES[site_order]
        if uniforms is not None:
            u = self._dev(uniforms, torch.float32).reshape(self.n_replicas, -1)
            keep.append(u)
            p.uniforms = u.data_ptr()
            p.uniforms_replica_stride = int(u.shape[1])
            p.uniforms_per_replica = int(u.shape[1])
            if cursor is None:
                cursor = torch.zeros(self.n_replicas, dtype=torch.int64, device=self.device)
            if cursor.device != self.device or cursor.dtype != torch.int64 or cursor.numel() != self.n_replicas:
                raise ValueError("cursor must be a device int64 [R] tensor")
            self.wolff_cursor = cursor
            p.cursor = cursor.data_ptr()
            p.rng_mode = _lib.SG_RNG_INJECTED
        else:
            p.rng_mode = _lib.SG_RNG_PHILOX
        trace = None
        if energy_trace:
            trace = torch.empty((n_sweeps, self.n_replicas), dtype=torch.float32, device=self.device)
            p.energy_trace = trace.data_ptr()
        check(self._lib.sg_sweep_wolff(self._h, ctypes.byref(p), self.stream), "sg_sweep_wolff")
        self._keep += keep
        if len(self._keep) > 256:
            self.synchronize()
        return trace

    def adaptive_temperature(self, sweep: int, base_temps: torch.Tensor, state: torch.Tensor,
                             temps_out: torch.Tensor, *, accepted_base: int = 0, replica: int = 0,
                             window: int = 100, target_acceptance: float = 0.44,
                             adaptation_rate: float = 0.1, final_temp: float = 0.0) -> None:
        """One ADAPTIVE-schedule step on the device: temps_out[sweep] from the geometric base
        temperature and the running acceptance rate of ``replica`` (no host read-back).
        ``base_temps`` / ``temps_out``: device float64 [n_sweeps]; ``state``: device float64
        [window + 1], zeroed before the first step."""
        for t in (base_temps, state, temps_out):
            if t.device != self.device or t.dtype != torch.float64 or not t.is_contiguous():
                raise ValueError("adaptive_temperature needs contiguous device float64 tensors")
        check(self._lib.sg_adaptive_temperature(
            self._h, int(replica), int(accepted_base), int(sweep), int(window), float(target_acceptance),
            float(adaptation_rate), float(final_temp), self._ptr(base_temps), self._ptr(state),
            self._ptr(temps_out), self.stream), "sg_adaptive_temperature")

    # ------------------------------------------------------------------ parallel tempering
    def set_ladder(self, ladder_temps: Sequence[float], *, n_global: Optional[int] = None,
                   replica_offset: int = 0) -> None:
        """Temperature ladder (rung 0 = hottest).  ``n_global`` / ``replica_offset``: the ladders
        span several engines; this one holds the global replicas [offset, offset + R)."""
        arr = (ctypes.c_double * len(ladder_temps))(*[float(t) for t in ladder_temps])
        ng = self.n_replicas if n_global is None else int(n_global)
        check(self._lib.sg_set_ladder_sharded(self._h, len(ladder_temps), arr, ng, int(replica_offset),
                                              self.stream), "sg_set_ladder")
        self.n_rungs = len(ladder_temps)
        self.n_global = ng
        self.replica_offset = int(replica_offset)

    def exchange(self, parity: int = 0, *, seed: int = 0, round: int = 0,
                 uniforms: Optional[ArrayLike] = None, method: str = "nearest_neighbor",
                 energies_all: Optional[torch.Tensor] = None) -> None:
        """One exchange round.  ``energies_all``: device float32 [n_global] table of every replica's
        energy by global id (sharded ladders: the all-gather of ``energies()`` over the ranks)."""
        p = ExchangeParams()
        p.struct_size = ctypes.sizeof(ExchangeParams)
        p.parity = int(parity)
        p.method = _lib.SG_EXCHANGE[method]
        p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        p.round = int(round)
        if uniforms is not None:
            u = self._dev(uniforms, torch.float64)
            self._keep.append(u)
            p.uniforms = u.data_ptr()
            p.rng_mode = _lib.SG_RNG_INJECTED
        else:
            p.rng_mode = _lib.SG_RNG_PHILOX
        if energies_all is not None:
            ea = energies_all
            if (ea.device != self.device or ea.dtype != torch.float32 or not ea.is_contiguous()
                    or ea.numel() != self.n_global):
                raise ValueError("energies_all must be a contiguous device float32 [n_global] tensor")
            self._keep.append(ea)
            p.energies_all = ea.data_ptr()
        check(self._lib.sg_exchange(self._h, ctypes.byref(p), self.stream), "sg_exchange")

    def check_target(self, target: float, round: int, hit: torch.Tensor, *, best: bool = False) -> None:
        """Asynchronous early-stop test: hit (int32[2], device or PINNED host tensor, hit[0] = -1
        initially) receives (round, replica) the first time min energy <= target."""
        assert hit.dtype == torch.int32 and hit.numel() >= 2 and (hit.is_cuda or hit.is_pinned())
        check(self._lib.sg_check_target(self._h, 1 if best else 0, float(target), int(round),
                                        ctypes.c_void_p(hit.data_ptr()), self.stream), "sg_check_target")

    def ladder_state(self):
        """(replica_at_rung[R], replica_temps[R], attempts[L, K-1], accepts[L, K-1]) on device."""
        R, K = self.n_replicas, self.n_rungs
        Rg = getattr(self, "n_global", R)
        L = Rg // K
        rep_at = torch.empty(Rg, dtype=torch.int32, device=self.device)
        temps = torch.empty(R, dtype=torch.float64, device=self.device)
        att = torch.empty((L, max(K - 1, 1)), dtype=torch.int32, device=self.device)
        acc = torch.empty((L, max(K - 1, 1)), dtype=torch.int32, device=self.device)
        check(self._lib.sg_get_ladder_state(self._h, self._ptr(rep_at), self._ptr(temps),
                                            self._ptr(att), self._ptr(acc), 1, self.stream),
              "sg_get_ladder_state")
        return rep_at, temps, att, acc

    # ------------------------------------------------------------------ checkpoint / resume
    def checkpoint(self) -> dict:
        """Everything a run in progress needs to resume bit for bit (CPU tensors): spins, best
        records, acceptance counters, ladder state.  Fields / energies are recomputed on restore;
        the caller keeps its own sweep counter and seeds (the Philox streams are counter based)."""
        best_e, best_s = self.best()
        ck = {"n": self.n, "n_replicas": self.n_replicas, "spins": self.spins().cpu(),
              "best_energy": best_e.cpu(), "best_spins": best_s.cpu(), "accepted": self.accepted().cpu()}
        if getattr(self, "n_rungs", 0):
            rep_at, _, att, acc = self.ladder_state()
            ck.update(rung_replica=rep_at.cpu(), exchange_attempts=att.cpu(), exchange_accepts=acc.cpu(),
                      n_rungs=self.n_rungs, n_global=getattr(self, "n_global", self.n_replicas),
                      replica_offset=getattr(self, "replica_offset", 0))
        return ck

    def restore(self, ck: dict, ladder_temps: Optional[Sequence[float]] = None) -> None:
        """Inverse of ``checkpoint`` on an engine that holds the same model (``set_model`` done)."""
        assert ck["n"] == self.n, "checkpoint belongs to a model of another size"
        if self.n_replicas != ck["n_replicas"]:
            self.alloc_replicas(ck["n_replicas"])
        self.set_spins(ck["spins"])
        self.init_fields()
        be = ck["best_energy"].to(self.device, torch.float32).contiguous()
        bs = ck["best_spins"].to(self.device, torch.int8).contiguous()
        check(self._lib.sg_set_best(self._h, self._ptr(be), self._ptr(bs), 1, self.stream), "sg_set_best")
        ac = ck["accepted"].to(self.device, torch.int64).contiguous()
        check(self._lib.sg_set_accepted(self._h, self._ptr(ac), 1, self.stream), "sg_set_accepted")
        self._keep += [be, bs, ac]
        if "rung_replica" in ck:
            if ladder_temps is None:
                raise ValueError("the checkpoint holds a ladder state: pass the ladder temperatures")
            self.set_ladder(ladder_temps, n_global=ck["n_global"], replica_offset=ck["replica_offset"])
            ra = ck["rung_replica"].to(self.device, torch.int32).contiguous()
            at = ck["exchange_attempts"].to(self.device, torch.int32).contiguous()
            ac2 = ck["exchange_accepts"].to(self.device, torch.int32).contiguous()
            check(self._lib.sg_set_ladder_state(self._h, self._ptr(ra), self._ptr(at), self._ptr(ac2), 1,
                                                self.stream), "sg_set_ladder_state")
            self._keep += [ra, at, ac2]

    # ------------------------------------------------------------------ batched energies
    def batch_energies(self, spins: ArrayLike, want_fields: bool = False):
        """Energies (and local fields) of arbitrary configurations spins[B][n]."""
        B = int(spins.shape[0])
        s = self._dev(spins, torch.int8)
        e = torch.empty(B, dtype=torch.float32, device=self.device)
        f = torch.empty((B, self.n), dtype=torch.float32, device=self.device) if want_fields else None
        check(self._lib.sg_batch_energies(self._h, B, self._ptr(s), self._ptr(e), self._ptr(f), 1,
                                          self.stream), "sg_batch_energies")
        return (e, f) if want_fields else e

    # ------------------------------------------------------------------ facts
    def query(self) -> dict:
        v = [ctypes.c_int32() for _ in range(5)]
        check(self._lib.sg_query(self._h, *[ctypes.byref(x) for x in v]), "sg_query")
        return dict(n=v[0].value, n_pad=v[1].value, n_replicas=v[2].value,
                    max_replicas_per_block=v[3].value, sm_count=v[4].value)

    def tc_cluster_size(self) -> int:
        """CTAs per replica group of the tensor-core sweep (2 = cluster pairs, 1, or 0 = n/a)."""
        return int(self._lib.sg_tc_cluster_size(self._h))

    def tc_side_replicas(self, n_sweeps: int, coupling_planes: int = 0) -> int:
        """Replicas a tensor-core launch of ``n_sweeps`` sweeps runs as cluster pairs on the SMs the
        clusters of 4 leave idle (0 = none)."""
        return int(self._lib.sg_tc_side_replicas(self._h, int(n_sweeps), int(coupling_planes)))

    def set_profiling(self, enable: bool) -> None:
        check(self._lib.sg_set_profiling(self._h, 1 if enable else 0), "sg_set_profiling")

    def profile(self) -> dict:
        """Device time of the sweep / gather kernels since the last call (CUDA events)."""
        sm, gm = ctypes.c_double(), ctypes.c_double()
        sc, gc = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._lib.sg_get_profile(self._h, ctypes.byref(sm), ctypes.byref(sc), ctypes.byref(gm),
                                       ctypes.byref(gc)), "sg_get_profile")
        return dict(sweep_ms=sm.value, sweep_launches=sc.value, gather_ms=gm.value,
                    gather_launches=gc.value)

    def launch_count(self) -> int:
        return int(self._lib.sg_launch_count(self._h))

    def measure_stream_bandwidth(self, nbytes: int, iters: int = 20, stagger: bool = False) -> float:
        out = ctypes.c_double()
        check(self._lib.sg_measure_stream_bandwidth(self._h, int(nbytes), int(iters),
                                                    1 if stagger else 0, ctypes.byref(out)),
              "sg_measure_stream_bandwidth")
        return out.value

    def measure_tma_stream(self, nbytes: int, row_bytes: int, depth: int = 8, n_rows: int = 4096,
                           stagger: bool = False) -> float:
        out = ctypes.c_double()
        check(self._lib.sg_measure_tma_stream(self._h, int(nbytes), int(row_bytes), int(depth),
                                              int(n_rows), 1 if stagger else 0, ctypes.byref(out)),
              "sg_measure_tma_stream")
        return out.value
