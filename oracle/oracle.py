"""CPU oracle: restatement of the reference's annealing path (host-side half).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (cpu_baseline / ``--impl reference``) may import this module, and
only as the checker or the timed CPU baseline.  The product package
``spin_glass_anneal_rl_b200`` never imports it.

Parity status: PINNED against traces recorded from the reference itself
(``tests/golden/make_golden.py`` / ``make_wolff_golden.py`` / ``make_operator_golden.py`` ->
``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).

The sweeps run in C (``sg_oracle.c``); this file restates the control flow
around them, citing the reference lines (paths relative to
``/root/reference/spin_glass_rl/``):

* ``wolff_sweeps*``         <- core/spin_dynamics.py:193-262 (UpdateRule.WOLFF, dense branch), pinned
  by ``tests/golden/wolff_*.npz`` (``tests/golden/make_wolff_golden.py``)
* ``schedule_temperature``  <- annealing/temperature_scheduler.py:69-269
* ``anneal``                <- annealing/gpu_annealer.py:96-183, 254-269
* ``temperature_ladder``    <- annealing/parallel_tempering.py:146-173
* ``parallel_tempering``    <- annealing/parallel_tempering.py:82-144, 175-258, 295-313
* ``result_postprocess``    <- annealing/result.py:37-77
* ``operator_*``            <- annealing/cuda_kernels.py:371-436 (the loops behind the
  CUDAKernelManager entry points), pinned by ``tests/golden/op_*.npz``
  (``tests/golden/make_operator_golden.py``)
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsg_oracle.so")
_lib = None

RULES = {"metropolis": 0, "glauber": 1, "heat_bath": 2}


def build(force: bool = False) -> str:
    """Compile sg_oracle.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "sg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsg_oracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_fp = ctypes.POINTER(ctypes.c_float)
        c_dp = ctypes.POINTER(ctypes.c_double)
        c_i64p = ctypes.POINTER(ctypes.c_int64)
        c_i32p = ctypes.POINTER(ctypes.c_int32)
        c_u32p = ctypes.POINTER(ctypes.c_uint32)
        L.sgo_local_field.restype = ctypes.c_double
        L.sgo_local_field.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int, ctypes.c_int]
        L.sgo_energy.restype = ctypes.c_double
        L.sgo_energy.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int, c_fp]
        L.sgo_sweeps.restype = ctypes.c_int
        L.sgo_sweeps.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int, ctypes.c_int, c_dp,
                                 ctypes.c_int, c_u32p, ctypes.c_int64, c_i64p, c_dp, c_i64p, c_i32p,
                                 c_fp, ctypes.c_int]
        L.sgo_sweeps_scheduled.restype = ctypes.c_int
        L.sgo_sweeps_scheduled.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int,
                                           ctypes.c_int, c_dp, ctypes.c_int, c_i32p, c_fp, c_dp,
                                           c_i64p]
        L.sgo_batch_fields_energies.restype = None
        L.sgo_batch_fields_energies.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int,
                                                ctypes.c_int, c_dp, c_dp]
        L.sgo_baseline_run.restype = ctypes.c_int64
        L.sgo_baseline_run.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_double, ctypes.c_uint64, ctypes.c_int,
                                       c_dp]
        L.sgo_num_threads.restype = ctypes.c_int
        L.sgo_wolff_sweeps.restype = ctypes.c_int
        L.sgo_wolff_sweeps.argtypes = [c_fp, ctypes.c_int64, c_fp, c_fp, ctypes.c_int, c_dp, ctypes.c_int,
                                       c_u32p, ctypes.c_int64, c_i64p, c_i32p, c_fp, ctypes.c_int64,
                                       c_dp, c_i64p, c_i32p, c_fp, ctypes.c_int64, c_i32p, c_i64p]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray], ctype):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ctype))
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- RNG streams
def mt_raw_stream(seed: int, n: int, skip: int = 0) -> np.ndarray:
    """Raw 32-bit mt19937 outputs for ``torch.manual_seed(seed)`` /
    ``np.random.seed(seed)`` (both use init_genrand; verified bit-identical)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed))
    raw = bg.random_raw(int(skip) + int(n))
    return np.ascontiguousarray(raw[skip:].astype(np.uint32))


def raw_to_spins(raw: np.ndarray) -> np.ndarray:
    """(torch.randint(0, 2, (n,)) * 2 - 1).float()  -- core/ising_model.py:67,202."""
    return ((raw % 2).astype(np.float32) * 2.0 - 1.0).astype(np.float32)


# --------------------------------------------------------------------------- model-level
def local_field(J, h, spins, i: int) -> float:
    J, h, spins = _f32(J), _f32(h), _f32(spins)
    n = spins.shape[0]
    return float(lib().sgo_local_field(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                                       _p(spins, ctypes.c_float), n, int(i)))


def energy(J, h, spins) -> float:
    J, h, spins = _f32(J), _f32(h), _f32(spins)
    n = spins.shape[0]
    scratch = np.empty(n, dtype=np.float32)
    return float(lib().sgo_energy(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                                  _p(spins, ctypes.c_float), n, _p(scratch, ctypes.c_float)))


def batch_fields_energies(J, h, S):
    J, h, S = _f32(J), _f32(h), _f32(S)
    b, n = S.shape
    F = np.empty((b, n), dtype=np.float64)
    E = np.empty(b, dtype=np.float64)
    lib().sgo_batch_fields_energies(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                                    _p(S, ctypes.c_float), n, b, _p(F, ctypes.c_double),
                                    _p(E, ctypes.c_double))
    return F, E


class RawStream:
    """Cursor over a raw mt19937 stream (one per global generator)."""

    def __init__(self, raw: np.ndarray, pos: int = 0):
        self.raw = np.ascontiguousarray(raw, dtype=np.uint32)
        self.pos = int(pos)

    def take(self, n: int) -> np.ndarray:
        out = self.raw[self.pos:self.pos + n]
        if out.shape[0] != n:
            raise RuntimeError("raw stream exhausted")
        self.pos += n
        return out


def sweeps(J, h, spins: np.ndarray, temps, rule: str, stream: RawStream, trace: bool = False,
           redo_flip_dot: bool = False):
    """n = len(temps) reference sweeps on ONE replica, in place on ``spins``.

    Returns (energies[n], accepted[n], trace_site or None, trace_u or None)."""
    J, h = _f32(J), _f32(h)
    assert spins.dtype == np.float32 and spins.flags.c_contiguous
    n = spins.shape[0]
    temps = np.ascontiguousarray(np.asarray(temps, dtype=np.float64))
    ns = temps.shape[0]
    energies = np.empty(ns, dtype=np.float64)
    accepted = np.empty(ns, dtype=np.int64)
    tsite = np.empty(ns * n, dtype=np.int32) if trace else None
    tu = np.empty(ns * n, dtype=np.float32) if trace else None
    pos = ctypes.c_int64(stream.pos)
    rc = lib().sgo_sweeps(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                          _p(spins, ctypes.c_float), n, RULES[rule], _p(temps, ctypes.c_double), ns,
                          _p(stream.raw, ctypes.c_uint32), stream.raw.shape[0], ctypes.byref(pos),
                          _p(energies, ctypes.c_double), _p(accepted, ctypes.c_int64),
                          _p(tsite, ctypes.c_int32), _p(tu, ctypes.c_float), int(redo_flip_dot))
    if rc != 0:
        raise RuntimeError("raw stream exhausted inside sgo_sweeps")
    stream.pos = pos.value
    return energies, accepted, tsite, tu


def sweeps_scheduled(J, h, spins: np.ndarray, temps, rule: str, sites, uniforms):
    """Sweeps driven by explicit (site, uniform) per attempt; in place on spins."""
    J, h = _f32(J), _f32(h)
    assert spins.dtype == np.float32 and spins.flags.c_contiguous
    n = spins.shape[0]
    temps = np.ascontiguousarray(np.asarray(temps, dtype=np.float64))
    ns = temps.shape[0]
    sites = np.ascontiguousarray(np.asarray(sites, dtype=np.int32).reshape(-1))
    uniforms = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float32).reshape(-1))
    assert sites.shape[0] == ns * n and uniforms.shape[0] == ns * n
    energies = np.empty(ns, dtype=np.float64)
    accepted = np.empty(ns, dtype=np.int64)
    lib().sgo_sweeps_scheduled(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                               _p(spins, ctypes.c_float), n, RULES[rule],
                               _p(temps, ctypes.c_double), ns, _p(sites, ctypes.c_int32),
                               _p(uniforms, ctypes.c_float), _p(energies, ctypes.c_double),
                               _p(accepted, ctypes.c_int64))
    return energies, accepted


# --------------------------------------------------------------------------- Wolff cluster updates
def wolff_sweeps(J, h, spins: np.ndarray, temps, stream: RawStream, trace: bool = False):
    """n = len(temps) reference sweeps with UpdateRule.WOLFF on ONE replica, in place on ``spins``
    (core/spin_dynamics.py:73-94 with _wolff_cluster_dense, :211-262): every sweep is n cluster
    updates, each from a start site drawn from the stream, the uniforms of the cluster growth
    drawn from the same stream as they are needed.

    Returns (energies[n], cluster_flips[n], trace) with trace = None or a dict: ``sites``
    [n_sweeps, n] start sites, ``uniforms`` every uniform consumed in order, ``draws``
    [n_sweeps, n] how many each update consumed."""
    J, h = _f32(J), _f32(h)
    assert spins.dtype == np.float32 and spins.flags.c_contiguous
    n = spins.shape[0]
    temps = np.ascontiguousarray(np.asarray(temps, dtype=np.float64))
    ns = temps.shape[0]
    energies = np.empty(ns, dtype=np.float64)
    flips = np.empty(ns, dtype=np.int64)
    cap = ns * n * n * n if trace else 0   # an update draws at most one uniform per (visited, other) pair
    cap = min(cap, stream.raw.shape[0] - stream.pos)
    tsite = np.empty(ns * n, dtype=np.int32) if trace else None
    tu = np.empty(max(cap, 1), dtype=np.float32) if trace else None
    draws = np.empty(ns * n, dtype=np.int32) if trace else None
    pos = ctypes.c_int64(stream.pos)
    used = ctypes.c_int64(0)
    rc = lib().sgo_wolff_sweeps(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                                _p(spins, ctypes.c_float), n, _p(temps, ctypes.c_double), ns,
                                _p(stream.raw, ctypes.c_uint32), stream.raw.shape[0], ctypes.byref(pos),
                                _p(None, ctypes.c_int32), _p(None, ctypes.c_float), 0,
                                _p(energies, ctypes.c_double), _p(flips, ctypes.c_int64),
                                _p(tsite, ctypes.c_int32), _p(tu, ctypes.c_float), cap,
                                _p(draws, ctypes.c_int32), ctypes.byref(used))
    if rc != 0:
        raise RuntimeError("raw stream exhausted inside sgo_wolff_sweeps")
    stream.pos = pos.value
    tr = None
    if trace:
        tr = {"sites": tsite.reshape(ns, n), "uniforms": tu[:used.value].copy(),
              "draws": draws.reshape(ns, n)}
    return energies, flips, tr


def wolff_sweeps_scheduled(J, h, spins: np.ndarray, temps, sites, uniforms):
    """The same sweeps driven by explicit start sites [n_sweeps, n] and a list of uniforms consumed
    in order (the form the CUDA kernel takes in injected mode).  In place on ``spins``; returns
    (energies, cluster_flips, uniforms_used)."""
    J, h = _f32(J), _f32(h)
    assert spins.dtype == np.float32 and spins.flags.c_contiguous
    n = spins.shape[0]
    temps = np.ascontiguousarray(np.asarray(temps, dtype=np.float64))
    ns = temps.shape[0]
    sites = np.ascontiguousarray(np.asarray(sites, dtype=np.int32).reshape(-1))
    uniforms = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float32).reshape(-1))
    assert sites.shape[0] == ns * n
    energies = np.empty(ns, dtype=np.float64)
    flips = np.empty(ns, dtype=np.int64)
    used = ctypes.c_int64(0)
    rc = lib().sgo_wolff_sweeps(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                                _p(spins, ctypes.c_float), n, _p(temps, ctypes.c_double), ns,
                                _p(None, ctypes.c_uint32), 0, ctypes.cast(None, ctypes.POINTER(ctypes.c_int64)),
                                _p(sites, ctypes.c_int32), _p(uniforms, ctypes.c_float),
                                uniforms.shape[0], _p(energies, ctypes.c_double),
                                _p(flips, ctypes.c_int64), _p(None, ctypes.c_int32),
                                _p(None, ctypes.c_float), 0, _p(None, ctypes.c_int32),
                                ctypes.byref(used))
    if rc != 0:
        raise RuntimeError("uniform list exhausted inside sgo_wolff_sweeps")
    return energies, flips, int(used.value)


# --------------------------------------------------------------------------- operator API
def operator_metropolis_update(J, h, spins, temperature: float, n_updates: int, *,
                               uniforms=None, uniform_stream=None):
    """CUDAKernelManager._metropolis_update_fallback (annealing/cuda_kernels.py:371-397), float32.

    ``n_updates`` passes over the sites in index order; local field WITHOUT the diagonal term;
    dE = 2 s_i lf; accept iff dE <= 0 or u < exp(-dE / T).  The reference draws a uniform only
    when dE > 0: ``uniform_stream`` (1-D) is consumed like that, in order; ``uniforms``
    ([n_updates, n]) gives attempt (p, i) its own value instead (what the CUDA path takes).
    Returns (spins, accepted, energy_changes[n], positional uniforms [n_updates, n] -- the stream
    values placed at the attempts that consumed them, 0.5 elsewhere)."""
    J = _f32(J)
    h = _f32(h)
    s = np.array(spins, dtype=np.float32)
    n = s.shape[0]
    T = np.float32(temperature)
    changes = np.zeros(n, dtype=np.float32)
    placed = np.full((n_updates, n), 0.5, dtype=np.float32)
    accepted = 0
    pos = 0
    for p in range(n_updates):
        for i in range(n):
            lf = np.float32(h[i] + np.sum(J[i] * s, dtype=np.float32) - J[i, i] * s[i])
            dE = np.float32(np.float32(2.0) * s[i] * lf)
            take = dE <= 0
            if not take:
                if uniforms is not None:
                    u = np.float32(uniforms[p][i])
                else:
                    u = np.float32(uniform_stream[pos])
                    pos += 1
                placed[p, i] = u
                take = float(u) < float(np.exp(np.float32(-dE / T)))
            if take:
                s[i] = -s[i]
                changes[i] += dE
                accepted += 1
    return s, accepted, changes, placed


def operator_energy(J, h, spins) -> float:
    """CUDAKernelManager._compute_energy_fallback (annealing/cuda_kernels.py:398-403):
    -1/2 sum_ij s_i J_ij s_j - sum_i h_i s_i, diagonal included (accumulated in double here)."""
    J = np.asarray(J, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    s = np.asarray(spins, dtype=np.float64)
    return float(-0.5 * s @ J @ s - h @ s)


def operator_exchange(spins_arrays, energies, temperatures, uniforms):
    """CUDAKernelManager._parallel_tempering_fallback (annealing/cuda_kernels.py:405-436), float32:
    one ordered pass over (i, i+1); p = exp((1/T[i+1] - 1/T[i]) * (E[i] - E[i+1])); a uniform per
    pair; rows and energies swapped in place.  Returns (rows, energies, accepted)."""
    S = np.array(spins_arrays)
    E = np.array(energies, dtype=np.float32)
    T = np.asarray(temperatures, dtype=np.float32)
    accepted = 0
    for i in range(S.shape[0] - 1):
        beta1 = np.float32(1.0) / T[i]
        beta2 = np.float32(1.0) / T[i + 1]
        prob = np.exp(np.float32((beta2 - beta1) * (E[i] - E[i + 1])))
        if float(np.float32(uniforms[i])) < float(prob):
            S[[i, i + 1]] = S[[i + 1, i]]
            E[[i, i + 1]] = E[[i + 1, i]]
            accepted += 1
    return S, E, accepted


# --------------------------------------------------------------------------- schedules
def schedule_temperature(kind: str, sweep: int, T0: float, Tf: float, total: int, *,
                         alpha: float = 0.95, k: float = 1.0, c: float = 1.0) -> float:
    """T(sweep) for the stateless schedules, annealing/temperature_scheduler.py:69-203."""
    if kind == "linear":  # :72-82
        if sweep >= total:
            return Tf
        return max(T0 - (T0 - Tf) * (sweep / total), Tf)
    if kind == "exponential":  # :95-106
        lam = -np.log(Tf / T0) / total if Tf > 0 else 0.01
        return max(T0 * np.exp(-lam * sweep), Tf)
    if kind == "geometric":  # :119-122 (ignores total)
        return max(T0 * (alpha ** sweep), Tf)
    if kind == "logarithmic":  # :135-142
        if sweep == 0:
            return T0
        t = c / np.log(1 + sweep)
        return max(t * T0 / c, Tf)
    if kind == "power_law":  # :155-158
        return max(T0 / ((1 + sweep) ** k), Tf)
    if kind == "fast":  # :171-177
        if sweep == 0:
            return T0
        return max(T0 / sweep, Tf)
    if kind == "boltzmann":  # :190-196
        if sweep == 0:
            return T0
        return max(T0 / np.log(1 + sweep), Tf)
    raise ValueError(kind)


class AdaptiveState:
    """AdaptiveSchedule.update, annealing/temperature_scheduler.py:206-249."""

    def __init__(self, T0, Tf, alpha=0.95, target_acceptance=0.44, adaptation_window=100,
                 adaptation_rate=0.1):
        self.T0, self.Tf, self.alpha = T0, Tf, alpha
        self.target, self.window, self.rate = target_acceptance, adaptation_window, adaptation_rate
        self.hist: List[float] = []

    def update(self, sweep: int, acceptance_rate: float) -> float:
        self.hist.append(acceptance_rate)
        base = max(self.T0 * (self.alpha ** sweep), self.Tf)
        if len(self.hist) >= self.window:
            recent = np.mean(self.hist[-self.window:])
            if recent > self.target:
                adj = 1.0 - self.rate
            elif recent < self.target:
                adj = 1.0 + self.rate
            else:
                adj = 1.0
            return max(base * adj, self.Tf)
        return base


# --------------------------------------------------------------------------- annealer
@dataclass
class OracleResult:
    best_configuration: np.ndarray
    best_energy: float
    energy_history: List[float]
    temperature_history: List[float]
    acceptance_rate_history: List[float]
    n_sweeps: int
    final_spins: np.ndarray
    sweep_energies: List[float] = field(default_factory=list)
    raw_consumed: int = 0
    extra: Dict = field(default_factory=dict)


def _converged(energy_history: List[float], tol: float) -> bool:
    """GPUAnnealer._check_convergence, annealing/gpu_annealer.py:254-269."""
    if len(energy_history) < 50:
        return False
    recent = energy_history[-20:]
    std, mean = np.std(recent), np.mean(recent)
    if abs(mean) > 0:
        return (std / abs(mean)) < tol
    return std < tol


def anneal(J, h, spins0, *, n_sweeps: int, T0: float, Tf: float, schedule: str = "geometric",
           schedule_params: Optional[dict] = None, record_interval: int = 10,
           energy_tolerance: float = 1e-8, rule: str = "metropolis",
           stream: RawStream, trace: bool = False) -> OracleResult:
    """GPUAnnealer.anneal on the CPU path, annealing/gpu_annealer.py:96-183.

    trace=True also returns, in ``extra``, the temperature of every sweep and the
    (site, uniform) of every attempt -- the reference's stream under its own update
    order, which is what the CUDA kernel is fed in injected mode."""
    sp = dict(schedule_params or {"alpha": 0.95})
    spins = np.ascontiguousarray(np.asarray(spins0, dtype=np.float32)).copy()
    start_pos = stream.pos
    adaptive = AdaptiveState(T0, Tf, **sp) if schedule == "adaptive" else None
    best_e = energy(J, h, spins)  # :130
    best_cfg = spins.copy()  # :131
    e_hist, t_hist, a_hist = [best_e], [T0], [0.0]  # :134-136
    n_acc = n_rej = 0
    sweep_energies = []
    tr_T, tr_site, tr_u = [], [], []
    sweep = -1
    for sweep in range(n_sweeps):  # :139
        rate = n_acc / (n_acc + n_rej) if (n_acc + n_rej) else 0.0
        if adaptive is not None:
            T = adaptive.update(sweep, rate)  # :141
        else:
            T = schedule_temperature(schedule, sweep, T0, Tf, n_sweeps, **sp)
        T = max(float(T), 1e-10)  # set_temperature clamp, core/spin_dynamics.py:57-59
        if rule == "wolff":   # every cluster site counts as accepted, nothing is ever rejected
            es, acc, wtr = wolff_sweeps(J, h, spins, [T], stream, trace=trace)
            if trace:
                tr_T.append(T)
                tr_site.append(wtr["sites"][0])
                tr_u.append(wtr["uniforms"])
            n_acc += int(acc[0])
        else:
            es, acc, ts, tu = sweeps(J, h, spins, [T], rule, stream, trace=trace)
            if trace:
                tr_T.append(T)
                tr_site.append(ts)
                tr_u.append(tu)
            n_acc += int(acc[0])
            n_rej += spins.shape[0] - int(acc[0])
        cur = float(es[0])
        sweep_energies.append(cur)
        if cur < best_e:  # :151-153
            best_e, best_cfg = cur, spins.copy()
        if sweep % record_interval == 0:  # :156-164
            e_hist.append(cur)
            t_hist.append(float(T))
            a_hist.append(n_acc / (n_acc + n_rej))
            if _converged(e_hist, energy_tolerance):
                break
    res = OracleResult(best_cfg, best_e, e_hist, t_hist, a_hist, sweep + 1, spins, sweep_energies,
                       stream.pos - start_pos)
    if trace:
        res.extra = {"temps": np.array(tr_T, np.float64), "sites": np.stack(tr_site),
                     "uniforms": np.concatenate(tr_u) if rule == "wolff" else np.stack(tr_u)}
        if rule == "wolff":   # the uniforms of sweep s are uniforms[offsets[s]:offsets[s + 1]]
            res.extra["uniform_offsets"] = np.concatenate([[0], np.cumsum([len(u) for u in tr_u])])
    return res


# --------------------------------------------------------------------------- parallel tempering
def temperature_ladder(n: int, tmin: float, tmax: float, dist: str = "geometric") -> List[float]:
    """ParallelTempering._generate_temperature_ladder, parallel_tempering.py:146-173."""
    if dist == "geometric":
        ratio = tmin / tmax
        return [tmax * (ratio ** (i / (n - 1))) for i in range(n)]
    if dist == "linear":
        return np.linspace(tmax, tmin, n).tolist()
    if dist == "exponential":
        return np.logspace(np.log10(tmax), np.log10(tmin), n).tolist()
    raise ValueError(dist)


def parallel_tempering(J, h, *, n_replicas: int, n_sweeps: int, temp_min: float, temp_max: float,
                       temp_distribution: str = "geometric", exchange_interval: int = 10,
                       record_interval: int = 10, rule: str = "metropolis", stream: RawStream,
                       np_rng: np.random.RandomState, exchange_method: str = "nearest_neighbor",
                       trace: bool = False) -> OracleResult:
    """ParallelTempering.run with n_threads=1 on device='cpu'.

    parallel_tempering.py:82-144 (loop), :175-189 (replica init: model.copy()
    draws n randints that are discarded, reset_to_random() draws n more),
    :214-220 (nearest-neighbour exchange), :222-232 (all_pairs, CPU branch: every pair i < j
    is attempted with probability 0.1, statistics filed under min(i, j)), :234-258 (one
    exchange), :295-313 (statistics / best).

    ``trace=True`` also records what a replay needs: the initial spins per temperature slot,
    the (site, uniform) of every attempt per (sweep, slot) and the numpy draws of every
    exchange round (res.extra["spins0" / "sites" / "uniforms" / "exchange_draws"])."""
    J, h = _f32(J), _f32(h)
    n = J.shape[0]
    temps = temperature_ladder(n_replicas, temp_min, temp_max, temp_distribution)
    start_pos = stream.pos
    reps = []
    for _ in range(n_replicas):  # :180-189
        stream.take(n)  # IsingModel(config) inside copy(): spins overwritten
        reps.append(raw_to_spins(stream.take(n)).copy())  # reset_to_random()
    spins0 = np.stack(reps).copy()
    n_acc = [0] * n_replicas
    n_tot = [0] * n_replicas
    attempts = np.zeros(n_replicas - 1)
    accepts = np.zeros(n_replicas - 1)
    e_hists: List[List[float]] = [[] for _ in range(n_replicas)]
    best_e, best_cfg = float("inf"), None
    tr_sites = np.zeros((n_sweeps, n_replicas, n), np.int32) if trace else None
    tr_uni = np.zeros((n_sweeps, n_replicas, n), np.float32) if trace else None
    draws: List[List[float]] = []

    def attempt(i, j, log):  # :234-258
        bi, bj = 1.0 / temps[i], 1.0 / temps[j]
        ei, ej = energy(J, h, reps[i]), energy(J, h, reps[j])
        prob = min(1.0, np.exp((bj - bi) * (ej - ei)))  # :244-246
        attempts[min(i, j)] += 1
        u = np_rng.rand()  # :252 (always drawn)
        log.append(float(u))
        if u < prob:
            reps[i], reps[j] = reps[j], reps[i]  # :254-256 swap configurations
            accepts[min(i, j)] += 1

    tr_wolff: List[List[np.ndarray]] = []   # rule "wolff": the uniforms of (sweep, slot), in order
    for sweep in range(n_sweeps):  # :108
        if rule == "wolff":
            tr_wolff.append([])
        for r in range(n_replicas):  # :193-196 (sequential, shared global stream)
            if rule == "wolff":   # cluster sizes count as accepted, nothing as rejected
                _, acc, wtr = wolff_sweeps(J, h, reps[r], [max(temps[r], 1e-10)], stream, trace=trace)
                if trace:
                    tr_sites[sweep, r] = wtr["sites"][0]
                    tr_wolff[-1].append(wtr["uniforms"])
                n_acc[r] += int(acc[0])
                n_tot[r] += int(acc[0])
                continue
            _, acc, ts, tu = sweeps(J, h, reps[r], [max(temps[r], 1e-10)], rule, stream, trace=trace)
            if trace:
                tr_sites[sweep, r], tr_uni[sweep, r] = ts, tu
            n_acc[r] += int(acc[0])
            n_tot[r] += n
        if sweep % exchange_interval == 0 and sweep > 0:  # :113-114
            log: List[float] = []
            if exchange_method == "nearest_neighbor":
                start = int(np_rng.randint(0, 2))  # :217
                log.append(float(start))
                for i in range(start, n_replicas - 1, 2):  # :219-220
                    attempt(i, i + 1, log)
            elif exchange_method == "all_pairs":  # :228-232
                for i in range(n_replicas - 1):
                    for j in range(i + 1, n_replicas):
                        pick = np_rng.rand()
                        log.append(float(pick))
                        if pick < 0.1:
                            attempt(i, j, log)
            else:
                raise ValueError(exchange_method)
            draws.append(log)
        if sweep % record_interval == 0:  # :117-125
            es = [energy(J, h, s) for s in reps]
            for r in range(n_replicas):
                e_hists[r].append(es[r])
            cur_best = min(range(n_replicas), key=lambda r: (es[r], r))
            if es[cur_best] < best_e:
                best_e, best_cfg = es[cur_best], reps[cur_best].copy()
    res = OracleResult(best_cfg, best_e, e_hists[0], [temps[0]] * len(e_hists[0]),
                       [n_acc[r] / n_tot[r] if n_tot[r] else 0.0 for r in range(n_replicas)],
                       n_sweeps, np.stack(reps), [], stream.pos - start_pos)
    res.extra = {"temperatures": temps, "exchange_attempts": attempts, "exchange_accepts": accepts,
                 "energy_histories": e_hists, "spins0": spins0, "sites": tr_sites,
                 "uniforms": tr_wolff if rule == "wolff" else tr_uni, "exchange_draws": draws}
    return res


# --------------------------------------------------------------------------- result post-processing
def result_postprocess(energy_history: List[float], best_energy: float):
    """AnnealingResult.__post_init__ derived fields, annealing/result.py:55-71."""
    energy_std = float(np.std(energy_history)) if energy_history else 0.0
    conv = None
    if len(energy_history) > 10:
        w = min(50, len(energy_history) // 4)
        e = np.array(energy_history)
        for i in range(w, len(e)):
            if np.std(e[i - w:i]) < 0.01 * abs(best_energy):
                conv = i - w
                break
    return energy_std, conv


# --------------------------------------------------------------------------- CPU baseline timing
def baseline_run(J, h, spins: np.ndarray, n_sweeps: int, T: float, seed: int = 1,
                 n_threads: int = 0):
    """Reference path (per-attempt dot products + per-sweep O(N^2) energy) for
    spins[R, N] independent replicas over OpenMP threads.  Returns
    (attempts, energies[R]); the caller times it."""
    J, h = _f32(J), _f32(h)
    assert spins.dtype == np.float32 and spins.flags.c_contiguous
    R, n = spins.shape
    E = np.empty(R, dtype=np.float64)
    att = lib().sgo_baseline_run(_p(J, ctypes.c_float), J.shape[1], _p(h, ctypes.c_float),
                                 _p(spins, ctypes.c_float), n, R, int(n_sweeps), float(T),
                                 int(seed), int(n_threads), _p(E, ctypes.c_double))
    return int(att), E


def baseline_run_csr(rowptr, colidx, val, h, spins: np.ndarray, n_sweeps: int, T: float, seed: int = 1,
                     n_threads: int = 0):
    """The same baseline on CSR rows of J (models that cannot be densified at full size: cfg2,
    cfg5); timing arm only.  Returns (attempts, energies[R])."""
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    v, h = _f32(val), _f32(h)
    assert spins.dtype == np.float32 and spins.flags.c_contiguous
    R, n = spins.shape
    E = np.empty(R, dtype=np.float64)
    fn = lib().sgo_baseline_run_csr
    fn.restype = ctypes.c_int64
    att = fn(_p(rp, ctypes.c_int64), _p(ci, ctypes.c_int32), _p(v, ctypes.c_float), _p(h, ctypes.c_float),
             _p(spins, ctypes.c_float), ctypes.c_int(n), ctypes.c_int(R), ctypes.c_int(int(n_sweeps)),
             ctypes.c_double(float(T)), ctypes.c_uint64(int(seed)), ctypes.c_int(int(n_threads)),
             _p(E, ctypes.c_double))
    return int(att), E


def num_threads() -> int:
    return int(lib().sgo_num_threads())
