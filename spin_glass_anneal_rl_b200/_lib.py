"""ctypes binding of libsg_b200.so -- the C ABI declared in include/sg_b200.h.

There is no CPU fallback: if the shared library is missing or a symbol is
absent, loading fails loudly.  ``build()`` compiles the library in-tree with
nvcc for sm_100a (it cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int8, c_int32,
                    c_int64, c_uint32, c_uint64, c_void_p)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SG_B200_LIB") or os.path.join(_PKG, "libsg_b200.so")   # (override: A/B builds)
CSRC = os.path.join(_PKG, "csrc")

SG_RULE = {"metropolis": 0, "glauber": 1, "heat_bath": 2, "wolff": 3}
SG_RNG_PHILOX, SG_RNG_INJECTED = 0, 1
SG_SITES = {"sequential": 0, "random": 1, "explicit": 2, "random_per_block": 3, "checkerboard": 4}
SG_KERNEL = {"auto": 0, "simt": 1, "tc": 2, "small": 3}
SG_EXCHANGE = {"nearest_neighbor": 0, "all_pairs": 1}
SG_ABI_VERSION = 5


class SweepParams(Structure):
    _fields_ = [
        ("struct_size", c_uint32), ("n_sweeps", c_int32), ("rule", c_int32), ("rng_mode", c_int32),
        ("site_mode", c_int32), ("replicas_per_block", c_int32),
        ("temps", c_void_p), ("temps_sweep_stride", c_int64), ("temps_replica_stride", c_int64),
        ("seed", c_uint64), ("sweep_base", c_uint64),
        ("sites", c_void_p), ("sites_block_stride", c_int64), ("sites_sweep_stride", c_int64),
        ("uniforms", c_void_p), ("energy_trace", c_void_p),
        ("track_best", c_int32), ("kernel", c_int32), ("coupling_planes", c_int32),
        ("replica_base", c_int32), ("site_energy_changes", c_void_p),
    ]


class WolffParams(Structure):
    _fields_ = [
        ("struct_size", c_uint32), ("n_sweeps", c_int32), ("rng_mode", c_int32), ("site_mode", c_int32),
        ("temps", c_void_p), ("temps_sweep_stride", c_int64), ("temps_replica_stride", c_int64),
        ("seed", c_uint64), ("sweep_base", c_uint64),
        ("sites", c_void_p), ("sites_replica_stride", c_int64), ("sites_sweep_stride", c_int64),
        ("uniforms", c_void_p), ("uniforms_replica_stride", c_int64), ("uniforms_per_replica", c_int64),
        ("cursor", c_void_p), ("energy_trace", c_void_p),
        ("track_best", c_int32), ("replica_base", c_int32),
    ]


class ExchangeParams(Structure):
    _fields_ = [
        ("struct_size", c_uint32), ("parity", c_int32), ("rng_mode", c_int32),
        ("method", c_int32), ("seed", c_uint64), ("round", c_uint64), ("uniforms", c_void_p),
        ("energies_all", c_void_p),
    ]


# every symbol include/sg_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "sg_abi_version": (c_int, []),
    "sg_last_error": (c_char_p, []),
    "sg_create": (c_int, [c_int, POINTER(c_void_p)]),
    "sg_destroy": (None, [c_void_p]),
    "sg_set_model_dense": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "sg_set_model_dense_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "sg_set_model_csr": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "sg_set_model_lattice2d": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "sg_lattice_sequence_index": (c_int, [c_int, c_int, c_int]),
    "sg_set_model_groups": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sg_adaptive_temperature": (c_int, [c_void_p, c_int, c_uint64, c_int, c_int, c_double, c_double, c_double,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "sg_exchange_chain": (c_int, [c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                  c_uint64, c_uint64, POINTER(c_int32), c_void_p]),
    "sg_alloc_replicas": (c_int, [c_void_p, c_int, c_void_p]),
    "sg_set_spins": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_get_spins": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_set_best": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sg_set_accepted": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_set_ladder_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sg_upload_spins_async": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_set_spins_staged": (c_int, [c_void_p, c_int, c_void_p]),
    "sg_get_best_config": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sg_init_fields": (c_int, [c_void_p, c_void_p]),
    "sg_refresh_fields": (c_int, [c_void_p, c_void_p]),
    "sg_get_energies": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_get_fields": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_get_accepted": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sg_reset_best": (c_int, [c_void_p, c_void_p]),
    "sg_get_best": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sg_sweep": (c_int, [c_void_p, POINTER(SweepParams), c_void_p]),
    "sg_sweep_wolff": (c_int, [c_void_p, POINTER(WolffParams), c_void_p]),
    "sg_set_ladder": (c_int, [c_void_p, c_int, POINTER(c_double), c_void_p]),
    "sg_set_ladder_sharded": (c_int, [c_void_p, c_int, POINTER(c_double), c_int, c_int, c_void_p]),
    "sg_exchange": (c_int, [c_void_p, POINTER(ExchangeParams), c_void_p]),
    "sg_check_target": (c_int, [c_void_p, c_int, c_float, c_int32, c_void_p, c_void_p]),
    "sg_get_ladder_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                    c_void_p]),
    "sg_batch_energies": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sg_measure_stream_bandwidth": (c_int, [c_void_p, c_int64, c_int, c_int, POINTER(c_double)]),
    "sg_measure_tma_stream": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                      POINTER(c_double)]),
    "sg_tc_selftest": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sg_set_profiling": (c_int, [c_void_p, c_int]),
    "sg_get_profile": (c_int, [c_void_p, POINTER(c_double), POINTER(c_uint64), POINTER(c_double),
                               POINTER(c_uint64)]),
    "sg_tc_cluster_size": (c_int, [c_void_p]),
    "sg_tc_side_replicas": (c_int, [c_void_p, c_int, c_int]),
    "sg_query": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32),
                         POINTER(c_int32), POINTER(c_int32)]),
    "sg_launch_count": (c_uint64, [c_void_p]),
}

_lib = None


class SGError(RuntimeError):
    """A C-ABI call returned a negative sg_status."""


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libsg_b200.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(_PKG, "..", "include", "sg_b200.h"))
    stale = (not os.path.exists(LIB_PATH) or
             any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s)))
    if force or stale:
        cmd = ["make", "-C", CSRC, "-j4"] + (["-B"] if force else [])
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout)
        if out.returncode != 0:
            raise RuntimeError("building libsg_b200.so failed")
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the CUDA library; raises if it is missing (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (nvcc, sm_100a).  spin_glass_anneal_rl_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.sg_abi_version() != SG_ABI_VERSION:
        raise RuntimeError("libsg_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sg_last_error()
        raise SGError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
