#!/usr/bin/env python
"""Throughput of the Wolff cluster kernel (csrc/sg_wolff.cu) with the CPU arm beside it.

A dense-stored L x L periodic lattice with bonds of the sign the reference's cluster growth follows
(J = -1, a tenth of them +1), R replicas, Philox mode, fixed temperature.  Every cluster site costs
one walk over its coupling row (n columns, read from L2), so the unit is cluster sites (= flips =
row walks) per second; bytes = walks x 4 n.  CPU arm: the oracle's C restatement of the reference's
_wolff_cluster_dense on one core (the reference itself spends three .item() calls per column).

    python tools/wolff_bench.py [L] [R] [T] > profiles/r2_wolff.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spin_glass_anneal_rl_b200.engine import Engine  # noqa: E402


def lattice(L, seed=1, p_pos=0.1):
    rs = np.random.RandomState(seed)
    n = L * L
    J = np.zeros((n, n), np.float32)
    for x in range(L):
        for y in range(L):
            i = x * L + y
            for j in (((x + 1) % L) * L + y, x * L + (y + 1) % L):
                v = 1.0 if rs.rand() < p_pos else -1.0
                J[i, j] = J[j, i] = v
    return J


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
    T = float(sys.argv[3]) if len(sys.argv) > 3 else 3.5
    n = L * L
    J = lattice(L)
    h = np.zeros(n, np.float32)
    rs = np.random.RandomState(2)
    S0 = (rs.randint(0, 2, (R, n)) * 2 - 1).astype(np.int8)
    eng = Engine(0)
    eng.set_model(J, h)
    eng.alloc_replicas(R)
    eng.set_spins(S0)
    eng.init_fields()
    temps = np.array([T])
    eng.sweep_wolff(1, temps, seed=3, sweep_base=0)           # warm-up (builds the row-major copy)
    eng.synchronize()
    a0 = eng.accepted().sum().item()
    ns = 2
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.sweep_wolff(ns, temps, seed=3, sweep_base=1)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    walks = eng.accepted().sum().item() - a0
    out = {"config": f"Wolff, {L}x{L} periodic lattice stored dense (n={n}), J=-1 (10% +1), {R} replicas, T={T}",
           "n_spins": n, "replicas": R, "sweeps": ns, "ms": ms,
           "cluster_updates_per_s": R * ns * n / (ms * 1e-3),
           "walks_per_s": walks / (ms * 1e-3), "mean_cluster_size": walks / (R * ns * n),
           "row_bytes_per_s": walks * 4.0 * n / (ms * 1e-3),
           "note": "time includes the exact field/energy refresh after every sweep"}
    # CPU arm: same model, same temperature, a few replicas on one core
    from oracle import oracle as orc
    reps = 4
    dt = 0.0
    cw = 0
    for r in range(reps):
        s = S0[r].astype(np.float32).copy()
        st = orc.RawStream(orc.mt_raw_stream(50 + r, 30_000_000))
        orc.wolff_sweeps(J, h, s, [T], st)          # warm-up sweep, like the GPU arm
        t0 = time.perf_counter()
        _, flips, _ = orc.wolff_sweeps(J, h, s, [T] * ns, st)
        dt += time.perf_counter() - t0
        cw += int(flips.sum())
    out["cpu_baseline"] = {"walks_per_s": cw / dt if dt > 0 else None, "cores": 1,
                           "kind": "port", "mean_cluster_size": cw / (reps * ns * n),
                           "sample": f"{reps} replicas x {ns} sweeps ({dt:.1f} s)"}
    if out["cpu_baseline"]["walks_per_s"]:
        out["gpu_over_cpu"] = out["walks_per_s"] / out["cpu_baseline"]["walks_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
