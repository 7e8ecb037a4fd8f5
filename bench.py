#!/usr/bin/env python
"""bench.py -- spin-flip attempts/s of the batched Monte Carlo sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one parallel-tempering step of cfg3 (BASELINE.json): `--sweeps` Metropolis sweeps of
every replica of the SK N=4096 instance at its ladder temperature (64-rung ladders), an exact
refresh of fields / energies, and one replica-exchange round.  8192 replicas PER GPU (weak
scaling, the default) or in total (--scaling strong).  On several GPUs the ladders are global:
every exchange round all-gathers the per-replica energies over NCCL (the path's one collective,
inside the timed region) and every rank applies the same decisions.  `value` times the steps with
the state resident in HBM (site-table + operand-gather + sweep + refresh + exchange kernels, all
inside the timed region); `e2e` times the same step through the C ABI's host-buffer path (pinned
host spins -> device, local-field initialisation, the step, best configuration -> pinned host).
The roofline object is about the dominant kernel alone (sg::sweep_tc_kernel), timed with CUDA
events by the library around each of its launches.

`--impl reference` times the reference's CPU algorithm (the oracle port in C, all host
threads, per-attempt dot products + per-sweep O(N^2) energy exactly like the reference's
Python loop) on a bounded sample of the same workload.  The Python reference itself cannot
travel to the GPU box; its measured speed in the build container is in BASELINE.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SPINS = 4096
REPLICAS_PER_GPU = 8192
TEMPERATURE = 1.0
METRIC = "spin_flip_attempts_per_s"


def sk_instance(n=N_SPINS, seed=3003):
    """G ~ N(0, 1/sqrt(n)), J = (G + G^T)/2, zero diagonal, h = 0."""
    rs = np.random.RandomState(seed)
    G = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((G + G.T) / 2).astype(np.float32)
    np.fill_diagonal(J, 0.0)
    return J, np.zeros(n, np.float32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def summary(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference(n_threads, budget_s, sweeps=1, quench_sweeps=0):
    """Reference algorithm (oracle port) on the host cores; returns (attempts/s, threads, sample
    text, best energy of an optional fixed-temperature quench within the same budget class)."""
    from oracle import oracle as orc
    orc.build()
    J, h = sk_instance()
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is for the GPU
    # ranks' host side, not for the CPU baseline)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    threads = n_threads or max(orc.num_threads(), avail)
    rng = np.random.default_rng(1)
    # calibrate: one replica-sweep per thread
    R = threads
    S = (rng.integers(0, 2, size=(R, N_SPINS)) * 2 - 1).astype(np.float32)
    t0 = time.perf_counter()
    att, _ = orc.baseline_run(J, h, S, 1, TEMPERATURE, seed=1, n_threads=threads)
    dt = time.perf_counter() - t0
    reps = max(1, int(budget_s / max(dt, 1e-3)))
    R = threads * reps
    S = (rng.integers(0, 2, size=(R, N_SPINS)) * 2 - 1).astype(np.float32)
    t0 = time.perf_counter()
    att, _ = orc.baseline_run(J, h, S, sweeps, TEMPERATURE, seed=2, n_threads=threads)
    dt = time.perf_counter() - t0
    best = None
    if quench_sweeps > 0:
        # what the CPU arm reaches in a comparable budget: one replica per core, fixed T = 0.3
        Sq = (rng.integers(0, 2, size=(threads, N_SPINS)) * 2 - 1).astype(np.float32)
        tq = time.perf_counter()
        _, eq = orc.baseline_run(J, h, Sq, quench_sweeps, 0.3, seed=3, n_threads=threads)
        best = {"best_energy": float(np.min(eq)), "sweeps": quench_sweeps, "replicas": threads,
                "temperature": 0.3, "seconds": time.perf_counter() - tq}
    return att / dt, threads, f"{R} replicas x {sweeps} sweep(s) of SK N={N_SPINS} at T={TEMPERATURE} ({dt:.1f} s)", best


LADDER_RUNGS = 64
LADDER_T = (2.0, 0.1)   # hottest, coldest (geometric)


def workload_name(replicas, sweeps):
    return (f"SK dense N={N_SPINS} Gaussian J (cfg3), parallel tempering step: {sweeps} Metropolis sweeps of "
            f"{replicas} replicas/GPU on {LADDER_RUNGS}-rung ladders (T geometric {LADDER_T[0]} -> {LADDER_T[1]}) "
            f"+ exact energy refresh + one replica-exchange round, shared random site order")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference(0, 2.0)
    sample = ""
    threads = 1
    best = None
    t_all = time.perf_counter()
    for i in range(args.steps):
        v, threads, sample, b = cpu_reference(0, args.ref_budget, quench_sweeps=200 if i == 0 else 0)
        best = b or best
        vals.append(v)
    ms = (time.perf_counter() - t_all) * 1e3 / max(1, args.steps)
    value = float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "attempts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.replicas, args.sweeps),
                   "replicas_per_gpu": args.replicas, "sweeps_per_step": args.sweeps,
                   "note": "the reference's algorithm (CPU port, all host threads) on a bounded sample "
                           "of the same workload: replicas x 1 sweep per step at T=1, see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": "attempts/s", "cores": threads, "kind": "port",
                         "sample": sample, "quench": best},
        "e2e": {"value": value, "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from spin_glass_anneal_rl_b200.annealing.multi_gpu import gather_energies
    from spin_glass_anneal_rl_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    J, h = sk_instance()
    n, sweeps, K = N_SPINS, args.sweeps, LADDER_RUNGS
    kernel, planes = args.kernel, args.planes
    use_tc = kernel in ("auto", "tc")
    ladder = np.geomspace(LADDER_T[0], LADDER_T[1], K)
    eng = Engine(local)
    eng.set_model(torch.from_numpy(J).to(dev), torch.from_numpy(h).to(dev))
    q = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fresh_spins(R, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        return (torch.randint(0, 2, (R, n), device=dev, generator=g, dtype=torch.int8) * 2 - 1).to(torch.int8)

    def setup(R):
        """R replicas on this GPU = the global replicas [rank R, rank R + R) of world R: ladders
        are global (64 consecutive replicas), so under strong scaling at 8 GPUs a rank holds 16."""
        if eng.n_replicas != R:
            eng.alloc_replicas(R)
        eng.set_spins(fresh_spins(R, 1234 + rank))
        eng.init_fields()
        eng.set_ladder(ladder, n_global=R * world, replica_offset=rank * R)

    def step(i, R, seed=99):
        """One parallel-tempering step (reference parallel_tempering.py:108-114): `sweeps` sweeps of
        every replica at its current temperature, exact energies, one exchange round.  With more
        than one GPU the exchange is decided from the all-gathered energy table (collective C1:
        4 B per replica over NCCL), identically on every rank."""
        eng.sweep(sweeps, None, seed=seed, sweep_base=i * sweeps, site_order="random", track_best=True,
                  kernel=kernel, coupling_planes=planes, replica_base=rank * R)
        eng.refresh_fields()
        e_all = gather_energies(eng.energies(), R * world) if world > 1 else None
        eng.exchange(i & 1, seed=seed + 1, round=i, energies_all=e_all)
        return e_all

    def measure(R, steps, warmup):
        setup(R)
        for i in range(warmup):
            step(i, R)
        barrier()
        acc0 = eng.accepted().sum().item()
        eng.set_profiling(True)
        eng.profile()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        l0 = eng.launch_count()
        ev[0].record()
        for i in range(steps):
            step(warmup + i, R)
        ev[1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[1])
        launches = eng.launch_count() - l0
        prof = eng.profile()
        eng.set_profiling(False)
        acc1 = eng.accepted().sum().item()
        _, _, att, acc = eng.ladder_state()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        return {"ms_total": ms, "value": float(R) * world * n * sweeps * steps / (ms * 1e-3),
                "launches": int(launches), "prof": prof,
                "acceptance_rate": (acc1 - acc0) / (float(R) * n * sweeps * steps),
                "exchange_rate": float(acc.sum().item()) / max(1.0, float(att.sum().item()))}

    R_main = args.replicas if args.scaling == "weak" else max(K, args.replicas // world // K * K)
    sampler = ClockSampler(local)
    sampler.start()
    main_res = measure(R_main, args.steps, args.warmup)
    q = eng.query()
    cluster = max(1, eng.tc_cluster_size()) if use_tc else 1
    R = R_main
    # replicas the launch hands to cluster pairs on the SMs the clusters of 4 leave idle
    r_pairs = eng.tc_side_replicas(sweeps, planes) if use_tc else 0

    # ---- end to end through the C ABI's host-buffer path, every step: pinned host spins ->
    # sg_upload_spins_async (side stream, double buffered) -> sg_set_spins_staged -> field init ->
    # the same step -> sg_get_best_config (argmin energy, replica, configuration) -> pinned host
    spins_host = fresh_spins(R, 4321 + rank).cpu().pin_memory()
    out_e = torch.empty(1, dtype=torch.float32).pin_memory()
    out_r = torch.empty(1, dtype=torch.int32).pin_memory()
    out_s = torch.empty(n, dtype=torch.int8).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    eng.upload_spins_async(spins_host, 0, side)   # allocates the staging buffers outside the timed region
    eng.upload_spins_async(spins_host, 1, side)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)

    def upload(i):
        if i >= 2:
            side.wait_event(consumed[i % 2])
        eng.upload_spins_async(spins_host, i % 2, side)
        ready[i % 2].record(side)

    t0.record()
    side.wait_event(t0)
    upload(0)
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            upload(i + 1)
        main.wait_event(ready[i % 2])
        eng.set_spins_staged(i % 2)
        consumed[i % 2].record(main)
        eng.init_fields()
        step(1000 + i, R)
        eng.best_config(out_e, out_r, out_s)
    t1.record()
    barrier()
    clocks = sampler.summary()   # sampled over the timed steps and the end-to-end steps
    ms_e2e = t0.elapsed_time(t1) / e2e_steps
    e2e_best = float(out_e.item())

    # ---- time to target energy (BASELINE metric, second half): the same parallel tempering until
    # some replica's EXACT current energy (after the refresh of a round) is <= E_target = 0.97 x the
    # Parisi ground-state energy for Var J = 1/(2N).  The test runs on the device after
    # every exchange round (sg_check_target into pinned host memory; the all-gathered table when
    # there are several GPUs); the host polls the flag without synchronising, so a run stops within
    # a round or two of the hit and reports the round of the hit itself.
    ttt = None
    if args.ttt_budget > 0:
        e_target = 0.97 * (-0.7632 / np.sqrt(2.0)) * n
        runs = []
        budget_left = args.ttt_budget
        max_rounds = 400 if world == 1 else 60
        for run in range(args.ttt_seeds):
            eng.set_spins(fresh_spins(R, 777 + rank + 7919 * run))
            eng.init_fields()
            eng.set_ladder(ladder, n_global=R * world, replica_offset=rank * R)
            hit = torch.full((2,), -1, dtype=torch.int32).pin_memory()
            hit_dev = torch.full((2,), -1, dtype=torch.int32, device=dev)
            evs = []
            barrier()
            w0 = time.perf_counter()
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            rounds = 0
            while rounds < max_rounds:
                if rounds >= 2:
                    # the host runs at most two rounds ahead of the device.  Every rank must take the
                    # same branch (the next round contains a collective): after this wait the flag is
                    # final for rounds <= rounds-2 on every rank, later hits are ignored until then
                    evs[rounds - 2].synchronize()
                    if 0 <= int(hit[0]) <= rounds - 2:
                        break
                eng.sweep(sweeps, None, seed=4242 + 1000 * run, sweep_base=rounds * sweeps, site_order="random",
                          track_best=False, kernel=kernel, coupling_planes=planes, replica_base=rank * R)
                eng.refresh_fields()
                if world > 1:
                    e_all = gather_energies(eng.energies(), R * world)
                    m, arg = torch.min(e_all, dim=0)
                    now = torch.stack([torch.full_like(arg, rounds), arg]).to(torch.int32)
                    hit_dev = torch.where((m <= e_target) & (hit_dev[0] < 0), now, hit_dev)
                    hit.copy_(hit_dev, non_blocking=True)
                else:
                    e_all = None
                    eng.check_target(e_target, rounds, hit)
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs.append(e)
                eng.exchange(rounds & 1, seed=77 + 1000 * run, round=rounds, energies_all=e_all)
                rounds += 1
                # (wall-clock budget: single GPU only -- on several GPUs it would let ranks leave the
                # loop at different rounds; there max_rounds bounds a run)
                if world == 1 and rounds % 8 == 0 and time.perf_counter() - w0 > budget_left:
                    break
            torch.cuda.synchronize()
            hr = int(hit[0])
            reached = hr >= 0
            secs = ev0.elapsed_time(evs[hr]) * 1e-3 if reached else time.perf_counter() - w0
            rec = {"seconds": secs, "sweeps": (hr + 1) * sweeps if reached else rounds * sweeps, "reached": reached}
            if reached and int(hit[1]) // R == rank:
                # exact energy of the configuration that hit the target (it has not been swept since
                # only when the loop stopped at once; report the target test's own exact value)
                rec["replica"] = int(hit[1])
            runs.append(rec)
            spent = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(spent, op=dist.ReduceOp.MAX)
            budget_left -= spent.item()
            if budget_left <= 0.0:
                break
        ok = sorted(r["seconds"] for r in runs if r["reached"])
        ttt = {"e_target": e_target, "reached": all(r["reached"] for r in runs), "seeds": len(runs),
               "seconds": float(np.median(ok)) if ok else None,
               "sweeps_median": int(np.median([r["sweeps"] for r in runs])),
               "sweeps_per_seed": [r["sweeps"] for r in runs],
               "seconds_per_seed": [round(r["seconds"], 4) for r in runs],
               "decision": "exact energies (K2 refresh) of the current configurations after every "
                           f"{sweeps}-sweep round, tested on the device; time = CUDA events up to the round of the hit",
               "ladder": f"{K} rungs, T geometric {LADDER_T[0]} -> {LADDER_T[1]}, {R * world // K} ladders over "
                         f"{world} GPU(s), exchange every {sweeps} sweeps"}

    # ---- the other scaling curve in the same run: strong scaling = args.replicas in total
    strong = None
    if world > 1 and args.scaling == "weak":
        Rs = max(K, args.replicas // world // K * K)
        r2 = measure(Rs, args.steps, args.warmup)
        strong = {"value": r2["value"], "unit": "attempts/s", "replicas_total": Rs * world,
                  "replicas_per_gpu": Rs, "ms_per_step": r2["ms_total"] / args.steps,
                  "replica_groups_per_gpu": (Rs + 16 * (max(1, eng.tc_cluster_size()) if use_tc else 1) - 1)
                  // (16 * (max(1, eng.tc_cluster_size()) if use_tc else 1)),
                  "note": "same step, the 8192 replicas of cfg3 sharded over the GPUs (ladders global, "
                          "exchange through the all-gathered energy table)"}

    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    best_local = eng.best_energies().min().reshape(1).double()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(best_local, op=dist.ReduceOp.MIN)
    ms_e2e = t.item()
    ms_total = main_res["ms_total"]
    prof = main_res["prof"]
    attempts_per_step = float(R) * n * sweeps * world
    value = main_res["value"]
    e2e_value = attempts_per_step / (ms_e2e * 1e-3)

    if rank == 0:
        # roofline of the dominant kernel (the sweep): algorithmic J-stream bytes (SURVEY 8d,
        # shared-order mode: d * b_J / G per attempt) = every replica group takes every one of the
        # n coupling rows (n_tc couplings, 2 bytes per bf16 plane) in once per sweep, spread over
        # the C SMs of its cluster; peak = bandwidth of the same transport (TMA bulk copies of an
        # L2-resident buffer into a shared-memory ring, one block per SM) measured in this run
        ng = 16 * cluster
        groups = (R - r_pairs + ng - 1) // ng + r_pairs // 32
        if use_tc:
            n_tc = (n + 127) // 128 * 128
            bytes_per_group_sweep = float(n) * n_tc * 2 * planes
            kname = "sg::sweep_tc_kernel"
            stream_desc = f"{planes} bf16 planes of J in UMMA operand layout ({planes * 2 * n * n_tc / 1e6:.0f} MB per sweep)"
        else:
            gmax = q["max_replicas_per_block"]
            groups = (R + gmax - 1) // gmax
            ng = gmax
            bytes_per_group_sweep = float(n) * q["n_pad"] * 4
            kname = "sg::sweep_kernel"
            stream_desc = "fp32 rows of J (73 MB padded)"
        n_klaunch = max(1, int(prof["sweep_launches"]))
        bytes_per_launch = float(groups) * sweeps * bytes_per_group_sweep * args.steps / n_klaunch
        ms_launch = prof["sweep_ms"] / n_klaunch
        achieved = bytes_per_launch / (ms_launch * 1e-3) / 1e9
        l2_peak = max(eng.measure_tma_stream(J.nbytes + (1 << 20), 17920, 8, 4096, False),
                      eng.measure_tma_stream(J.nbytes + (1 << 20), 49152, 4, 2048, False))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        traffic, traffic_src, lts = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_sweep_tc_traffic.json")))
            traffic, lts, traffic_src = tj["dram_bytes_per_launch"], tj.get("lts_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
        smem_factor = 2.0 + ng / 128.0   # TMA write + operand read + B operand (32 NG bytes per 4 KB A tile)
        sm_clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        smem_peak = q["sm_count"] * 128 * sm_clk / 1e9
        roofline = {"bound": "l2", "achieved": achieved, "peak": l2_peak, "unit": "GB/s",
                    "frac": achieved / l2_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "lts_bytes_per_launch": lts,
                    "bytes_per_attempt": bytes_per_group_sweep / (n * ng),
                    "peak_source": "measured in this run (sg_measure_tma_stream): TMA bulk-copy stream of an "
                                   "L2-resident J-sized buffer, one block per SM, no compute; the J stream ("
                                   + stream_desc + ") is read by all SMs in the same order, so it is served "
                                   "from the 126 MB L2 and the HBM copy peak is not the bound.  The kernel "
                                   "itself is bound by shared-memory bandwidth (smem_frac): growing the replica "
                                   "group (clusters) cuts the J bytes an attempt needs, which lowers `frac` "
                                   "while the throughput rises",
                    "hbm_peak": hbm, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "frac_of_hbm_peak": achieved / hbm,
                    "kernel": kname, "bytes_per_launch": bytes_per_launch,
                    "ms_per_launch": ms_launch, "launches_timed": n_klaunch,
                    "kernel_share_of_step": prof["sweep_ms"] / ms_total,
                    "smem_traffic_gbs": smem_factor * achieved if use_tc else None,
                    "smem_peak_gbs": smem_peak,
                    "smem_frac": (smem_factor * achieved) / smem_peak if use_tc else None,
                    "gather_ms_per_launch": prof["gather_ms"] / max(1, int(prof["gather_launches"]))}
        if use_tc:
            # what the kernel is actually bound by (profiles/r2_notes.md): the tensor core's rate per
            # INSTRUCTION.  A block of 16 attempts needs (n_tc / 128) tiles x planes MMAs per replica
            # group, spread over the group's `cluster` SMs; one M128 x N(16 cluster) x K16 MMA takes
            # `clk_per_mma` clocks in isolation (tools/mma_dep_bench.py: 41 / 41 / 49 / 65 at N = 16 /
            # 32 / 64 / 128), so the MMA time per SM is a floor for the launch
            clk_per_mma = {1: 41.0, 2: 41.0, 4: 49.0, 8: 65.0}.get(cluster, 49.0)
            nblk = (n + 15) // 16
            mma_per_launch = float(groups) * sweeps * nblk * (n_tc // 128) * planes * args.steps / n_klaunch
            slots = min(groups, max(1, (q["sm_count"] // cluster) * cluster // cluster)) * cluster
            if cluster == 4:
                slots = min(groups, 33) * 4   # cudaOccupancyMaxActiveClusters: 33 clusters of 4 (132 SMs)
                if r_pairs:
                    slots = q["sm_count"]     # ... and cluster pairs on the other 16
            roofline["tensor_issue"] = {
                "mma_per_launch": mma_per_launch, "clk_per_mma_isolated": clk_per_mma, "sms_used": slots,
                "frac": mma_per_launch * clk_per_mma / slots / (ms_launch * 1e-3 * sm_clk),
                "mma_tflops": float(R) * sweeps * nblk * (n_tc // 128) * planes * args.steps / n_klaunch
                              * 2.0 * 128 * 16 / (ms_launch * 1e-3) / 1e12,
                "note": "share of the kernel's time that the tensor pipe of a used SM needs for its MMA "
                        "instructions at their isolated rate; the remainder is the two barrier round trips "
                        "per block between MMA completion, the raw field reads and the next issue"}
        v, threads, sample, quench = cpu_reference(0, args.cpu_budget, quench_sweeps=200)
        cpu = {"value": v, "unit": "attempts/s", "cores": threads, "kind": "port", "sample": sample,
               "quench": quench}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "attempts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(R, sweeps),
                       "kernel": (f"tensor-core (tcgen05, TMEM-resident fields, {16 * cluster} replicas per "
                                  f"cluster of {cluster} SM(s), 1/{cluster} of the field columns each"
                                  + (f"; the last {r_pairs} replicas as cluster pairs of 32 on the 16 SMs the "
                                     "clusters of 4 leave idle, in a concurrent launch" if r_pairs else "") + ")")
                       if use_tc else "simt",
                       "replicas_on_cluster_pairs": r_pairs,
                       "coupling_planes": planes if use_tc else None,
                       "replicas_per_gpu": R, "replicas_total": R * world, "sweeps_per_step": sweeps,
                       "replicas_per_group": ng, "ctas_per_group": cluster, "replica_groups": groups,
                       "collective": ("all_gather_into_tensor of the per-replica energies (NCCL, "
                                      f"{4 * R * world} bytes) before every exchange round") if world > 1 else None,
                       "n_pad": q["n_pad"], "acceptance_rate": main_res["acceptance_rate"],
                       "exchange_acceptance": main_res["exchange_rate"],
                       "l2_policy": f"inputs (J planes 100 MB + operand stream {sweeps * 100} MB per step + "
                                    "170 MB replica state) exceed the 126 MB L2; no flush between steps",
                       "best_energy": best_local.item()},
            "e2e": {"value": e2e_value, "unit": "attempts/s", "h2d_bytes_per_step": int(R) * n * world,
                    "d2h_bytes_per_step": (8 + n) * world, "ms_per_step": ms_e2e,
                    "path": "sg_upload_spins_async (pinned host) -> sg_set_spins_staged -> sg_init_fields -> "
                            "sg_sweep + sg_refresh_fields + sg_exchange -> sg_get_best_config (pinned host)",
                    "best_energy_last_step": e2e_best},
            "gpu_launches": main_res["launches"], "roofline": roofline, "cpu_baseline": cpu,
            "time_to_target": ttt, "strong_scaling": strong, "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=REPLICAS_PER_GPU,
                    help="replicas per GPU (weak scaling) or in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --replicas per GPU; strong: --replicas over all GPUs (cfg3: 8192 over 8). "
                         "A weak run on several GPUs also reports the strong-scaling number of the same step")
    ap.add_argument("--sweeps", type=int, default=10, help="sweeps per step (= the reference's default exchange_interval)")
    ap.add_argument("--ttt-seeds", type=int, default=32,
                    help="independent time-to-target runs (different initial spins and RNG streams)")
    ap.add_argument("--ttt-budget", type=float, default=25.0,
                    help="wall-clock budget (s) of the time-to-target runs; 0 skips them")
    ap.add_argument("--kernel", default="auto", choices=["auto", "tc", "simt"])
    ap.add_argument("--planes", type=int, default=3,
                    help="bf16 planes per coupling on the tensor-core path (3 = exact fp32 couplings)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work")
    ap.add_argument("--ref-budget", type=float, default=10.0, help="seconds per reference step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
