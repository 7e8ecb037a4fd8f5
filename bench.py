#!/usr/bin/env python
"""bench.py -- spin-flip attempts/s of the batched Monte Carlo sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: `--sweeps` Metropolis sweeps of every
replica of the SK N=4096 instance (cfg3 of BASELINE.json), 8192 replicas PER GPU (weak
scaling: replicas are independent, no data-path collective; the only collective is the final
argmin allgather of the best energies).  `value` times the sweep launches with the state
resident in HBM (site-table + operand-gather + sweep kernels of the tensor-core path, all
inside the timed region); `e2e` times the same step through the host-buffer C-ABI path (pinned
host spins -> device, local-field initialisation, sweeps, best energies -> host).  The roofline
object is about the dominant kernel alone (sg::sweep_tc_kernel), timed with CUDA events by the
library around each of its launches.

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


def cpu_reference(n_threads, budget_s, sweeps=1):
    """Reference algorithm (oracle port) on the host cores; returns (attempts/s, sample text)."""
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
    return att / dt, threads, f"{R} replicas x {sweeps} sweep(s) of SK N={N_SPINS} at T={TEMPERATURE} ({dt:.1f} s)"


def workload_name(replicas, sweeps):
    return (f"SK dense N={N_SPINS} Gaussian J (cfg3), Metropolis sweep, T={TEMPERATURE}, "
            f"{replicas} replicas/GPU x {sweeps} sweeps per step, shared random site order")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference(0, 2.0)
    sample = ""
    threads = 1
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, threads, sample = cpu_reference(0, args.ref_budget)
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
                           "of the same workload: replicas x 1 sweep per step, see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": "attempts/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from spin_glass_anneal_rl_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    J, h = sk_instance()
    R, n, sweeps = args.replicas, N_SPINS, args.sweeps
    eng = Engine(local)
    eng.set_model(torch.from_numpy(J).to(dev), torch.from_numpy(h).to(dev))
    eng.alloc_replicas(R)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    spins_dev = (torch.randint(0, 2, (R, n), device=dev, generator=g, dtype=torch.int8) * 2 - 1).to(torch.int8)
    spins_host = spins_dev.cpu().pin_memory()
    eng.set_spins(spins_dev)
    eng.init_fields()
    temps = torch.full((1,), TEMPERATURE, dtype=torch.float64, device=dev)
    q = eng.query()
    gmax = q["max_replicas_per_block"]
    blocks = (R + gmax - 1) // gmax
    launches0 = eng.launch_count()

    kernel = args.kernel
    planes = args.planes

    def step(i):
        eng.sweep(sweeps, temps, seed=99 + rank, sweep_base=i * sweeps, site_order="random",
                  track_best=True, kernel=kernel, coupling_planes=planes)

    use_tc = kernel in ("auto", "tc")
    cluster = 1
    if use_tc:
        gmax = 16
        blocks = (R + gmax - 1) // gmax
        cluster = max(1, eng.tc_cluster_size())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    acc0 = eng.accepted().sum().item()
    eng.set_profiling(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    l0 = eng.launch_count()
    ev[0].record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev[1].record()
    barrier()
    ms_total = ev[0].elapsed_time(ev[1])
    gpu_launches = eng.launch_count() - l0
    prof = eng.profile()
    eng.set_profiling(False)
    acc1 = eng.accepted().sum().item()
    clocks = sampler.summary()

    # ---- end to end through the host-buffer path: pinned spins -> device, field init, sweeps,
    # best energies -> host, every step
    best_host = torch.empty(R, dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    # the upload of step i+1 runs on a side stream while step i computes (double buffer); every
    # step's upload and read-back are inside the timed region
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    bufs = [spins_dev, torch.empty_like(spins_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        with torch.cuda.stream(side):
            if i >= 2:
                side.wait_event(consumed[i % 2])
            bufs[i % 2].copy_(spins_host, non_blocking=True)
            ready[i % 2].record(side)

    t0.record()
    side.wait_event(t0)
    upload(0)
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            upload(i + 1)
        main.wait_event(ready[i % 2])
        eng.set_spins(bufs[i % 2])
        consumed[i % 2].record(main)
        eng.init_fields()
        step(1000 + i)
        best_host.copy_(eng.best_energies(), non_blocking=True)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1) / e2e_steps

    # ---- time to target energy (BASELINE metric, second half): parallel tempering on the same
    # instance, 64 rungs per ladder, R/64 ladders per GPU, 10 sweeps between exchanges; stop when
    # the best energy over all replicas of all ranks reaches E_target (0.97 x the Parisi
    # ground-state energy for Var J = 1/(2N))
    ttt = None
    if args.ttt_budget > 0 and R % 64 == 0:
        e_target = 0.97 * (-0.7632 / np.sqrt(2.0)) * n
        ladder = np.geomspace(2.0, 0.1, 64)
        runs = []
        exact = None
        budget_left = args.ttt_budget
        for run in range(args.ttt_seeds):
            if run == 0:
                eng.set_spins(spins_dev)
            else:
                g.manual_seed(1234 + rank + 7919 * run)
                eng.set_spins((torch.randint(0, 2, (R, n), device=dev, generator=g, dtype=torch.int8) * 2 - 1)
                              .to(torch.int8))
            eng.init_fields()
            eng.set_ladder(ladder)
            barrier()
            w0 = time.perf_counter()
            rounds, reached = 0, False
            while True:
                for _ in range(4):
                    eng.sweep(10, None, seed=4242 + rank + 1000 * run, sweep_base=rounds * 10,
                              site_order="random", track_best=True, kernel=kernel, coupling_planes=planes)
                    eng.refresh_fields()
                    eng.exchange(rounds & 1, seed=77 + rank + 1000 * run, round=rounds)
                    rounds += 1
                # one reduction carries both the best energy and "some rank is out of budget", so
                # every rank takes the same branch
                over = 1.0 if time.perf_counter() - w0 >= budget_left else 0.0
                b = torch.stack([eng.best_energies().min().double(),
                                 torch.tensor(-over, dtype=torch.float64, device=dev)])
                if world > 1:
                    dist.all_reduce(b, op=dist.ReduceOp.MIN)
                if b[0].item() <= e_target:
                    reached = True
                    break
                if b[1].item() < 0.0:
                    break
            torch.cuda.synchronize()
            secs = time.perf_counter() - w0
            runs.append({"seconds": secs, "sweeps": rounds * 10, "reached": reached})
            if run == 0:
                be, bs = eng.best()
                exact = eng.batch_energies(bs[int(torch.argmin(be).item())].reshape(1, n)).item()
            spent = torch.tensor([secs], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(spent, op=dist.ReduceOp.MAX)
            budget_left -= spent.item()
            if not reached or budget_left <= 0.0:
                break
        ok = [r["seconds"] for r in runs if r["reached"]]
        ttt = {"e_target": e_target, "reached": all(r["reached"] for r in runs),
               "seconds": float(np.median(ok)) if ok else runs[0]["seconds"],
               "sweeps": int(np.median([r["sweeps"] for r in runs])),
               "seeds": len(runs), "seconds_per_seed": [round(r["seconds"], 4) for r in runs],
               "best_energy_exact_local": exact, "ladder": "64 rungs, T geometric 2.0 -> 0.1, "
               f"{R // 64} ladders/GPU, exchange every 10 sweeps; seconds = median over the seeds run"}

    # max over ranks, whole-job aggregate
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    best_local = eng.best_energies().min().reshape(1).double()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(best_local) for _ in range(world)]
        dist.all_gather(gathered, best_local)  # the only collective of the path: final argmin
        best_global = torch.cat(gathered).min().item()
    else:
        best_global = best_local.item()
    ms_total, ms_e2e = t[0].item(), t[1].item()
    attempts_per_step = float(R) * n * sweeps * world
    value = attempts_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = attempts_per_step / (ms_e2e * 1e-3)
    acc_rate = (acc1 - acc0) / (float(R) * n * sweeps * args.steps)

    if rank == 0:
        # roofline of the dominant kernel (the sweep): algorithmic J-stream bytes = every block
        # streams the 16 coupling rows of every attempt block once (P bf16 planes on the
        # tensor-core path, one padded fp32 row per attempt on the SIMT path); peak = bandwidth of
        # the same transport (TMA bulk copies of an L2-resident buffer into a shared-memory ring,
        # one block per SM) measured on this box
        if use_tc:
            n_tc = (n + 127) // 128 * 128
            # per group of 16 replicas and sweep: every one of the n rows (n_tc couplings, 2 bytes
            # per plane) enters an SM once; a cluster pair takes each row in once for 32 replicas
            # (half of it per SM), i.e. half as many bytes per replica
            bytes_per_sweep_block = float(n) * n_tc * 2 * planes / cluster
            kname = "sg::sweep_tc_kernel"
            stream_desc = f"{planes} bf16 planes of J in UMMA operand layout ({planes * 2 * n * n_tc / 1e6:.0f} MB per sweep)"
        else:
            bytes_per_sweep_block = float(n) * q["n_pad"] * 4
            kname = "sg::sweep_kernel"
            stream_desc = "fp32 rows of J (73 MB padded)"
        n_klaunch = max(1, int(prof["sweep_launches"]))
        bytes_per_launch = float(blocks) * sweeps * bytes_per_sweep_block * args.steps / n_klaunch
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
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_sweep_tc_traffic.json")))[
                "dram_bytes_per_launch"]
        except Exception:
            pass
        # B operand per MMA: 16 attempts x (16 x cluster) replicas x 2 bytes against a 4 KB A tile
        smem_factor = 2.0 + 0.125 * cluster
        roofline = {"bound": "hbm", "achieved": achieved, "peak": l2_peak, "unit": "GB/s",
                    "frac": achieved / l2_peak, "traffic": traffic,
                    "peak_source": "measured on this box (sg_measure_tma_stream): TMA bulk-copy stream of "
                                   "an L2-resident J-sized buffer, one block per SM; the J stream ("
                                   + stream_desc + ") is read by all SMs in the same order, so it is "
                                   "served from the 126 MB L2 and the HBM copy peak is not the bound",
                    "hbm_peak": hbm, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "frac_of_hbm_peak": achieved / hbm,
                    "kernel": kname, "bytes_per_launch": bytes_per_launch,
                    "ms_per_launch": ms_launch, "launches_timed": n_klaunch,
                    # what actually bounds the tensor-core kernel: every J byte is written to shared
                    # memory by TMA and read from it by tcgen05.mma, which also reads the 512-byte
                    # B operand (the block's decisions) once per 4 KB A tile: 2.125 x the J stream
                    # through a 128 B/clk/SM port
                    "smem_traffic_gbs": smem_factor * achieved if use_tc else None,
                    "smem_peak_gbs": q["sm_count"] * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                    "smem_frac": (smem_factor * achieved) / (q["sm_count"] * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9) if use_tc else None,
                    "gather_ms_per_launch": prof["gather_ms"] / max(1, int(prof["gather_launches"]))}
        cpu = None
        if world == 1 or True:
            v, threads, sample = cpu_reference(0, args.cpu_budget)
            cpu = {"value": v, "unit": "attempts/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "attempts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(R, sweeps),
                       "kernel": ("tensor-core (tcgen05, TMEM-resident fields"
                                  + (", cluster pairs: 32 replicas per 2 SMs, half of the field columns each)"
                                     if cluster == 2 else ")")) if use_tc else "simt",
                       "coupling_planes": planes if use_tc else None,
                       "replicas_per_gpu": R, "sweeps_per_step": sweeps,
                       "replicas_per_group": gmax * cluster, "ctas_per_group": cluster,
                       "replica_groups": (R + gmax * cluster - 1) // (gmax * cluster),
                       "schedule": ("persistent CTAs (cluster pairs) walk (sweep chunk, replica group) work "
                                    "items; a group's fields/spins move through HBM between its items")
                                   if use_tc and (R + gmax * cluster - 1) // (gmax * cluster) > q["sm_count"] // cluster
                                   else "one CTA (cluster pair) per replica group",
                       "n_pad": q["n_pad"], "acceptance_rate": acc_rate,
                       "l2_policy": f"inputs (J planes 100 MB + operand stream {sweeps * 100} MB per step + "
                                    "170 MB replica state) exceed the 126 MB L2; no flush between steps",
                       "best_energy": best_global},
            "e2e": {"value": e2e_value, "unit": "attempts/s", "h2d_bytes_per_step": int(R) * n * world,
                    "d2h_bytes_per_step": int(R) * 4 * world, "ms_per_step": ms_e2e},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu,
            "time_to_target": ttt, "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=REPLICAS_PER_GPU)
    ap.add_argument("--sweeps", type=int, default=10, help="sweeps per step (= the reference's default exchange_interval)")
    ap.add_argument("--ttt-seeds", type=int, default=8,
                    help="independent time-to-target runs (different initial spins and RNG streams)")
    ap.add_argument("--ttt-budget", type=float, default=20.0,
                    help="wall-clock budget (s) of the time-to-target run; 0 skips it")
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
