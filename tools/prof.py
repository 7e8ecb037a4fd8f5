"""Small profiling drivers (one script instead of a file each): python tools/prof.py <what> [args]

    tc         one short launch of the tensor-core sweep at the headline shape (for ncu --set full): [n] [replicas] [sweeps] [planes]
    simt       one short launch of the sequential-FMA sweep kernel: [n] [replicas] [sweeps]
    tma        the L2 -> shared-memory TMA stream probe alone (the roofline denominator; for its ncu capture)
    anneal     cProfile of 50 short GPUAnnealer.anneal() calls on N=100 (host overhead)
    batch      cProfile of BatchProcessor.process_models_batch on 64 models of N=100
    setmodels  host timings of the stacked-model cycle (set_models / alloc / sweep / best)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def tc(argv):
    """one short launch of the tensor-core sweep at the headline shape (for ncu --set full): [n] [replicas] [sweeps] [planes]"""
    import numpy as np, torch
    from spin_glass_anneal_rl_b200.engine import Engine
    n = int(argv[0]) if len(argv) > 0 else 4096
    R = int(argv[1]) if len(argv) > 1 else 8192
    sw = int(argv[2]) if len(argv) > 2 else 10
    P = int(argv[3]) if len(argv) > 3 else 3
    rs = np.random.RandomState(3003)
    Gm = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((Gm + Gm.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
    eng = Engine(0)
    eng.set_model(torch.from_numpy(J).cuda(), torch.zeros(n, device="cuda"))
    eng.alloc_replicas(R)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8))
    eng.init_fields()
    eng.sweep(sw, np.array([1.0]), seed=1, site_order="random", kernel="tc", coupling_planes=P)
    eng.sweep(sw, np.array([1.0]), seed=1, sweep_base=sw, site_order="random", kernel="tc", coupling_planes=P)
    torch.cuda.synchronize()
    print("ok", eng.energies().mean().item() / n)


def simt(argv):
    """one short launch of the sequential-FMA sweep kernel: [n] [replicas] [sweeps]"""
    import numpy as np, torch
    from spin_glass_anneal_rl_b200.engine import Engine
    n = int(argv[0]) if len(argv) > 0 else 4096
    R = int(argv[1]) if len(argv) > 1 else 148 * 10
    sw = int(argv[2]) if len(argv) > 2 else 1
    rs = np.random.RandomState(3003)
    Gm = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((Gm + Gm.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
    eng = Engine(0)
    eng.set_model(torch.from_numpy(J).cuda(), torch.zeros(n, device="cuda"))
    eng.alloc_replicas(R)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8))
    eng.init_fields()
    eng.sweep(sw, np.array([1.0]), seed=1, site_order="random")
    eng.sweep(sw, np.array([1.0]), seed=1, sweep_base=sw, site_order="random")
    torch.cuda.synchronize()
    print("ok", eng.energies().mean().item() / n)


def tma(argv):
    """the L2 -> shared-memory TMA stream probe alone (the roofline denominator; for its ncu capture)"""
    import os, sys
    from spin_glass_anneal_rl_b200.engine import Engine
    eng = Engine(0)
    nbytes = 4096 * 4096 * 4 + (1 << 20)
    a = eng.measure_tma_stream(nbytes, 17920, 8, 4096, False)
    b = eng.measure_tma_stream(nbytes, 49152, 4, 2048, False)
    print(f"tma stream probe: {a:.0f} GB/s (17.9 KB copies, 8 stages), {b:.0f} GB/s (48 KB copies, 4 stages)")


def anneal(argv):
    """cProfile of 50 short GPUAnnealer.anneal() calls on N=100 (host overhead)"""
    import time, cProfile, pstats
    import numpy as np, torch
    import spin_glass_anneal_rl_b200 as sg
    from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType
    n = 100
    rng = np.random.default_rng(n)
    a = rng.normal(size=(n, n)).astype(np.float32); J = np.triu(a, 1); J = J + J.T
    m = sg.IsingModel(sg.IsingModelConfig(n_spins=n, use_sparse=False)); m.set_couplings_from_matrix(torch.from_numpy(J))
    cfg = sg.GPUAnnealerConfig(n_sweeps=10, initial_temp=1.0, final_temp=1.0, schedule_type=ScheduleType.GEOMETRIC,
                               schedule_params={"alpha": 1.0}, record_interval=10, n_replicas=1, random_seed=1)
    ann = sg.GPUAnnealer(cfg)
    for _ in range(5): ann.anneal(m)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): ann.anneal(m)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)


def batch(argv):
    """cProfile of BatchProcessor.process_models_batch on 64 models of N=100"""
    import time, cProfile, pstats
    import numpy as np, torch
    import spin_glass_anneal_rl_b200 as sg
    from spin_glass_anneal_rl_b200.annealing.temperature_scheduler import ScheduleType
    from spin_glass_anneal_rl_b200.annealing.batch_processor import BatchConfig, BatchProcessor
    rng = np.random.default_rng(0)
    models = []
    for _ in range(64):
        a = rng.normal(size=(100, 100)).astype(np.float32); J = np.triu(a, 1); J = J + J.T
        m = sg.IsingModel(sg.IsingModelConfig(n_spins=100, use_sparse=False)); m.set_couplings_from_matrix(torch.from_numpy(J)); models.append(m)
    cfg = sg.GPUAnnealerConfig(n_sweeps=10, initial_temp=1.0, final_temp=1.0, schedule_type=ScheduleType.GEOMETRIC,
                               schedule_params={"alpha": 1.0}, record_interval=10, n_replicas=32, random_seed=1)
    bp = BatchProcessor(BatchConfig(batch_size=64), cfg)
    bp.process_models_batch(models); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3): bp.process_models_batch(models)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


def setmodels(argv):
    """host timings of the stacked-model cycle (set_models / alloc / sweep / best)"""
    import time
    import numpy as np, torch
    from spin_glass_anneal_rl_b200.engine import Engine
    rng = np.random.default_rng(0)
    M, n, R = 64, 100, 32
    J = rng.normal(size=(M, n, n)).astype(np.float32); h = rng.normal(size=(M, n)).astype(np.float32)
    eng = Engine(0)
    def t(label, f, reps=5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): f()
        torch.cuda.synchronize(); print(f"{label}: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms")
    t("set_models (numpy)", lambda: eng.set_models(J, h))
    Jt, ht = torch.from_numpy(J), torch.from_numpy(h)
    t("set_models (cpu tensor)", lambda: eng.set_models(Jt, ht))
    Jd, hd = Jt.cuda(), ht.cuda()
    t("set_models (cuda tensor)", lambda: eng.set_models(Jd, hd))
    eng.alloc_replicas(M * R)
    S = (torch.randint(0, 2, (M * R, n), device="cuda") * 2 - 1).to(torch.int8)
    eng.set_spins(S); eng.init_fields()
    t("alloc_replicas", lambda: eng.alloc_replicas(M * R))
    eng.set_spins(S); eng.init_fields()
    t("batch_energies", lambda: eng.batch_energies(S))
    t("sweep 10", lambda: eng.sweep(10, np.array([1.0]), seed=1))
    t("refresh", lambda: eng.refresh_fields())
    t("best", lambda: eng.best())
    print("---- full cycle")
    for it in range(3):
        marks = []
        def mark(label):
            torch.cuda.synchronize(); marks.append((label, time.perf_counter()))
        mark("start")
        eng.set_models(Jt, ht); mark("set_models")
        eng.alloc_replicas(M * R); mark("alloc")
        eng.set_spins(S); eng.init_fields(); mark("spins+fields")
        tr = eng.sweep(10, np.array([1.0]), seed=1, energy_trace=True); eng.refresh_fields(); mark("sweep")
        _, bs = eng.best(); mark("best")
        be = eng.batch_energies(bs); mark("batch_energies")
        print(" ".join(f"{l}={1e3 * (t1 - t0):.2f}" for (l, t1), (_, t0) in zip(marks[1:], marks[:-1])))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else ""
    table = {"tc": tc, "simt": simt, "tma": tma, "anneal": anneal, "batch": batch, "setmodels": setmodels}
    if what not in table:
        sys.exit(__doc__)
    table[what](sys.argv[2:])
