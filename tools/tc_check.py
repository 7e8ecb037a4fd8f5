"""GPU checks of the tensor-core sweep kernel: oracle replay (integer J, bit-exact),
TC vs SIMT in Philox mode (integer J, identical), float-J drift, and throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
from oracle import oracle as orc

what = sys.argv[1] if len(sys.argv) > 1 else "all"
ok = True


def int_instance(rng, n, amp=2):
    a = rng.integers(-amp, amp + 1, size=(n, n))
    J = np.triu(a, 1); J = (J + J.T).astype(np.float32)
    h = rng.integers(-amp, amp + 1, size=n).astype(np.float32)
    return J, h


def replay(n, R, ns, rule="metropolis", planes=1):
    global ok
    rng = np.random.default_rng(n + R)
    J, h = int_instance(rng, n)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
    uni = rng.random((R, ns, n), dtype=np.float32)
    temps = np.linspace(3.0, 0.5, ns)
    eng = Engine(0)
    eng.set_model(J, h); eng.alloc_replicas(R); eng.set_spins(S0); eng.init_fields()
    trace = eng.sweep(ns, temps, temps_sweep_stride=1, rule=rule, sites=sites, uniforms=uni,
                      energy_trace=True, kernel="tc", coupling_planes=planes).cpu().numpy()
    final = eng.spins().cpu().numpy()
    fields = eng.fields().cpu().numpy()
    be, bs = eng.best()
    bad = 0
    for r in range(R):
        s = S0[r].astype(np.float32).copy()
        es, _ = orc.sweeps_scheduled(J, h, s, temps, rule, sites, uni[r])
        if not np.array_equal(final[r], s.astype(np.int8)) or not np.array_equal(trace[:, r].astype(np.float64), es):
            bad += 1
    Fo, Eo = orc.batch_fields_energies(J, h, final.astype(np.float32))
    fe = np.array_equal(fields.astype(np.float64), Fo)
    print(f"replay n={n} R={R} ns={ns} rule={rule} P={planes}: mismatching replicas {bad}/{R}, fields exact {fe}")
    ok &= (bad == 0) and fe


def tc_vs_simt(n, R, ns, T=1.5, planes=1):
    global ok
    rng = np.random.default_rng(n * 3 + R)
    J, h = int_instance(rng, n)
    S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    out = []
    for kern in ("simt", "tc"):
        eng = Engine(0)
        eng.set_model(J, h); eng.alloc_replicas(R); eng.set_spins(S0); eng.init_fields()
        tr = eng.sweep(ns, np.array([T]), seed=77, sweep_base=5, site_order="random", energy_trace=True,
                       kernel=kern, coupling_planes=planes).cpu().numpy()
        out.append((eng.spins().cpu().numpy(), tr, eng.accepted().cpu().numpy(), eng.best()[0].cpu().numpy()))
    same = all(np.array_equal(a, b) for a, b in zip(out[0], out[1]))
    print(f"tc_vs_simt n={n} R={R} ns={ns}: identical={same}  acc={out[1][2].sum() / (R * n * ns):.3f}")
    ok &= same


def sk(n, seed=3003):
    rs = np.random.RandomState(seed)
    G = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((G + G.T) / 2).astype(np.float32); np.fill_diagonal(J, 0.0)
    return J, np.zeros(n, np.float32)


def drift(n=4096, R=32, ns=10, planes=3):
    global ok
    J, h = sk(n)
    eng = Engine(0)
    eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda()); eng.alloc_replicas(R)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8)); eng.init_fields()
    eng.sweep(ns, np.array([1.0]), seed=3, kernel="tc", coupling_planes=planes)
    f = eng.fields().double(); e = eng.energies().double()
    e2, f2 = eng.batch_energies(eng.spins(), want_fields=True)
    df = (f - f2.double()).abs().max().item(); de = ((e - e2.double()).abs() / e2.double().abs()).max().item()
    print(f"drift n={n} P={planes} after {ns} sweeps: max|f_tc - f_exact|={df:.3e}  max rel energy err={de:.3e}  E/N={e.mean().item() / n:.4f}")
    ok &= df < 1e-2


def perf(n=4096, R=148 * 16, sweeps=5, T=1.0, planes=3, reps=3):
    J, h = sk(n)
    eng = Engine(0)
    eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda()); eng.alloc_replicas(R)
    eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8)); eng.init_fields()
    temps = np.array([T])
    eng.sweep(1, temps, seed=1, kernel="tc", coupling_planes=planes); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    best = 1e9; a0 = eng.accepted().sum().item()
    for i in range(reps):
        t0.record(); eng.sweep(sweeps, temps, seed=1, sweep_base=1 + i * sweeps, kernel="tc", coupling_planes=planes); t1.record()
        torch.cuda.synchronize(); best = min(best, t0.elapsed_time(t1))
    a1 = eng.accepted().sum().item()
    att = R * n * sweeps
    blocks = (R + 15) // 16
    jbytes = blocks * sweeps * n * n * 2 * planes
    print(f"perf n={n} R={R} P={planes} sweeps={sweeps} T={T}: {best:.3f} ms, {att / best / 1e6:.2f} Gattempts/s, "
          f"acc={(a1 - a0) / (att * reps):.3f}, J-stream {jbytes / best / 1e6:.0f} GB/s, E/N={eng.energies().mean().item() / n:.4f}")


if what in ("all", "replay"):
    replay(96, 5, 4); replay(100, 33, 3, planes=3); replay(128, 16, 2, rule="glauber"); replay(500, 20, 2, rule="heat_bath", planes=2)
    replay(1024, 17, 2)
if what in ("all", "vs"):
    tc_vs_simt(256, 40, 6); tc_vs_simt(1000, 64, 3); tc_vs_simt(4096, 32, 2)
if what in ("all", "drift"):
    drift(planes=3); drift(planes=2); drift(planes=1)
if what in ("all", "perf"):
    for P in (3, 2, 1):
        perf(planes=P)
    perf(R=8192, planes=3, sweeps=10); perf(R=8192, planes=2, sweeps=10); perf(T=0.3, planes=3)
print("TC CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
