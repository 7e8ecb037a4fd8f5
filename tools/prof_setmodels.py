import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
