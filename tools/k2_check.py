"""K2-TC (int8 tensor-core field init) vs fp64 numpy + timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
ok = True
for n, R, integer in ((100, 33, False), (48, 5, True), (1500, 70, False), (4096, 300, True), (4096, 8192, False), (5000, 40, False)):
    rng = np.random.default_rng(n)
    if integer:
        a = rng.integers(-3, 4, size=(n, n)); J = np.triu(a, 1); J = (J + J.T).astype(np.float32)
        h = rng.integers(-2, 3, size=n).astype(np.float32)
    else:
        a = rng.standard_normal((n, n)) / np.sqrt(n); J = ((a + a.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
        h = (0.3 * rng.standard_normal(n)).astype(np.float32)
    S = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
    eng = Engine(0)
    eng.set_model(J, h); eng.alloc_replicas(R); eng.set_spins(S)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    eng.init_fields(); torch.cuda.synchronize()
    t0.record(); eng.init_fields(); t1.record(); torch.cuda.synchronize()
    F = eng.fields().cpu().numpy().astype(np.float64)
    E = eng.energies().cpu().numpy().astype(np.float64)
    Sd = S.astype(np.float64)
    Fo = Sd @ J.astype(np.float64).T + h.astype(np.float64)
    Eo = -0.5 * np.einsum("ri,ri->r", Sd, Fo + h.astype(np.float64))
    ef = np.abs(F - Fo).max(); ee = (np.abs(E - Eo) / np.maximum(1, np.abs(Eo))).max()
    exact = np.array_equal(F, Fo) if integer else None
    print(f"n={n} R={R} integer={integer}: init {t0.elapsed_time(t1):.3f} ms  max|F-F64|={ef:.3e} (ulp(1)=1.2e-7) exact={exact}  max rel E err={ee:.2e}")
    ok &= (exact if integer else ef < 5e-7)
print("K2 CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
