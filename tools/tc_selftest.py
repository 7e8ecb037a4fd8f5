"""GPU check of the tensor-core rank-16 field update (sg_tc_selftest) against numpy."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
from spin_glass_anneal_rl_b200._lib import check


def bf16_round(x):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return t.to(torch.bfloat16).to(torch.float32).numpy()


def planes_of(J):
    hi = bf16_round(J); r1 = (J - hi).astype(np.float32)
    mid = bf16_round(r1); r2 = (r1 - mid).astype(np.float32)
    lo = bf16_round(r2)
    return hi, mid, lo


def run(n, integer, seed=0):
    rng = np.random.default_rng(seed)
    if integer:
        J = rng.integers(-3, 4, size=(n, n)).astype(np.float32)
    else:
        J = rng.normal(0, 1 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((J + J.T) / 2).astype(np.float32) if not integer else (np.triu(J, 1) + np.triu(J, 1).T).astype(np.float32)
    h = np.zeros(n, np.float32)
    eng = Engine(0)
    eng.set_model(J, h)
    sites = rng.integers(0, n, size=16).astype(np.int32)
    deltas = rng.choice([-2.0, 0.0, 2.0], size=(16, 16)).astype(np.float32)
    f_in = rng.normal(0, 1, size=(16, n)).astype(np.float32)
    if integer:
        f_in = np.round(f_in * 4).astype(np.float32)
    planes = planes_of(J)
    ok = True
    for P in (1, 2, 3):
        out = np.empty_like(f_in)
        check(eng._lib.sg_tc_selftest(eng._h, P, sites.ctypes.data_as(ctypes.c_void_p),
                                      deltas.ctypes.data_as(ctypes.c_void_p),
                                      f_in.ctypes.data_as(ctypes.c_void_p),
                                      out.ctypes.data_as(ctypes.c_void_p)), "sg_tc_selftest")
        Jq = sum(p.astype(np.float64) for p in planes[:P])
        # Jt[i][j] = J[j][i]; J symmetric here.  F[r][:] += sum_k deltas[k][r] * Jq[sites[k]][:]
        ref = f_in.astype(np.float64) + np.einsum("kr,kj->rj", deltas.astype(np.float64), Jq[sites])
        err = np.abs(out - ref).max()
        exact = np.array_equal(out.astype(np.float64), ref) if integer else None
        jerr = np.abs(Jq - J).max()
        print(f"n={n} integer={integer} P={P}: max|out-ref|={err:.3e} exact={exact} max|J'-J|={jerr:.3e}")
        ok &= (err < 1e-5) if not integer else bool(exact)
    return ok


if __name__ == "__main__":
    good = True
    for n, integer in ((4096, True), (4096, False), (1000, True), (384, False)):
        good &= run(n, integer)
    print("TC SELFTEST", "PASS" if good else "FAIL")
    sys.exit(0 if good else 1)
