"""Clock-stamp timeline of block 0 of the tensor-core sweep kernel (development aid)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
P = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 4096; R = 148 * 16
rs = np.random.RandomState(3003)
Gm = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
J = ((Gm + Gm.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
eng = Engine(0)
eng.set_model(torch.from_numpy(J).cuda(), torch.zeros(n, device="cuda"))
eng.alloc_replicas(R)
eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8))
eng.init_fields()
eng.sweep(1, np.array([1.0]), seed=1, kernel="tc", coupling_planes=P)
buf = torch.zeros(512 * 16 + 2048 + 512, dtype=torch.int64, device="cuda")
eng._lib.sg_debug_set_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
eng._lib.sg_debug_set_timeline(eng._h, ctypes.c_void_p(buf.data_ptr()))
eng.sweep(2, np.array([1.0]), seed=1, sweep_base=1, kernel="tc", coupling_planes=P)
torch.cuda.synchronize()
full_buf = buf.cpu().numpy()
t = full_buf[:512 * 16].reshape(512, 16)
cta = full_buf[8192:8192 + 4 * 148].reshape(148, 4)
print('CTA start (us, rel) min/med/max:', (cta[:, 0] - cta[:, 0].min()).min() / 1e3, np.median(cta[:, 0] - cta[:, 0].min()) / 1e3, (cta[:, 0] - cta[:, 0].min()).max() / 1e3)
dur = (cta[:, 1] - cta[:, 0]) / 1e3
print('CTA duration us min/med/max:', dur.min(), np.median(dur), dur.max(), ' clocks/ns:', np.median((cta[:, 3] - cta[:, 2]) / np.maximum(1, cta[:, 1] - cta[:, 0])))
print('durations by CTA (first 16):', np.round(dur[:16], 1))
ist = full_buf[10240:10248]
if ist[7] > 0:
    print("item phases (clk): dep wait", ist[1] - ist[0], " bit planes", ist[2] - ist[1], " fields->TMEM", ist[3] - ist[2],
          " sweeps", ist[4] - ist[3], " state->HBM", ist[5] - ist[4], " fence", ist[6] - ist[5], " sync", ist[7] - ist[6])
t0 = t[0, 0]
names = ["q_start", "q_tabs", "q_h0ok", "q_h1ok", "d_start", "d_pre", "d_rawok", "d_dec", "d_done", "m_wait", "m_decok", "m_r0ok", "m_done", "m_h0end", "m_h1go", "q_done"]
print("blk " + " ".join(f"{x:>8s}" for x in names))
for k in list(range(0, 4)) + list(range(100, 112)) + list(range(254, 258)):
    print(f"{k:3d} " + " ".join(f"{int(t[k, i] - t0):8d}" for i in range(16)))
d = t[40:240].astype(np.int64)
def m(a, b): return float((d[:, a] - d[:, b]).mean())
print("period", float((d[1:, 12] - d[:-1, 12]).mean()))
print("quarter: tables+theta", m(1, 0), " wait h0(k-2)", m(2, 1), " read h0 + wait h1", m(3, 2), " read h1", m(15, 3))
print("decision: wait tab", "n/a", " tables+cross prefetch", m(5, 4), " wait raw", m(6, 5), " 16 attempts", m(7, 6), " epilogue", m(8, 7))
print("mma: wait dec", m(10, 9), " wait r0(k+1)", m(11, 10), " chunks", m(12, 11))
print("mma detail: issue h0", m(13, 11), " wait r1(k+1)", m(14, 13), " issue h1", m(12, 14))
print("lags: q_done(k)->d_rawok(k)", m(6, 15), " d_done(k)->m_decok(k)", m(10, 8))
