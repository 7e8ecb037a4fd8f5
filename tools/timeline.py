import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
n = 4096; G = int(sys.argv[1]) if len(sys.argv) > 1 else 10
R = 148 * G
rs = np.random.RandomState(3003)
Gm = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
J = ((Gm + Gm.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
eng = Engine(0)
eng.set_model(torch.from_numpy(J).cuda(), torch.zeros(n, device="cuda"))
eng.alloc_replicas(R)
eng.set_spins((torch.randint(0, 2, (R, n), device="cuda") * 2 - 1).to(torch.int8))
eng.init_fields()
eng.sweep(1, np.array([1.0]), seed=1, replicas_per_block=G)
buf = torch.zeros(256 * 8 + 8, dtype=torch.int64, device="cuda")
eng._lib.sg_debug_set_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
eng._lib.sg_debug_set_timeline(eng._h, ctypes.c_void_p(buf.data_ptr()))
eng.sweep(1, np.array([1.0]), seed=1, sweep_base=1, replicas_per_block=G)
torch.cuda.synchronize()
full = buf.cpu().numpy().astype(np.int64); t = full[:2048].reshape(256, 8)
print('per attempt warp0 wait/update/release:', full[2048:2051] / 4096., ' warp3:', full[2052:2055] / 4096.)
t0 = t[0, 1]
print("blk | dec: wait_raw_from  raw_ok  done | bulk0: at_decbar  dec_ok  attempts_done  published   (clk, relative)")
for k in list(range(0, 12)) + list(range(100, 108)):
    r = t[k] - t0
    print(f"{k:3d} | {r[1]:8d} {r[0]:8d} {r[2]:8d} | {r[3]:8d} {r[4]:8d} {r[5]:8d} {r[6]:8d}   dec={r[2]-r[0]} bulk={r[5]-r[4]} pub={r[6]-r[5]} decwait={r[4]-r[3]} rawwait={r[0]-r[1]}")
d = t[20:200]
print("mean per block: decide", (d[:,2]-d[:,0]).mean(), "rawwait", (d[:,0]-d[:,1]).mean(), "bulk", (d[:,5]-d[:,4]).mean(), "publish", (d[:,6]-d[:,5]).mean(), "decwait", (d[:,4]-d[:,3]).mean(), "period", (d[1:,5]-d[:-1,5]).mean())
