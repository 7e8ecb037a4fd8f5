"""One short launch of the tensor-core sweep at the headline shape, for ncu --set full."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
sw = int(sys.argv[3]) if len(sys.argv) > 3 else 10
P = int(sys.argv[4]) if len(sys.argv) > 4 else 3
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
