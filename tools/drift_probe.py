"""Drift of the TMEM-resident fields / in-kernel energies of the TC sweep on float couplings
(SK N=4096): resident vs exactly refreshed values after k sweeps."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from spin_glass_anneal_rl_b200.engine import Engine

n, R = 4096, 64
rs = np.random.RandomState(3003)
G = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
J = ((G + G.T) / 2).astype(np.float32)
np.fill_diagonal(J, 0.0)
eng = Engine(0)
eng.set_model(J, np.zeros(n, np.float32))
eng.alloc_replicas(R)
g = torch.Generator(device="cuda").manual_seed(1)
S0 = (torch.randint(0, 2, (R, n), device="cuda", generator=g) * 2 - 1).to(torch.int8)
for T in (1.0, 0.3):
    for k in (1, 2, 5, 10):
        for planes in (3, 1):
            eng.set_spins(S0); eng.init_fields()
            eng.sweep(k, np.array([T]), seed=3, kernel="tc", coupling_planes=planes)
            f = eng.fields().double(); e = eng.energies().double()
            e2, f2 = eng.batch_energies(eng.spins(), want_fields=True)
            de = (e - e2.double()); df = (f - f2.double())
            s = eng.spins().double()
            print(f"T={T} k={k:2d} P={planes}: dE mean {de.mean().item():+.5f} std {de.std().item():.5f} "
                  f"rel max {(de.abs()/e2.double().abs()).max().item():.2e} | df mean {df.mean().item():+.2e} "
                  f"|df| max {df.abs().max().item():.2e}  mean(s*df) {(s*df).mean().item():+.2e}  "
                  f"mean(sign(f)*df) {(f2.double().sign()*df).mean().item():+.2e}")
