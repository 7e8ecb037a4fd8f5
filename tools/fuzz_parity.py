"""Randomised differential test on the GPU: oracle vs tensor-core vs sequential-FMA vs sparse
kernels on random integer models of random shapes (all must agree bit for bit)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
from oracle import oracle as orc



def run(seed0=0, cases=40, verbose=True):
  orc.build()
  eng = Engine(0)
  bad = 0
  for c in range(cases):
      rng = np.random.default_rng(seed0 * 1000 + c)
      n = int(rng.choice([16, 17, 31, 32, 33, 100, 127, 128, 129, 255, 300, 511, 513, 777, 1000, 1024, 1024, 1100, 2048]))
      R = int(rng.integers(1, 70))
      ns = int(rng.integers(1, 4))
      rule = str(rng.choice(["metropolis", "glauber", "heat_bath"]))
      planes = int(rng.integers(1, 4))
      dens = float(rng.choice([0.05, 0.3, 1.0]))
      a = rng.integers(-3, 4, size=(n, n)) * (rng.random((n, n)) < dens)
      J = np.triu(a, 1); J = (J + J.T).astype(np.float32)
      h = rng.integers(-2, 3, size=n).astype(np.float32)
      S0 = (rng.integers(0, 2, size=(R, n)) * 2 - 1).astype(np.int8)
      T = float(rng.choice([0.7, 1.5, 4.0]))
      # --- replay: oracle vs TC (injected uniforms, explicit sites with duplicates inside blocks)
      sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
      if rng.random() < 0.5:
          sites[:, : n // 2] = sites[:, n // 2: n // 2 + n // 2][:, ::-1] // 3   # many repeats
      uni = rng.random((R, ns, n), dtype=np.float32)
      temps = np.full(ns, T)
      eng.set_model(J, h); eng.alloc_replicas(R); eng.set_spins(S0); eng.init_fields()
      tr = eng.sweep(ns, temps, temps_sweep_stride=1, rule=rule, sites=sites, uniforms=uni, energy_trace=True,
                     kernel="tc", coupling_planes=planes).cpu().numpy()
      fin = eng.spins().cpu().numpy()
      ok = True
      for r in range(R):
          s = S0[r].astype(np.float32).copy()
          es, _ = orc.sweeps_scheduled(J, h, s, temps, rule, sites, uni[r])
          ok &= np.array_equal(fin[r], s.astype(np.int8)) and np.array_equal(tr[:, r].astype(np.float64), es)
      # --- Philox: TC == SIMT == CSR
      outs = []
      order = str(rng.choice(["random", "sequential"]))
      for mode in ("tc", "simt", "csr"):
          if mode == "csr":
              rows, cols = np.nonzero(J)
              rowptr = np.zeros(n + 1, np.int64); np.add.at(rowptr, rows + 1, 1)
              eng.set_model_csr(np.cumsum(rowptr), cols.astype(np.int32), J[rows, cols], h)
          else:
              eng.set_model(J, h)
          eng.alloc_replicas(R); eng.set_spins(S0); eng.init_fields()
          kw = {} if mode == "csr" else {"kernel": mode, "coupling_planes": planes}
          t2 = eng.sweep(ns, np.array([T]), rule=rule, seed=c, sweep_base=7, site_order=order, energy_trace=True, **kw).cpu().numpy()
          outs.append((eng.spins().cpu().numpy(), t2, eng.accepted().cpu().numpy(), eng.best()[0].cpu().numpy(), eng.best()[1].cpu().numpy()))
      for o in outs[1:]:
          for x, y in zip(outs[0], o):
              ok &= np.array_equal(x, y)
      if not ok:
          bad += 1
          print(f"MISMATCH case {c}: n={n} R={R} ns={ns} rule={rule} planes={planes} dens={dens} T={T} order={order}")
  if verbose:
    print(f"fuzz seed {seed0}: {cases - bad}/{cases} cases agree")
  return bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    sys.exit(1 if run(seed, ncases) else 0)
