"""Randomised differential test on the GPU for the round-1 additions: the small-model kernel
(alone and with stacked models) and the partitioned group kernel, against the oracle (replay) and
against the kernels they must be indistinguishable from (Philox mode).  Integer models, so every
comparison is bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine
from oracle import oracle as orc


def state(eng):
    return (eng.spins().cpu().numpy(), eng.accepted().cpu().numpy(), eng.best()[0].cpu().numpy(),
            eng.best()[1].cpu().numpy(), eng.energies().cpu().numpy())


def run(seed0=0, cases=30):
    orc.build()
    eng, solo = Engine(0), Engine(0)
    bad = 0
    for c in range(cases):
        rng = np.random.default_rng(seed0 * 1000 + c)
        ok = True
        # ---------------- stacked small models: replay vs oracle, Philox vs SIMT per model
        M = int(rng.integers(1, 6)); n = int(rng.choice([2, 3, 31, 32, 33, 64, 65, 100, 160, 223, 224]))
        r = int(rng.integers(1, 20)); ns = int(rng.integers(1, 4))
        rule = str(rng.choice(["metropolis", "glauber", "heat_bath"]))
        a = rng.integers(-3, 4, size=(M, n, n)); J = np.triu(a, 1); J = (J + J.transpose(0, 2, 1)).astype(np.float32)
        if rng.random() < 0.3:
            J[:, np.arange(n), np.arange(n)] = rng.integers(-2, 3, size=(M, n))
        h = rng.integers(-2, 3, size=(M, n)).astype(np.float32)
        S0 = (rng.integers(0, 2, size=(M * r, n)) * 2 - 1).astype(np.int8)
        sites = rng.integers(0, n, size=(ns, n)).astype(np.int32)
        uni = rng.random((M * r, ns, n), dtype=np.float32)
        temps = np.full(ns, float(rng.choice([0.7, 1.5, 4.0])))
        eng.set_models(J, h); eng.alloc_replicas(M * r); eng.set_spins(S0); eng.init_fields()
        tr = eng.sweep(ns, temps, temps_sweep_stride=1, rule=rule, sites=sites, uniforms=uni, energy_trace=True).cpu().numpy()
        fin = eng.spins().cpu().numpy()
        for k in range(M * r):
            s = S0[k].astype(np.float32).copy()
            es, _ = orc.sweeps_scheduled(J[k // r], h[k // r], s, temps, rule, sites, uni[k])
            ok &= np.array_equal(fin[k], s.astype(np.int8)) and np.array_equal(tr[:, k].astype(np.float64), es)
        m = int(rng.integers(0, M))
        eng.set_models(J, h); eng.alloc_replicas(M * r); eng.set_spins(S0); eng.init_fields()
        eng.sweep(ns, temps, temps_sweep_stride=1, rule=rule, seed=c, sweep_base=3, site_order="random")
        stacked = eng.spins().cpu().numpy()[m * r:(m + 1) * r]
        if n >= 2:
            pad = np.ones(((m + 1) * r, n), np.int8); pad[m * r:] = S0[m * r:(m + 1) * r]
            solo.set_model(J[m], h[m]); solo.alloc_replicas((m + 1) * r); solo.set_spins(pad); solo.init_fields()
            solo.sweep(ns, temps, temps_sweep_stride=1, rule=rule, seed=c, sweep_base=3, site_order="random", kernel="simt")
            ok &= np.array_equal(solo.spins().cpu().numpy()[m * r:], stacked)
        # ---------------- group models: partitioned == one warp per word == CSR
        G = int(rng.integers(1, 40)); sizes = rng.integers(1, 30, size=G)
        group_of = np.repeat(np.arange(G), sizes).astype(np.int32); rng.shuffle(group_of)
        ng = group_of.shape[0]
        if ng >= 2:
            coup = rng.integers(-3, 6, size=G).astype(np.float32)
            hg = rng.integers(-20, 21, size=ng).astype(np.float32)
            Rg = int(rng.integers(1, 80)); nsg = int(rng.integers(1, 4)); Tg = float(rng.choice([3.0, 15.0, 60.0]))
            Sg = (rng.integers(0, 2, size=(Rg, ng)) * 2 - 1).astype(np.int8)
            Jg = (coup[group_of][:, None] * (group_of[:, None] == group_of[None, :])).astype(np.float32)
            np.fill_diagonal(Jg, 0.0)
            outs = []
            for mode in ("0", "1", "csr"):
                if mode == "csr":
                    rows, cols = np.nonzero(Jg); rowptr = np.zeros(ng + 1, np.int64); np.add.at(rowptr, rows + 1, 1)
                    eng.set_model_csr(np.cumsum(rowptr), cols.astype(np.int32), Jg[rows, cols], hg)
                else:
                    os.environ["SG_GRP_PART"] = mode
                    eng.set_model_groups(group_of, coup, hg)
                eng.alloc_replicas(Rg); eng.set_spins(Sg); eng.init_fields()
                t2 = eng.sweep(nsg, np.array([Tg]), rule=rule, seed=c, sweep_base=11, site_order="random", energy_trace=True).cpu().numpy()
                outs.append(state(eng) + (t2,))
            os.environ.pop("SG_GRP_PART", None)
            for o in outs[1:]:
                for x, y in zip(outs[0], o):
                    ok &= np.array_equal(x, y)
        if not ok:
            bad += 1
            print(f"MISMATCH case {c}: M={M} n={n} r={r} ns={ns} rule={rule} G={G} ng={ng}")
    print(f"fuzz seed {seed0}: {cases - bad}/{cases} cases agree")
    return bad


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    sys.exit(1 if run(seed, ncases) else 0)
