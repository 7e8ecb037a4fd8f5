"""Quick device-timed probe of the sweep kernel (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spin_glass_anneal_rl_b200.engine import Engine

def run(n, R, sweeps, T, G=0, order="random", reps=3):
    rs = np.random.RandomState(3003)
    Gm = rs.normal(0.0, 1.0 / np.sqrt(n), size=(n, n)).astype(np.float32)
    J = ((Gm + Gm.T) / 2).astype(np.float32); np.fill_diagonal(J, 0)
    h = np.zeros(n, np.float32)
    eng = Engine(0)
    eng.set_model(torch.from_numpy(J).cuda(), torch.from_numpy(h).cuda())
    eng.alloc_replicas(R)
    S = (torch.randint(0, 2, (R, n), device="cuda", dtype=torch.int8) * 2 - 1).to(torch.int8)
    eng.set_spins(S)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); eng.init_fields(); t1.record(); torch.cuda.synchronize()
    init_ms = t0.elapsed_time(t1)
    temps = np.array([T])
    eng.sweep(2, temps, seed=1, site_order=order, replicas_per_block=G)  # warm
    torch.cuda.synchronize()
    best = 1e9
    a0 = eng.accepted().sum().item()
    for i in range(reps):
        t0.record(); eng.sweep(sweeps, temps, seed=1, sweep_base=2 + i * sweeps, site_order=order, replicas_per_block=G); t1.record()
        torch.cuda.synchronize(); best = min(best, t0.elapsed_time(t1))
    a1 = eng.accepted().sum().item()
    att = R * n * sweeps
    acc = (a1 - a0) / (att * reps)
    q = eng.query(); Gu = G or q["max_replicas_per_block"]
    blocks = (R + Gu - 1) // Gu
    jbytes = blocks * sweeps * n * q["n_pad"] * 4
    print(f"n={n} R={R} G={G} sweeps={sweeps} T={T} order={order}: init {init_ms:.2f} ms, sweep {best:.3f} ms, "
          f"{att / best / 1e6:.2f} Gattempts/s, acc={acc:.3f}, J-stream {jbytes / best / 1e6:.1f} GB/s, E/N={eng.energies().mean().item() / n:.4f}")

if __name__ == "__main__":
    eng = Engine(0)
    for rb in (3584, 7168, 17920):
        for depth in (2, 4, 8, 12):
            if depth * rb > 200 * 1024: continue
            a_ = eng.measure_tma_stream(73 << 20, rb, depth, 4096, False)
            b_ = eng.measure_tma_stream(73 << 20, rb, depth, 4096, True)
            clk = rb * 148 / (a_ * 1e9) * 1.96e9
            print(f"tma probe row={rb}B depth={depth}: same-order {a_:.0f} GB/s ({clk:.0f} clk/row/SM), staggered {b_:.0f} GB/s")
    run(4096, 148 * 10, 5, 1.0)
    run(4096, 148 * 10, 5, 0.3)
    run(4096, 148 * 1, 5, 1.0, G=1)
    run(4096, 8192, 5, 1.0)
