"""The L2 -> shared-memory TMA stream probe alone (the roofline denominator bench.py uses), for an
ncu capture of its kernel: every SM pulls the same J-sized, L2-resident buffer in the same order
through a shared-memory ring with TMA bulk copies, no compute."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spin_glass_anneal_rl_b200.engine import Engine
eng = Engine(0)
nbytes = 4096 * 4096 * 4 + (1 << 20)
a = eng.measure_tma_stream(nbytes, 17920, 8, 4096, False)
b = eng.measure_tma_stream(nbytes, 49152, 4, 2048, False)
print(f"tma stream probe: {a:.0f} GB/s (17.9 KB copies, 8 stages), {b:.0f} GB/s (48 KB copies, 4 stages)")
