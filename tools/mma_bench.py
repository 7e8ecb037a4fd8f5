import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spin_glass_anneal_rl_b200.engine import Engine
eng = Engine(0)
f = eng._lib.sg_debug_mma_bench
f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
names = ["MN nosw sbo128", "MN nosw sbo144", "K nosw", "K sw128", "MN sw128"]
for v in range(5):
    for nd in (16, 32, 64, 128):
        out = (ctypes.c_longlong * 2)()
        rc = f(eng._h, v, nd, 8, out)
        f(eng._h, v, nd, 8, out)
        print(f"{names[v]:16s} N={nd:3d}: rc={rc} issue {out[0] / 256:.1f} clk/MMA, complete {out[1] / 256:.1f} clk/MMA")
for ce, label in ((0, "32 per elect, no commit"), (13, "12 per elect, no commit"), (12, "12 per elect + commit"), (14, "12 x (elect, MMA) + commit"),
                  (7, "6 per elect, no commit"), (6, "6 per elect + commit")):
    out = (ctypes.c_longlong * 2)()
    f(eng._h, ce * 16, 16, 8, out)
    f(eng._h, ce * 16, 16, 8, out)
    print(f"MN nosw N=16, {label}: issue {out[0] / 256:.1f} clk/MMA, complete {out[1] / 256:.1f} clk/MMA")
for iters in (1, 2, 4, 8, 16):
    out = (ctypes.c_longlong * 2)()
    f(eng._h, 15 * 16, 16, iters, out)
    f(eng._h, 15 * 16, 16, iters, out)
    print(f"tcgen05.ld of an untouched column while {iters * 32} MMAs are issued: ld+wait {out[0]} clk, all MMAs done after {out[1]} clk")
