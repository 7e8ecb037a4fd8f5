"""Do back-to-back tcgen05.mma on ONE accumulator tile run slower than on alternating tiles, and what
does a tcgen05.commit cost?  (round-2 question: the sweep kernel's MMA warp needs ~100 clk per MMA.)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spin_glass_anneal_rl_b200.engine import Engine
eng = Engine(0)
f = eng._lib.sg_debug_mma_bench
f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
labels = {8: "commits only (per commit)", 9: "stage pattern plane-major (2 tiles alternate) + commit",
          10: "all MMAs into one accumulator tile + commit per 6", 11: "stage pattern tile-major (3 planes back to back) + commit"}
for nd in (32, 64, 128):
    for ce in (9, 11, 10, 8):
        out = (ctypes.c_longlong * 2)()
        f(eng._h, ce * 16, nd, 8, out)
        f(eng._h, ce * 16, nd, 8, out)
        print(f"N={nd:3d} {labels[ce]:58s}: issue {out[0] / 256:.1f} clk, complete {out[1] / 256:.1f} clk per MMA", flush=True)
