import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scoary_b200.engine import Engine
e = Engine(0)
out = (ctypes.c_double * 8)()
e._lib.sb_debug_pipe_rates.restype = ctypes.c_int
e._lib.sb_debug_pipe_rates.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
e._check(e._lib.sb_debug_pipe_rates(e._ctx, 4096, out))
names = ["viaddmax_s32", "viaddmax_s16x2", "vimax3_s16x2", "vadd2", "lop3+shf(x2 instr)", "vimax3_s32", "isetp+sel(+add)", "imad"]
for n, v in zip(names, out):
    print("%-22s %.3e warp-instr-groups/s  -> %.1f per clk per SM @1.965GHz" % (n, v, v / 148 / 1.965e9))
