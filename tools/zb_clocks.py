"""Development aid: per-warp clock64 trace of zonal_bwd3 (build with -DGPFY_ZB_CLOCKS into tools/libgpfy_clk.so)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from gpfy_b200 import build
build.LIB = os.path.abspath("tools/libgpfy_clk.so")
from gpfy_b200 import _lib
_lib.LIB = build.LIB
from gpfy_b200.synthetic import headline_problem
prob = headline_problem(227328, device="cuda")
hp, args = prob["hot"], prob["args"]
for _ in range(2):
    hp.elbo_valgrad_packed(*args)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (12 * 32 * 8))()
lib = _lib.load()
lib.gpfy_debug_zb_clocks.argtypes = [ctypes.c_void_p]
print("rc", lib.gpfy_debug_zb_clocks(buf))
c = np.array(buf[:], dtype=np.int64).reshape(32, 12, 8)
t0 = c[2, :, 0].min()
for it in range(2, 12):
    tw = it % 12
    print(f"tile {it}: tail warp {tw}")
    for w in range(12):
        r = c[it, w] - t0
        line = f"  w{w:2d} wait_full {r[0]:7d}->{r[1]:7d}  p1 {r[2]-r[1]:6d}  p2 {r[3]-r[2]:6d}"
        if w == tw:
            line += f"   TAIL wait_done->{r[4]:7d} gather+refill {r[5]-r[4]:6d} math {r[6]-r[5]:6d}"
        print(line)
