"""Diagnostics: look-back depth and per-phase cycles of the forward scan kernel (SSDE_STATS build).
    SSDE_LIB_SUFFIX=_stats python scripts/stats_run.py [tracks] [steps]"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smoothsde_b200 import devgen

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
eng, par, info = devgen.make_ctcrw_device(T, m, device=0)
for _ in range(3):
    eng.eval(par, 1)
out = (C.c_uint64 * 32)()
eng._lib.ssde_debug_stats(eng._h, out, 1)
reps = 5
for _ in range(reps):
    eng.eval(par, 1)
print("ms", eng.last_eval_ms)
rc = eng._lib.ssde_debug_stats(eng._h, out, 0)
s = np.array(list(out), dtype=float)
li = eng.launch_info()
ntile = li["tiles_fwd"] * reps
for nm, o in (("fwd", 0), ("bwd", 16)):
    lb, win, spin, cyc = s[o:o + 4]
    if lb == 0:
        continue
    print(f"{nm}: look-backs {lb:.0f} windows/look-back {win / lb:.2f} spins/look-back {spin / lb:.2f} cycles/look-back {cyc / lb:.0f}")
    w0 = s[o + 4:o + 8] / ntile
    wo = s[o + 8:o + 12] / (3 * ntile)
    print(f"  warp 0 : phase1-2 {w0[0]:.0f}  barrier1 {w0[1]:.0f}  agg/look-back/barrier2 {w0[2]:.0f}  phase4 {w0[3]:.0f}  cycles per tile")
    print(f"  warps1-3: phase1-2 {wo[0]:.0f}  barrier1 {wo[1]:.0f}  agg/look-back/barrier2 {wo[2]:.0f}  phase4 {wo[3]:.0f}")
    if nm == "bwd":
        print(f"  X'eta_bar scatter: warp 0 {s[o + 12] / ntile:.0f}  warps1-3 {s[o + 13] / (3 * ntile):.0f}")
