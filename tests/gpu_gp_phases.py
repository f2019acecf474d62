"""Developer script (GPU box): per-phase cycle counts of the GP evaluation (build with -DGP_PHASE_TIMING,
run with MLRN_LIB pointing at that library)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mallorn_astrophysics_b200 import _native
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_table
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
table = make_table(n, config_index=3)
L = _native.lib()
buf = (ctypes.c_ulonglong * 12)()
wbuf = (ctypes.c_ulonglong * (16 * 8))()
has_w = hasattr(L, "mlrn_gp_warp_cycles")
with LightCurveStore.from_table(table) as s:
    st = np.zeros((n, 4), np.int32)
    s.features("gp2d", status=st)
    L.mlrn_gp_phase_cycles(buf, 1)
    if has_w: L.mlrn_gp_warp_cycles(wbuf, 1)
    s.features("gp2d", status=st)
    L.mlrn_gp_phase_cycles(buf, 0)
    if has_w: L.mlrn_gp_warp_cycles(wbuf, 0)
names = ["build", "pivot0", "phaseA(scale)+sync", "warp0: tile update + b", "phaseB tail wait (sync)", "gradient", "warp0: pivot", "optimiser step (thread 0) + barrier"]
tot = sum(buf[:8])
for k, nm in enumerate(names):
    if buf[k]: print(f"{nm:28s} {buf[k]:14d} cycles  {100*buf[k]/tot:5.1f}%")
print("evaluations per object (mean):", st[:, 1].mean())
if buf[10]: print(f"CTA 0: {buf[10]} evaluations, {buf[11]} sweep steps, {tot / buf[10]:.0f} cycles per evaluation; per step: "
                  f"scale {buf[2] / buf[11]:.0f}, update+b {buf[3] / buf[11]:.0f}, pivot {buf[6] / buf[11]:.0f}, tail {buf[4] / buf[11]:.0f}; "
                  f"build {buf[0] / buf[10]:.0f}, pivot0 {buf[1] / buf[10]:.0f}, gradient {buf[5] / buf[10]:.0f} per evaluation")
if buf[9]: print(f"tile_update on warp 1: {buf[8] / buf[9]:.0f} cycles per call over {buf[9]} calls")
if has_w and buf[11]:
    w = np.array(list(wbuf), dtype=np.float64).reshape(16, 8) / buf[11]
    print("per sweep step, cycles per warp: acquire  update  finalise  small  end-wait  (#updates)  pivot-work  scale+barrier")
    for k in range(16):
        print(f"  warp {k:2d}: {w[k,0]:8.0f} {w[k,1]:8.0f} {w[k,2]:8.0f} {w[k,3]:8.0f} {w[k,4]:8.0f}   ({w[k,5]:.2f})  {w[k,6]:8.0f} {w[k,7]:8.0f}")
