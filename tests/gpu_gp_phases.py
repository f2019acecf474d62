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
with LightCurveStore.from_table(table) as s:
    st = np.zeros((n, 4), np.int32)
    s.features("gp2d", status=st)
    L.mlrn_gp_phase_cycles(buf, 1)
    s.features("gp2d", status=st)
    L.mlrn_gp_phase_cycles(buf, 0)
names = ["build", "pivot0", "phaseA(scale)+sync", "warp0: tile update + b", "phaseB tail wait (sync)", "gradient", "warp0: pivot", "-"]
tot = sum(buf[:8])
for k, nm in enumerate(names):
    if buf[k]: print(f"{nm:28s} {buf[k]:14d} cycles  {100*buf[k]/tot:5.1f}%")
print("evaluations per object (mean):", st[:, 1].mean())
if buf[9]: print(f"tile_update on warp 1: {buf[8] / buf[9]:.0f} cycles per call over {buf[9]} calls")
