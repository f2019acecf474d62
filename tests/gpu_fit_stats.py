"""Developer script (GPU): lane occupancy of the thread-per-fit trips (build with -DFIT_STATS, select with MLRN_LIB).
    MLRN_LIB=variants/lib_fitstats.so python tests/gpu_fit_stats.py bazin 1000000"""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200 import _native
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device
mod = sys.argv[1]; n = int(sys.argv[2])
c = make_columns_device(n, 4242, "cuda")
lib = ctypes.CDLL(os.environ["MLRN_LIB"])
buf = (ctypes.c_ulonglong * 8)()
with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
    out = torch.empty((n, _native.COLS[mod]), dtype=torch.float64, device="cuda")
    lib.mlrn_fit_stats(buf, 1)
    s.features(mod, out=out)
    lib.mlrn_fit_stats(buf, 1)
    trips, act, jac, sm, mx, smj, mxj, tj = [int(x) for x in buf]
    print(f"{mod} {n} objects: warp trips {trips}, active lanes/trip {act/trips:.2f}, lanes with a Jacobian pass/trip {jac/trips:.2f}, "
          f"rows: mean over active {sm/max(act,1):.1f} vs warp max {mx/trips:.1f}; Jacobian rows: mean {smj/max(jac,1):.1f} vs warp max {mxj/max(tj,1):.1f}; trips with a pass {tj/trips:.3f}")
