"""Developer script (GPU): one pass on a small batch in which object 17 has NaN / +inf / -inf fluxes.
    python tests/gpu_nonfinite_probe.py <module> [nan|pinf|ninf|all]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_table
m = sys.argv[1]; kind = sys.argv[2] if len(sys.argv) > 2 else "all"
tb = make_table(64, seed=5)
bad = tb.flux.copy()
rows = np.flatnonzero(tb.obj == 17)
if kind in ("nan", "all"): bad[rows[3]] = np.nan
if kind in ("pinf", "all"): bad[rows[7]] = np.inf
if kind in ("ninf", "all"): bad[rows[11]] = -np.inf
with LightCurveStore(tb.obj, tb.band, tb.t, bad, tb.err, tb.n_obj) as s:
    out = s.features(m, z=tb.z if m in ("physics", "research") else None)
print(m, kind, "ok", np.isnan(out[17]).mean())
