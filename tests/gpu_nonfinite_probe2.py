"""Developer script (GPU): every pass on a small batch in which object 17 has a non-finite TIME or ERROR.
    python tests/gpu_nonfinite_probe2.py <t_nan|t_inf|e_nan|e_inf|e_zero|e_neg>"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_table
kind = sys.argv[1]
tb = make_table(64, seed=5)
t, e = tb.t.copy(), tb.err.copy()
rows = np.flatnonzero(tb.obj == 17)
if kind == "t_nan": t[rows[3]] = np.nan
if kind == "t_inf": t[rows[5]] = np.inf
if kind == "e_nan": e[rows[3]] = np.nan
if kind == "e_inf": e[rows[3]] = np.inf
if kind == "e_zero": e[rows[:40]] = 0.0
if kind == "e_neg": e[rows[3]] = -1.0
try:
    with LightCurveStore(tb.obj, tb.band, t, tb.flux, e, tb.n_obj) as s:
        for m in ("statistical", "colors", "shape", "tde", "physics", "research", "bazin", "powerlaw", "gp2d", "gp1d"):
            out = s.features(m, z=tb.z if m in ("physics", "research") else None)
            print(kind, m, "ok", round(float(np.isnan(out[17]).mean()), 3), flush=True)
except Exception as ex:
    print(kind, "FAILED:", str(ex)[:300], flush=True)
