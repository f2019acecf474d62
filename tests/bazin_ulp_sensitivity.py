"""Developer script (CPU): how far does the REFERENCE's own Bazin fit move when its model is perturbed by one ulp?

The reference fits with scipy's TRF and a forward-difference Jacobian (h ~ 1.5e-8 |x|).  This script runs the
reference's call (oracle/bazin.py = bazin_fitting.py:63-179) twice on every band of the golden fixtures: once with
numpy's exp, once with an exp that differs from it by at most one unit in the last place (math.exp, i.e. glibc's
libm, applied element-wise -- numpy's SIMD exp and glibc disagree in the last bit on a few per cent of arguments).
Both runs are "the reference"; the acceptance rule of tests/fitparity.py is then applied between them.  The match
rate it prints is the ceiling any independent implementation (a GPU included) can reach against the reference
under the 1e-3 rule -- the other fits stop elsewhere along a flat valley.

    python tests/bazin_ulp_sensitivity.py            # writes profiles/r01_bazin_ulp_sensitivity.json
"""
import json, math, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import bazin as ob
from oracle.common import group_objects
from fitparity import classify_bazin

def _exp1(v):
    try:
        return math.exp(v)
    except OverflowError:
        return math.inf


_vexp = np.frompyfunc(_exp1, 1, 1)

def exp_libm(x):
    return _vexp(np.asarray(x, dtype=np.float64)).astype(np.float64)

def bazin_libm(t, A, t0, tau_rise, tau_fall, B):
    return A * exp_libm(-(t - t0) / tau_fall) / (1.0 + exp_libm(-(t - t0) / tau_rise)) + B

def rows(g, model):
    orig = ob.bazin
    ob.bazin = model
    try:
        out = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k, v in group_objects(g["obj"], g["band"], g["t"], g["flux"], g["err"], int(g["n_obj"])):
                out.append(ob.object_row(v))
        return np.array(out)
    finally:
        ob.bazin = orig

if __name__ == "__main__":
    x = np.random.default_rng(0).uniform(-30, 30, 200000)
    differ = float((np.exp(x) != exp_libm(x)).mean())
    total = {"fits": 0, "match": 0, "no_worse": 0, "worse": 0, "nan_diff": 0}
    per = {}
    for name in ("ref_sparse", "ref_dense", "ref_edge"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        a, b = rows(g, ob.bazin), rows(g, bazin_libm)
        c = classify_bazin(b, a)
        per[name] = c
        for k in total:
            total[k] += c[k]
        print(name, c)
    res = {"what": "reference Bazin fit with numpy exp vs the same fit with glibc exp (<= 1 ulp apart), rule of tests/fitparity.py",
           "exp_values_that_differ": differ, "total": total, "match_fraction": total["match"] / total["fits"], "per_fixture": per}
    print(json.dumps(res["total"]), "match fraction %.3f" % res["match_fraction"], "exp differs on %.2f%% of arguments" % (100 * differ))
    with open(os.path.join(ROOT, "profiles", "r01_bazin_ulp_sensitivity.json"), "w") as fh:
        json.dump(res, fh, indent=1)
