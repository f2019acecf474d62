"""Developer script (GPU box): fit-parity classification per fixture + throughput of the fit passes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from fitparity import classify_bazin, classify_r2
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_table
for name in (() if "nogolden" in sys.argv else ("ref_sparse", "ref_dense", "ref_edge")):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    with LightCurveStore(g["obj"], g["band"], g["t"], g["flux"], g["err"], int(g["n_obj"])) as s:
        st = np.zeros((s.n_obj, 6, 2), np.int32)
        gb = s.features("bazin", status=st); gp = s.features("powerlaw")
    keep = g["bazin_present"] != (15 if name == "ref_edge" else -1)
    print(name, "bazin", classify_bazin(gb[g["bazin_present"]][keep], g["bazin_rows"][keep]))
    print(name, "powerlaw", classify_r2(gp[g["powerlaw_present"]][keep], g["powerlaw_rows"][keep]))
    print("   nfev median/max", np.median(st[..., 1][st[..., 1] > 0]), st[..., 1].max())
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
table = make_table(n, config_index=3)
with LightCurveStore.from_table(table) as s:
    for m in (() if "nofits" in sys.argv else ("bazin", "powerlaw")) + (("gp2d",) if "gp2d" in sys.argv else ()):
        out = torch.empty((n, {"bazin": 52, "powerlaw": 27, "gp2d": 27}[m]), dtype=torch.float64, device="cuda")
        s.features(m, out=out); torch.cuda.synchronize()
        t0 = time.perf_counter(); s.features(m, out=out); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{m}: {n} objects in {dt*1e3:.1f} ms -> {n/dt:.0f} obj/s; NaN fraction {float(torch.isnan(out).double().mean()):.3f}")
