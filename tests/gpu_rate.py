"""Developer script (GPU): throughput of single passes on the C4 generator, device-resident inputs.
    python tests/gpu_rate.py bazin,powerlaw,gp2d [objects] [repeats]      (MLRN_LIB selects a developer build)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200 import _native
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device
mods = sys.argv[1].split(",")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = make_columns_device(n, 4242, "cuda")
tag = os.path.basename(os.environ.get("MLRN_LIB", "default"))
with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
    for m in mods:
        out = torch.empty((n, _native.COLS[m]), dtype=torch.float64, device="cuda")
        best = 1e30
        for r in range(reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.features(m, z=c["z"] if m in ("physics", "research") else None, out=out); b.record()
            torch.cuda.synchronize()
            if r: best = min(best, a.elapsed_time(b))
        nanf = float(torch.isnan(out).double().mean())
        chk = float(torch.nan_to_num(out, nan=0.0, posinf=0.0, neginf=0.0).abs().clamp(max=1e6).sum())
        print(f"[{tag}] {m}: {n} objects best {best:.2f} ms -> {n/best*1e3:.0f} obj/s; NaN fraction {nanf:.4f}; abs-sum {chk:.6e}", flush=True)
