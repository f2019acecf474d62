"""Developer script (GPU): throughput of the experimental research pass (K8) alone.
    python tests/gpu_research_rate.py [objects] [repeats]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = make_columns_device(n, 4242, "cuda")
with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
    out = torch.empty((n, 40), dtype=torch.float64, device="cuda")
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.features("research", z=c["z"], out=out); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"run {r}: {n} objects {dt*1e3:.1f} ms = {n/dt:.0f} obj/s; rows with an MHPS value {int((~torch.isnan(out[:, 29])).sum())}, "
              f"with a luminosity {int((~torch.isnan(out[:, 36])).sum())}")
