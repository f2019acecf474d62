"""Developer script (GPU): throughput of the per-band GP pass alone, with the work it did.
    python tests/gpu_gp1d_rate.py [objects] [repeats]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = make_columns_device(n, 4242, "cuda")
with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
    out = torch.empty((n, 21), dtype=torch.float64, device="cuda"); st = torch.zeros((n, 12), dtype=torch.int32, device="cuda")
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.features("gp1d", out=out, status=st); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ev = st[:, 0::3].double().sum(); it = st[:, 1::3].double().sum(); fitted = (st[:, 2::3] >= 0).double().sum()
        print(f"run {r}: {n} objects {dt*1e3:.1f} ms = {n/dt:.0f} obj/s; bands fitted {int(fitted)}, evaluations {int(ev)} "
              f"({float(ev/fitted):.1f}/band), iterations {int(it)}; {float(ev)/dt/1e6:.2f} M evaluations/s; "
              f"NaN rows {int(torch.isnan(out[:, 18]).sum())}")
