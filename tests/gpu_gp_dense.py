import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_table
for scale, n in ((1.0, 4000), (1.4, 4000), (2.0, 4000)):
    tb = make_table(n, seed=11, count_scale=scale)
    with LightCurveStore.from_table(tb) as s:
        out = torch.empty((n, 27), dtype=torch.float64, device="cuda"); st = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
        s.features("gp2d", out=out, status=st); torch.cuda.synchronize()
        t0 = time.perf_counter(); s.features("gp2d", out=out, status=st); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        N = st[:, 3].double(); ev = st[:, 1].double()
        flop = float(((ev + 1) * N ** 3).sum())
        print(f"count_scale {scale}: mean N {float(N.mean()):.0f} max {int(N.max())}, >192: {float((N > 192).double().mean()):.2f}; {n/dt:.0f} obj/s, {flop/dt/1e12:.2f} TFLOP/s")
