"""Developer script (GPU): do the straggler tails of the fit passes hide under each other when Bazin and the decline models
run on two streams?    python tests/gpu_overlap2.py [objects]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200 import _native
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
c = make_columns_device(n, 4242, "cuda")
with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
    outs = {m: torch.empty((n, _native.COLS[m]), dtype=torch.float64, device="cuda") for m in ("bazin", "powerlaw")}
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    def run(overlap):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if overlap:
            sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
            s.features("bazin", out=outs["bazin"], stream=sa.cuda_stream)
            s.features("powerlaw", out=outs["powerlaw"], stream=sb.cuda_stream)
            torch.cuda.current_stream().wait_stream(sa); torch.cuda.current_stream().wait_stream(sb)
        else:
            for m in ("bazin", "powerlaw"):
                s.features(m, out=outs[m])
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    for mode in (False, True, False, True):
        print(f"{n} objects, bazin + powerlaw {'on two streams' if mode else 'one after the other'}: {run(mode):.1f} ms", flush=True)
