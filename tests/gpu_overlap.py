"""Developer script (GPU): the GP pass (one 512-thread CTA per SM, 46 k of the 64 k registers, FP64 pipes half idle in its
serial sections) next to the fit passes on a second stream.  Needs fit kernels whose CTAs fit beside a GP CTA
(-DFIT_THREADS_N=128 -DFIT_MIN_BLOCKS=4; select with MLRN_LIB).
    python tests/gpu_overlap.py [objects]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200 import _native
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
c = make_columns_device(n, 4242, "cuda")
tag = os.path.basename(os.environ.get("MLRN_LIB", "default"))
with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
    outs = {m: torch.empty((n, _native.COLS[m]), dtype=torch.float64, device="cuda") for m in ("gp2d", "bazin", "powerlaw")}
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    def run(overlap):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if overlap:
            sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
            s.features("gp2d", out=outs["gp2d"], stream=sa.cuda_stream)
            s.features("bazin", out=outs["bazin"], stream=sb.cuda_stream)
            s.features("powerlaw", out=outs["powerlaw"], stream=sb.cuda_stream)
            torch.cuda.current_stream().wait_stream(sa); torch.cuda.current_stream().wait_stream(sb)
        else:
            for m in ("gp2d", "bazin", "powerlaw"):
                s.features(m, out=outs[m])
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)
    for mode in (False, True, False, True):
        ms = run(mode)
        chk = sum(float(torch.nan_to_num(o, nan=0.0, posinf=0.0, neginf=0.0).abs().clamp(max=1e6).sum()) for o in outs.values())
        print(f"[{tag}] {n} objects, gp2d + bazin + powerlaw {'on two streams' if mode else 'one after the other'}: {ms:.1f} ms; abs-sum {chk:.6e}", flush=True)
