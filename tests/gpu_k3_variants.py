"""Developer script (GPU): the K3 scalar passes (and the research block, which shares team.cuh) of several builds of the
library in ONE process -- time per pass on the C4 generator and bit-comparison of every row with the first build's.
    python tests/gpu_k3_variants.py objects tag=path[:MLRN_K3_TIGHT] ..."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from mallorn_astrophysics_b200 import _native
from mallorn_astrophysics_b200.store import LightCurveStore
from mallorn_astrophysics_b200.synth import make_columns_device

n = int(sys.argv[1])
c = make_columns_device(n, 4242, "cuda")
mods = ("colors", "shape", "tde", "physics", "research")
ref = {}
for spec in sys.argv[2:]:
    tag, rest = spec.split("=", 1)
    path, _, tight = rest.partition(":")
    os.environ["MLRN_K3_TIGHT"] = tight or "1"
    _native._lib, _native.LIB_PATH = None, os.path.abspath(path)
    line = []
    with LightCurveStore(c["obj"], c["band"], c["t"], c["flux"], c["err"], n) as s:
        for m in mods:
            out = torch.empty((n, _native.COLS[m]), dtype=torch.float64, device="cuda")
            best = 1e30
            for r in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); s.features(m, z=c["z"] if m in ("physics", "research") else None, out=out); b.record()
                torch.cuda.synchronize()
                if r: best = min(best, a.elapsed_time(b))
            bits = out.view(torch.int64)
            if m not in ref:
                ref[m] = bits.clone()
                same = "ref"
            else:
                diff = int((bits != ref[m]).sum())
                close = bool(torch.allclose(out, ref[m].view(torch.float64), rtol=1e-12, atol=0.0, equal_nan=True))
                same = "bit-identical" if diff == 0 else f"{diff} cells differ (allclose 1e-12: {close})"
            line.append(f"{m} {best:.2f} ms [{same}]")
    print(f"[{tag}] " + "; ".join(line), flush=True)
