"""Developer script: build named -D variants of the CUDA library into variants/ (select with MLRN_LIB on the GPU box).
    python tests/build_variants.py name=-DFLAG1,-DFLAG2 [name2=...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
os.environ.setdefault("MLRN_QUIET", "1")
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(ROOT, "variants", f"lib_{name}.so")
    fl = [f for f in flags.split(";" if ";" in flags or "{" in flags else ",") if f]
    csrc = g.CSRC
    for f in list(fl):                      # csrc=<dir>: compile another source tree (e.g. an older revision checked out under /tmp)
        if f.startswith("csrc="):
            csrc = f[5:]; fl.remove(f)
    g.compile_library(out, fl, csrc=csrc)
    print("built", out, flush=True)
