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
    g.compile_library(out, [f for f in flags.split(",") if f])
    print("built", out, flush=True)
