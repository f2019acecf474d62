"""Developer script: run the TEST-ONLY host instantiation of the device templates against the goldens."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from packref import pack_reference
from hostsim.build import build
from mallorn_astrophysics_b200 import schema
lib = ctypes.CDLL(build())
P = ctypes.c_void_p
COLS = {"colors": ("hs_color", 83), "shape": ("hs_shape", 65), "tde": ("hs_tde", 25), "physics": ("hs_physics", 32)}
mods = sys.argv[1:] or list(COLS)
for name in ("ref_sparse", "ref_dense", "ref_edge"):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz"))
    n_obj = int(g["n_obj"])
    pk = pack_reference(g["obj"], g["band"], g["t"], g["flux"], g["err"], n_obj)
    z = np.ascontiguousarray(g["z"])
    for m in mods:
        fn, nc = COLS[m]
        out = np.zeros((n_obj, nc))
        getattr(lib, fn)(ctypes.c_longlong(n_obj), pk["seg_off"].ctypes.data_as(P), pk["t"].ctypes.data_as(P),
                         pk["flux"].ctypes.data_as(P), pk["err"].ctypes.data_as(P), pk["rank"].ctypes.data_as(P),
                         z.ctypes.data_as(P), out.ctypes.data_as(P))
        ref, pres = g[m + "_rows"], g[m + "_present"]
        got = out[pres]
        if name == "ref_edge":
            keep = pres != 15
            got, ref, pres = got[keep], ref[keep], pres[keep]
        nm = np.isnan(got) != np.isnan(ref)
        ok = ~np.isnan(ref) & ~nm
        rel = np.where(ok, np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9), 0)
        print(name, m, "nan mismatch", int(nm.sum()), "max rel %.3g" % rel.max())
        cols = schema.MODULE_COLUMNS[m]
        for i, j in np.argwhere(nm | (rel > 1e-7))[:6]:
            print("    obj", pres[i], cols[j], got[i, j], ref[i, j])
