"""Developer script: TEST-ONLY host instantiation of the fit templates vs the reference goldens / oracle."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from packref import pack_reference
from hostsim.build import build
from mallorn_astrophysics_b200 import schema
lib = ctypes.CDLL(build())
P = ctypes.c_void_p

def run(fn, ncol, nstat, pk, n_obj):
    out = np.zeros((n_obj, ncol)); st = np.zeros((n_obj, nstat, 2), np.int32)
    t0 = time.time()
    getattr(lib, fn)(ctypes.c_longlong(n_obj), pk["seg_off"].ctypes.data_as(P), pk["t"].ctypes.data_as(P),
                     pk["flux"].ctypes.data_as(P), pk["err"].ctypes.data_as(P), pk["rank"].ctypes.data_as(P),
                     out.ctypes.data_as(P), st.ctypes.data_as(P))
    return out, st, time.time() - t0

if __name__ == "__main__":
    mods = sys.argv[1:] or ["bazin", "powerlaw"]
    for name in ("ref_sparse", "ref_dense", "ref_edge"):
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz"))
        n_obj = int(g["n_obj"])
        pk = pack_reference(g["obj"], g["band"], g["t"], g["flux"], g["err"], n_obj)
        for m in mods:
            fn, nc, ns = {"bazin": ("hs_bazin_t", 52, 6), "powerlaw": ("hs_powerlaw_t", 27, 27)}[m]
            out, st, dt = run(fn, nc, ns, pk, n_obj)
            ref, pres = g[m + "_rows"], g[m + "_present"]
            got = out[pres]
            if name == "ref_edge":
                keep = pres != 15
                got, ref, pres = got[keep], ref[keep], pres[keep]
            nm = np.isnan(got) != np.isnan(ref)
            ok = ~np.isnan(ref) & ~nm
            rel = np.where(ok, np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9), 0)
            print(f"{name} {m}: {dt:.2f}s nan mismatch {int(nm.sum())} of {nm.size}; rel>1e-3: {int((rel > 1e-3).sum())}; "
                  f"rel>1e-6: {int((rel > 1e-6).sum())}; max rel {rel.max():.3g}")
            cols = schema.MODULE_COLUMNS[m]
            for i, j in np.argwhere(nm | (rel > 1e-3))[:8]:
                print("    obj", pres[i], cols[j], got[i, j], ref[i, j], "status", st[pres[i]].reshape(-1)[:12] if m == "bazin" else st[pres[i], j])
