"""Developer script: per-band GP of the TEST-ONLY host instantiation (gp1d.cuh / lbfgsb.cuh) vs oracle/gp1d.py:
features, and the L-BFGS-B iteration / evaluation totals of the four starts of every band."""
import ctypes, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from packref import pack_reference
from hostsim.build import build
from mallorn_astrophysics_b200 import schema
from mallorn_astrophysics_b200.synth import make_table, make_edge_table
from oracle import gp1d
from oracle.common import group_objects

lib = ctypes.CDLL(build())
P = ctypes.c_void_p
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
tables = [("c0", make_table(n, config_index=0)), ("dense", make_table(max(n // 4, 2), seed=424242, count_scale=2.5)), ("edge", make_edge_table())]
warnings.simplefilter("ignore")
for name, tb in tables:
    pk = pack_reference(tb.obj, tb.band, tb.t, tb.flux, tb.err, tb.n_obj)
    out = np.zeros((tb.n_obj, 21)); st = np.zeros((tb.n_obj, 12), np.int32)
    t0 = time.time()
    lib.hs_gp1d(ctypes.c_longlong(tb.n_obj), *[pk[k].ctypes.data_as(P) for k in ("seg_off", "t", "flux", "err", "rank")],
                out.ctypes.data_as(P), st.ctypes.data_as(P))
    t1 = time.time()
    nb = same_counts = same_win = 0; worst = 0.0; nanbad = 0; bad = []
    for k, v in group_objects(tb.obj, tb.band, tb.t, tb.flux, tb.err, tb.n_obj):
        tr = []
        row = np.array(gp1d.object_row(v, tr))
        got = out[k]
        nm = np.isnan(row) != np.isnan(got)
        nanbad += int(nm.sum())
        ok = ~np.isnan(row) & ~nm
        rel = np.abs(got[ok] - row[ok]) / np.maximum(np.abs(row[ok]), 1e-12)
        if rel.size:
            worst = max(worst, rel.max())
            if rel.max() > 1e-3:
                bad.append((k, [schema.GP1D_COLUMNS[j] for j in np.flatnonzero(ok)[rel > 1e-3]]))
        for b, starts in tr:
            j = gp1d.BANDS.index(b)
            if not starts:
                continue
            nb += 1
            nfev = sum(s[1] for s in starts); nit = sum(s[0] for s in starts)
            win = int(np.argmin([s[2] for s in starts]))
            same_counts += (st[k, 3 * j] == nfev and st[k, 3 * j + 1] == nit)
            same_win += (st[k, 3 * j + 2] == win)
    t2 = time.time()
    print(f"{name}: {tb.n_obj} objects, hostsim {t1 - t0:.2f}s oracle {t2 - t1:.1f}s; bands fitted {nb}, identical (nit, nfev) totals {same_counts}, "
          f"same winner {same_win}, NaN mismatches {nanbad}, worst rel {worst:.3g}")
    for b in bad[:8]:
        print("   >1e-3:", b)
