"""Developer script (CPU): throughput of the native CSV ingest against the reference's pandas path on a C1-sized set
(10,178 objects, ~1.5 M rows, 20 split files).    python tests/ingest_rate.py [objects] [dir]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from mallorn_astrophysics_b200 import ingest
from mallorn_astrophysics_b200.store import frame_columns
from mallorn_astrophysics_b200.synth import make_table

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10178
root = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp(prefix="mlrn_ingest_")
paths = [os.path.join(root, f"split_{i:02d}", "train_full_lightcurves.csv") for i in range(1, 21)]
if not os.path.exists(paths[0]):
    lc = make_table(n, config_index=1).to_frame()
    edges = np.linspace(0, len(lc), 21).astype(int)
    for i, p in enumerate(paths):
        os.makedirs(os.path.dirname(p), exist_ok=True)
        lc.iloc[edges[i]:edges[i + 1]].to_csv(p, index=False)
size = sum(os.path.getsize(p) for p in paths)
t0 = time.perf_counter()
lc = pd.concat([pd.read_csv(p) for p in paths], ignore_index=True)      # data_loader.py:53-62
t1 = time.perf_counter()
ref = frame_columns(lc, None)
t2 = time.perf_counter()
print(f"{len(lc)} rows, {size / 1e6:.1f} MB in 20 files")
print(f"pandas: read_csv+concat {t1 - t0:.2f} s, factorise {t2 - t1:.2f} s -> {size / 1e6 / (t2 - t0):.0f} MB/s")
for th in (1, 4, 0):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); got = ingest.read_lightcurve_csv(paths, threads=th); best = min(best, time.perf_counter() - t0)
    ok = np.array_equal(got.obj, ref[0]) and np.array_equal(got.band, ref[1]) and all(
        np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in ((got.t, ref[2]), (got.flux, ref[3]), (got.err, ref[4])))
    print(f"native, threads={th or os.cpu_count()}: {best:.3f} s -> {size / 1e6 / best:.0f} MB/s, {len(lc) / best / 1e6:.1f} M rows/s, identical to pandas: {ok}")
