"""numpy statement of the packed-store layout (csrc/objview.cuh) for tests: what pack.cu must
produce bit-exactly from a flat long-format table."""
import numpy as np

N_SEG = 7


def pack_reference(obj, band, t, flux, err, n_obj):
    """Returns dict(seg_off int64[n_obj*7+1], t, flux, err float64[n], rank uint16[n], row int64[n])
    ordered by (object, band, time, file rank); rows with obj < 0 are dropped."""
    keep = np.flatnonzero(obj >= 0)
    o, b = obj[keep].astype(np.int64), np.minimum(band[keep], 6).astype(np.int64)
    # file rank inside the object
    order = np.argsort(o, kind="stable")
    so = o[order]
    start = np.searchsorted(so, so, side="left")
    rank = np.empty(len(keep), dtype=np.int64)
    rank[order] = np.arange(len(keep)) - start
    final = np.lexsort((rank, t[keep], b, o))       # last key is primary
    counts = np.bincount(o * N_SEG + b, minlength=n_obj * N_SEG)
    seg_off = np.zeros(n_obj * N_SEG + 1, dtype=np.int64)
    np.cumsum(counts, out=seg_off[1:])
    rows = keep[final]
    # time rank inside the object: position in (time, file rank) order over all bands
    torder = np.lexsort((rank, t[keep], o))
    trank = np.empty(len(keep), dtype=np.int64)
    so2 = o[torder]
    trank[torder] = np.arange(len(keep)) - np.searchsorted(so2, so2, side="left")
    return {"seg_off": seg_off, "t": np.ascontiguousarray(t[rows]), "flux": np.ascontiguousarray(flux[rows]),
            "err": np.ascontiguousarray(err[rows]), "rank": rank[final].astype(np.uint16),
            "trank": trank[final].astype(np.uint16), "row": rows}
