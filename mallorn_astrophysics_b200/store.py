"""Host-side handle of the packed, device-resident light-curve store.

``LightCurveStore`` owns one ``mlrn_store`` (C-ABI) -- the CSR (object, band) layout in HBM that
replaces the ``groupby('object_id')`` dict each reference extractor rebuilds
(``src/features/statistical.py:155`` and siblings).  Inputs may be numpy arrays (host; copied
inside the call) or torch CUDA tensors (device-resident; no copy).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native
from .schema import BAND_CODE, UNKNOWN_BAND


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x, dtype):
    """(address, keepalive, on_device) for a numpy array or torch tensor of the given numpy dtype."""
    if _is_torch(x):
        import torch
        want = {np.int32: torch.int32, np.uint8: torch.uint8, np.float64: torch.float64}[dtype]
        if x.dtype != want or not x.is_contiguous():
            x = x.to(want).contiguous()
        return x.data_ptr(), x, x.is_cuda
    a = np.ascontiguousarray(x, dtype=dtype)
    return a.ctypes.data, a, False


def _out_ptr(x, dtype, shape, device: int, what: str):
    """(address, on_device) of an OUTPUT buffer.  Outputs are never converted or copied: anything but a C-contiguous
    array / tensor of exactly this dtype and shape (and, for CUDA tensors, on the store's device) is an error,
    because the kernel would otherwise fill a temporary and leave the caller's buffer unwritten."""
    shape = tuple(int(v) for v in shape)
    if _is_torch(x):
        import torch
        want = {np.int32: torch.int32, np.float64: torch.float64}[dtype]
        if x.dtype != want or not x.is_contiguous() or tuple(x.shape) != shape:
            raise ValueError(f"{what} must be a contiguous {want} tensor of shape {shape}, got {x.dtype} {tuple(x.shape)}"
                             f"{'' if x.is_contiguous() else ' (non-contiguous)'}")
        if x.is_cuda and x.device.index != device:
            raise ValueError(f"{what} lives on cuda:{x.device.index}, the store on cuda:{device}")
        return x.data_ptr(), x.is_cuda
    if not isinstance(x, np.ndarray) or x.dtype != np.dtype(dtype) or not x.flags.c_contiguous or x.shape != shape \
            or not x.flags.writeable:
        raise ValueError(f"{what} must be a writeable C-contiguous {np.dtype(dtype)} array of shape {shape}")
    return x.ctypes.data, False


def frame_columns(lightcurves, object_ids: Optional[Sequence] = None):
    """Host ingest: reference-format long DataFrame -> the flat columns K1 takes.

    ``(obj int32, band uint8, t, flux, err float64, ids)``: ``obj[i]`` is the position of row i's ``object_id`` in
    ``ids`` (the unique requested ids in request order; default: order of first appearance, as
    ``lightcurves['object_id'].unique()``), -1 for rows of objects that were not requested or have no id;
    ``band[i]`` is 0..5 for u g r i z y and ``UNKNOWN_BAND`` for any other ``Filter`` value (NaN included).
    The strings are hashed once -- by ``libmallorn_ingest.so`` on all host cores straight from the Arrow buffers when the
    frame's strings are Arrow-backed (``ingest.factorize_strings``; pandas >= 3 default), by ``pandas.factorize``
    otherwise -- and the per-row work is integer gathers; row order is the
    file order, which K1 preserves inside each object.  Replaces the ``groupby('object_id')`` dict every
    reference extractor builds (``statistical.py:155`` ...) and the ``Filter == band`` masks.
    """
    import pandas as pd
    from . import ingest
    fast = ingest.factorize_strings(lightcurves["object_id"])                    # Arrow-backed strings (pandas >= 3): native, all cores
    row_code, uniq = fast if fast is not None else pd.factorize(lightcurves["object_id"], sort=False)      # -1 for a missing id
    if object_ids is None:
        ids = pd.Index(uniq)
        remap = np.arange(len(uniq), dtype=np.int32)
    else:
        ids = pd.Index(pd.unique(pd.Series(list(object_ids), dtype=object)))
        remap = ids.get_indexer(pd.Index(uniq)).astype(np.int32)                 # -1: present in the frame, not requested
    obj = np.append(remap, np.int32(-1))[row_code]                                # code -1 (no id) -> the appended -1
    band = ingest.band_codes(lightcurves["Filter"])
    if band is None:
        band_code, band_uniq = pd.factorize(lightcurves["Filter"], sort=False)
        lut = np.array([BAND_CODE.get(x, UNKNOWN_BAND) for x in band_uniq] + [UNKNOWN_BAND], dtype=np.uint8)
        band = lut[band_code]                                                     # code -1 (NaN) -> the appended UNKNOWN entry
    return (obj, band,
            lightcurves["Time (MJD)"].to_numpy(dtype=np.float64),
            lightcurves["Flux"].to_numpy(dtype=np.float64),
            lightcurves["Flux_err"].to_numpy(dtype=np.float64), ids)


class LightCurveStore:
    """Packed light curves of ``n_obj`` objects on one GPU."""

    def __init__(self, obj, band, t, flux, err, n_obj: int, device: int = 0, stream: int = 0):
        L = _native.lib()
        _native.require_device()
        n_rows = int(obj.shape[0])
        ptrs = [_ptr(obj, np.int32), _ptr(band, np.uint8), _ptr(t, np.float64),
                _ptr(flux, np.float64), _ptr(err, np.float64)]
        on_dev = {p[2] for p in ptrs}
        if len(on_dev) != 1:
            raise ValueError("input columns must be all host arrays or all CUDA tensors")
        self._h = C.c_void_p()
        self.device, self.stream = int(device), int(stream)
        rc = L.mlrn_store_create(self.device, n_rows, int(n_obj), *[p[0] for p in ptrs],
                                 1 if on_dev.pop() else 0, self.stream, C.byref(self._h))
        _native.check(rc, "mlrn_store_create")
        nr, no, mo = C.c_int64(), C.c_int32(), C.c_int32()
        _native.check(L.mlrn_store_info(self._h, C.byref(nr), C.byref(no), C.byref(mo)), "mlrn_store_info")
        self.n_rows, self.n_obj, self.max_obs = nr.value, no.value, mo.value

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_table(cls, table, device: int = 0, stream: int = 0) -> "LightCurveStore":
        return cls(table.obj, table.band, table.t, table.flux, table.err, table.n_obj, device, stream)

    @classmethod
    def from_frame(cls, lightcurves, object_ids: Optional[Sequence] = None, device: int = 0, stream: int = 0):
        """Factorise a reference-format DataFrame (``object_id, Time (MJD), Flux, Flux_err, Filter``).

        Returns ``(store, ids)`` where ``ids`` are the unique requested ids in request order; rows
        of other objects are skipped, as the reference's ``grouped.get(obj_id)`` never sees them.
        """
        codes, band, t, flux, err, ids = frame_columns(lightcurves, object_ids)
        return cls(codes, band, t, flux, err, len(ids), device, stream), ids

    @classmethod
    def from_csv(cls, paths, device: int = 0, stream: int = 0, threads: int = 0):
        """Light-curve CSV file(s) of the reference's layout -> packed store, without a DataFrame: the native host
        ingest (``ingest.read_lightcurve_csv``: threads over memory-mapped chunks, pandas-exact number conversion)
        feeds K1 directly.  Returns ``(store, ids)`` with ``ids`` in order of first appearance."""
        from . import ingest
        cols = ingest.read_lightcurve_csv(paths, threads)
        return cls(cols.obj, cols.band, cols.t, cols.flux, cols.err, cols.n_obj, device, stream), cols.object_ids

    # ------------------------------------------------------------------ K1 outputs
    def segments(self) -> np.ndarray:
        """int32 [n_obj, 22]: per segment (count, file-order peak pos, time-order peak pos), peak_band."""
        out = np.empty((self.n_obj, _native.SEGMETA_COLS), dtype=np.int32)
        _native.check(_native.lib().mlrn_store_segments(self._h, out.ctypes.data), "mlrn_store_segments")
        return out

    def export(self) -> dict:
        seg = np.empty(self.n_obj * _native.N_SEG + 1, dtype=np.int64)
        t, f, e = (np.empty(self.n_rows, dtype=np.float64) for _ in range(3))
        rank = np.empty(self.n_rows, dtype=np.uint16)
        trank = np.empty(self.n_rows, dtype=np.uint16)
        _native.check(_native.lib().mlrn_store_export(self._h, seg.ctypes.data, t.ctypes.data, f.ctypes.data,
                                                      e.ctypes.data, rank.ctypes.data, trank.ctypes.data), "mlrn_store_export")
        return {"seg_off": seg, "t": t, "flux": f, "err": e, "rank": rank, "trank": trank}

    def present(self) -> np.ndarray:
        """Indices of objects that have at least one observation."""
        return np.flatnonzero(self.segments()[:, 0:21:3].sum(axis=1) > 0)

    # ------------------------------------------------------------------ feature passes
    def features(self, module: str, z=None, out=None, status=None, stream: Optional[int] = None):
        """Run one pass.  ``out``: None (returns a new numpy array), a numpy array, or a CUDA
        tensor [n_obj, cols] float64 (asynchronous on ``stream``).  ``status``: optional int32
        numpy array / CUDA tensor for the fit passes (see include/mallorn_b200.h)."""
        L = _native.lib()
        cols = _native.COLS[module]
        st = self.stream if stream is None else int(stream)
        ret = None
        if out is None:
            out = ret = np.empty((self.n_obj, cols), dtype=np.float64)
        optr, odev = _out_ptr(out, np.float64, (self.n_obj, cols), self.device, "out")
        scols = {"bazin": 12, "powerlaw": 54, "gp2d": 4, "gp1d": 12}.get(module)
        sptr = 0
        if status is not None:
            if scols is None:
                raise ValueError(f"the '{module}' pass has no status table")
            n_stat = int(status.numel() if _is_torch(status) else status.size)
            if n_stat != self.n_obj * scols:
                raise ValueError(f"status must hold {scols} int32 values per object ({self.n_obj * scols}), got {n_stat}")
            sptr, sdev = _out_ptr(status, np.int32, tuple(status.shape), self.device, "status")
            if sdev != odev:
                raise ValueError("status and out must live on the same side")
        if module in ("physics", "research"):
            if z is None:
                z = np.full(self.n_obj, np.nan)
            zptr, zkeep, zdev = _ptr(z, np.float64)
            fn = L.mlrn_physics if module == "physics" else L.mlrn_research
            rc = fn(self._h, zptr, 1 if zdev else 0, optr, 1 if odev else 0, st)
        elif module in ("bazin", "powerlaw", "gp2d", "gp1d"):
            fn = getattr(L, {"bazin": "mlrn_bazin", "powerlaw": "mlrn_powerlaw", "gp2d": "mlrn_gp2d",
                              "gp1d": "mlrn_gp1d"}[module])
            rc = fn(self._h, optr, sptr, 1 if odev else 0, st)
        else:
            fn = getattr(L, {"statistical": "mlrn_stat", "colors": "mlrn_color", "shape": "mlrn_shape",
                             "tde": "mlrn_tde"}[module])
            rc = fn(self._h, optr, 1 if odev else 0, st)
        _native.check(rc, f"mlrn pass '{module}'")
        return ret if ret is not None else out

    def fit_passes(self, out_bazin, out_powerlaw, status_bazin=None, status_powerlaw=None, stream: Optional[int] = None):
        """Bazin and decline-model passes in one call (``mlrn_fits``): the decline models run on an internal side stream
        so that the straggler stages of the two passes overlap.  Outputs as in :meth:`features` (both on the same side);
        results are those of the two separate passes."""
        L = _native.lib()
        st = self.stream if stream is None else int(stream)
        bptr, bdev = _out_ptr(out_bazin, np.float64, (self.n_obj, _native.COLS["bazin"]), self.device, "out_bazin")
        pptr, pdev = _out_ptr(out_powerlaw, np.float64, (self.n_obj, _native.COLS["powerlaw"]), self.device, "out_powerlaw")
        if bdev != pdev:
            raise ValueError("out_bazin and out_powerlaw must live on the same side")
        sp = []
        for status, scols, name in ((status_bazin, 12, "status_bazin"), (status_powerlaw, 54, "status_powerlaw")):
            if status is None:
                sp.append(0)
                continue
            n_stat = int(status.numel() if _is_torch(status) else status.size)
            if n_stat != self.n_obj * scols:
                raise ValueError(f"{name} must hold {scols} int32 values per object ({self.n_obj * scols}), got {n_stat}")
            ptr, sdev = _out_ptr(status, np.int32, tuple(status.shape), self.device, name)
            if sdev != bdev:
                raise ValueError(f"{name} and the outputs must live on the same side")
            sp.append(ptr)
        _native.check(L.mlrn_fits(self._h, bptr, sp[0], pptr, sp[1], 1 if bdev else 0, st), "mlrn_fits")
        return out_bazin, out_powerlaw

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            _native.lib().mlrn_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
