"""Host ingest of the reference's light-curve files without a DataFrame (``libmallorn_ingest.so``).

``read_lightcurve_csv(paths)`` is the columnar counterpart of ``utils.data_loader.load_lightcurves``
(``src/utils/data_loader.py:36-62``): the split files are memory-mapped and parsed by a pool of threads straight
into the flat columns K1 takes (``obj`` int32 numbered by first appearance, ``band`` uint8, ``t / flux / err``
float64, file order) plus the id table -- bit-identical to ``pd.concat([pd.read_csv(p) ...])`` followed by
``store.frame_columns`` (``tests/test_ingest_csv.py``), at several times pandas' speed per core and on all cores.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence, Union

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmallorn_ingest.so")
_lib: Optional[C.CDLL] = None


class MallornIngestError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MallornIngestError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.mlrn_ingest_last_error.restype = C.c_char_p
        L.mlrn_ingest_csv.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(vp)]
        L.mlrn_table_rows.restype = C.c_int64
        L.mlrn_table_objects.restype = C.c_int32
        L.mlrn_table_id_bytes.restype = C.c_int64
        for name in ("mlrn_table_obj", "mlrn_table_band", "mlrn_table_time", "mlrn_table_flux", "mlrn_table_err",
                     "mlrn_table_id_offsets", "mlrn_table_id_chars"):
            getattr(L, name).restype = vp
        for name in ("mlrn_table_rows", "mlrn_table_objects", "mlrn_table_id_bytes", "mlrn_table_obj", "mlrn_table_band",
                     "mlrn_table_time", "mlrn_table_flux", "mlrn_table_err", "mlrn_table_id_offsets", "mlrn_table_id_chars",
                     "mlrn_table_free"):
            getattr(L, name).argtypes = [vp]
        L.mlrn_table_free.restype = None
        pp = C.POINTER(vp)
        chunk_args = [C.c_int, pp, pp, pp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.c_int]
        L.mlrn_factorize_utf8.argtypes = chunk_args + [vp, vp, C.POINTER(C.c_int32)]
        L.mlrn_band_codes_utf8.argtypes = chunk_args + [vp]
        _lib = L
    return _lib


@dataclass
class LightCurveColumns:
    """Flat long-format light curves in file order (the inputs of ``LightCurveStore``)."""
    obj: np.ndarray          # int32 [rows], position in ``object_ids``; -1 = row without an id
    band: np.ndarray         # uint8 [rows], 0..5 = u g r i z y, 6 = other
    t: np.ndarray            # float64 [rows]
    flux: np.ndarray
    err: np.ndarray
    object_ids: np.ndarray   # object [n_obj], order of first appearance (== DataFrame['object_id'].unique())

    @property
    def n_obj(self) -> int:
        return len(self.object_ids)

    @property
    def n_rows(self) -> int:
        return len(self.obj)

    def to_frame(self):
        """The reference's long DataFrame (for code that still wants one)."""
        import pandas as pd
        bands = np.array(["u", "g", "r", "i", "z", "y", None], dtype=object)
        ids = np.append(self.object_ids, None)
        return pd.DataFrame({"object_id": ids[self.obj], "Time (MJD)": self.t, "Flux": self.flux,
                             "Flux_err": self.err, "Filter": bands[self.band]})


class _Table:
    """Owns one ``mlrn_table``; freed when the last array that views its columns is garbage-collected."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        if self.handle:
            lib().mlrn_table_free(self.handle)
            self.handle = None


def _view(owner: _Table, ptr, n: int, dtype) -> np.ndarray:
    """Zero-copy numpy view of a column of the table; keeps ``owner`` (and so the C memory) alive."""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (np.ctypeslib.as_ctypes_type(dtype) * n).from_address(ptr)
    buf._owner = owner                               # ctypes instances accept attributes; numpy keeps `buf` as .base
    return np.frombuffer(buf, dtype=dtype)


def _arrow_string_chunks(column):
    """The Arrow arrays behind a pandas string column (``str`` / ``string[pyarrow]`` / ``ArrowDtype(string)``), or ``None``
    when the column is not Arrow-backed UTF-8 (object dtype, python-backed strings, categoricals ...)."""
    arr = getattr(column, "array", column)
    if not hasattr(arr, "__arrow_array__"):
        return None
    dt = getattr(arr, "dtype", None)
    if getattr(dt, "storage", None) != "pyarrow" and type(dt).__name__ != "ArrowDtype":
        return None                                            # python-backed strings would be converted (copied) by the protocol
    try:
        import pyarrow as pa
        pa_arr = arr.__arrow_array__()
    except Exception:
        return None
    chunks = list(pa_arr.chunks) if isinstance(pa_arr, pa.ChunkedArray) else [pa_arr]
    if chunks and all(pa.types.is_large_string(c.type) for c in chunks):
        width = 8
    elif chunks and all(pa.types.is_string(c.type) for c in chunks):
        width = 4
    elif chunks:
        return None
    else:
        width = 8
    return chunks, width


def _chunk_args(chunks, width: int, threads: int):
    n = len(chunks)
    vp = C.c_void_p
    offs, data, valid = (vp * max(n, 1))(), (vp * max(n, 1))(), (vp * max(n, 1))()
    aoff, length = (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))()
    for i, c in enumerate(chunks):
        v, o, d = c.buffers()
        offs[i] = o.address if o is not None else None
        data[i] = d.address if d is not None and d.size else None
        valid[i] = v.address if v is not None and c.null_count else None
        aoff[i], length[i] = c.offset, len(c)
    return [n, offs, data, valid, aoff, length, width, int(threads)]


def factorize_strings(column, threads: int = 0):
    """``pandas.factorize(column, sort=False)`` for an Arrow-backed string column, on all host cores and without creating
    a Python object per row: ``(codes int32 [rows] (-1 = null), uniques)`` with ``uniques`` the distinct strings in order of
    first appearance (an object ndarray).  Returns ``None`` when the column is not Arrow-backed (the caller then uses pandas)."""
    got = _arrow_string_chunks(column)
    if got is None:
        return None
    chunks, width = got
    rows = sum(len(c) for c in chunks)
    codes = np.empty(rows, dtype=np.int32)
    first = np.empty(rows, dtype=np.int64)
    n_unique = C.c_int32(0)
    L = lib()
    rc = L.mlrn_factorize_utf8(*_chunk_args(chunks, width, threads), codes.ctypes.data, first.ctypes.data, C.byref(n_unique))
    if rc != 0:
        raise MallornIngestError(f"mlrn_factorize_utf8 failed ({rc}): {L.mlrn_ingest_last_error().decode()}")
    import pyarrow as pa
    whole = pa.chunked_array(chunks) if chunks else pa.chunked_array([], type=pa.large_string())
    uniques = whole.take(pa.array(first[: n_unique.value])).to_numpy(zero_copy_only=False) if n_unique.value else np.empty(0, dtype=object)
    return codes, np.asarray(uniques, dtype=object)


def band_codes(column, threads: int = 0):
    """``Filter`` strings -> uint8 band codes (0..5 = u g r i z y, 6 = anything else or null) for an Arrow-backed string
    column; ``None`` when the column is not Arrow-backed."""
    got = _arrow_string_chunks(column)
    if got is None:
        return None
    chunks, width = got
    band = np.empty(sum(len(c) for c in chunks), dtype=np.uint8)
    L = lib()
    rc = L.mlrn_band_codes_utf8(*_chunk_args(chunks, width, threads), band.ctypes.data)
    if rc != 0:
        raise MallornIngestError(f"mlrn_band_codes_utf8 failed ({rc}): {L.mlrn_ingest_last_error().decode()}")
    return band


def read_lightcurve_csv(paths: Union[str, Path, Sequence[Union[str, Path]]], threads: int = 0) -> LightCurveColumns:
    """Parse one or more light-curve CSV files as one concatenated table.  Structural problems (missing column,
    ragged line, non-numeric field, unreadable file) raise ``MallornIngestError``.  The returned arrays are views of the
    parser's own buffers (no copy); they stay valid for as long as any of them is referenced."""
    if isinstance(paths, (str, Path)):
        paths = [paths]
    L = lib()
    enc = [os.fsencode(str(p)) for p in paths]
    arr = (C.c_char_p * len(enc))(*enc)
    h = C.c_void_p()
    rc = L.mlrn_ingest_csv(arr, len(enc), int(threads), C.byref(h))
    if rc != 0:
        raise MallornIngestError(f"mlrn_ingest_csv failed (code {rc}): {L.mlrn_ingest_last_error().decode()}")
    owner = _Table(h)
    n, k = L.mlrn_table_rows(h), L.mlrn_table_objects(h)
    off = _view(owner, L.mlrn_table_id_offsets(h), k + 1, np.int64)
    chars = C.string_at(L.mlrn_table_id_chars(h), L.mlrn_table_id_bytes(h)) if k else b""
    ids = np.empty(k, dtype=object)
    for i in range(k):
        ids[i] = chars[off[i]:off[i + 1]].decode()
    return LightCurveColumns(_view(owner, L.mlrn_table_obj(h), n, np.int32), _view(owner, L.mlrn_table_band(h), n, np.uint8),
                             _view(owner, L.mlrn_table_time(h), n, np.float64), _view(owner, L.mlrn_table_flux(h), n, np.float64),
                             _view(owner, L.mlrn_table_err(h), n, np.float64), ids)


def load_lightcurves(split: str = "train", data_path: Union[str, Path, None] = None, threads: int = 0) -> LightCurveColumns:
    """``utils.data_loader.load_lightcurves`` (data_loader.py:36-62) without the DataFrame: every existing
    ``split_XX/{split}_full_lightcurves.csv`` for XX = 01..20, concatenated in that order."""
    if data_path is None:
        raise ValueError("data_path is required (the reference defaults to <its repo>/data/raw)")
    data_path = Path(data_path)
    files = [data_path / f"split_{i:02d}" / f"{split}_full_lightcurves.csv" for i in range(1, 21)]
    files = [f for f in files if f.exists()]
    if not files:
        raise FileNotFoundError(f"No {split} lightcurve files found")
    return read_lightcurve_csv(files, threads)


def featurise_files(paths: Sequence[Union[str, Path]], modules: Sequence[str], metadata=None, device: int = 0, threads: int = 0,
                    chunk_bytes: int = 64 << 20):
    """Files -> feature table as an overlapped pipeline (``data_loader.py:36-62`` + every extractor's
    ``groupby``).  The files are taken in groups of about ``chunk_bytes`` (one native-reader call each):

        parse group i+1 (native reader, all host threads, GIL released)     |  producer thread
        merge its ids into the global table, re-base ``obj`` and start the host->device copies of its five columns on
        a copy stream                                                        |  this thread, while the next group parses

    and, once the last group is in HBM, ONE K1 pack over the concatenated device columns followed by the passes -- an
    object's rows may be spread over several files, so segmenting happens after the last file, on the GPU, where it costs
    milliseconds.  Wall time is then about max(parse, copies) + GPU instead of their sum.  (Measured on the B200 box: the
    reader parses 3.5 GB/s = 49 M rows/s on 16 threads while the passes consume 11 M rows/s, so parsing is at most a
    fifth of the wall time to begin with and small sets gain nothing; the 109 MB C1 set is two groups.)

    Returns ``(table, object_ids, info)``: the ``[n_obj, width]`` float64 CUDA table (module columns side by side, ids in
    order of first appearance, exactly as ``read_lightcurve_csv(paths)`` numbers them) and the split of the wall time.
    ``metadata``: optional frame with ``object_id`` and ``Z`` for the physics / research passes."""
    import queue
    import threading
    import time
    import pandas as pd
    import torch
    from . import shard
    t_start = time.perf_counter()
    paths = [str(p) for p in paths]
    groups, cur, size = [], [], 0
    for p in paths:
        cur.append(p)
        size += os.path.getsize(p)
        if size >= chunk_bytes:
            groups.append(cur)
            cur, size = [], 0
    if cur:
        groups.append(cur)
    dev = torch.device("cuda", int(device))
    q: "queue.Queue" = queue.Queue(maxsize=2)

    def producer():
        try:
            for grp in groups:
                t0 = time.perf_counter()
                cols = read_lightcurve_csv(grp, threads)
                q.put((cols, time.perf_counter() - t0))
        except BaseException as exc:                     # handed to the consumer
            q.put(exc)
        q.put(None)

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    copy_stream = torch.cuda.Stream(device=dev)
    known = pd.Index([], dtype=object)                   # global id table, order of first appearance over all files
    chunks = []
    parse_s = stage_s = 0.0
    n_rows = 0
    while True:
        item = q.get()
        if item is None:
            break
        if isinstance(item, BaseException):
            raise item
        cols, dt = item
        parse_s += dt
        t0 = time.perf_counter()
        lut = known.get_indexer(pd.Index(cols.object_ids, dtype=object)).astype(np.int32)
        new = lut < 0
        if new.any():
            lut[new] = len(known) + np.arange(int(new.sum()), dtype=np.int32)
            known = known.append(pd.Index(cols.object_ids[new], dtype=object))
        lut = np.append(lut, np.int32(-1))               # rows without an id (-1) pick the appended entry
        with torch.cuda.stream(copy_stream):
            # the parser's buffers are pageable: the driver stages them through its own pinned ring, which for chunks of
            # tens of MB costs less than pinning a fresh buffer per chunk did (0.88 s for twenty 5 MB chunks)
            dcols = {"obj": torch.from_numpy(lut[cols.obj]).to(dev, non_blocking=True)}
            for k in ("band", "t", "flux", "err"):
                dcols[k] = torch.from_numpy(getattr(cols, k)).to(dev, non_blocking=True)
        chunks.append((dcols, cols))                     # keep the host arrays alive until the copies are done
        n_rows += cols.n_rows
        stage_s += time.perf_counter() - t0
    th.join()
    t0 = time.perf_counter()
    torch.cuda.current_stream(dev).wait_stream(copy_stream)
    if chunks:
        flat = {k: (torch.cat([c[0][k] for c in chunks]) if len(chunks) > 1 else chunks[0][0][k]) for k in ("obj", "band", "t", "flux", "err")}
    else:
        flat = {"obj": torch.empty(0, dtype=torch.int32, device=dev), "band": torch.empty(0, dtype=torch.uint8, device=dev),
                **{k: torch.empty(0, dtype=torch.float64, device=dev) for k in ("t", "flux", "err")}}
    object_ids = np.asarray(known, dtype=object)
    if metadata is not None and any(m in ("physics", "research") for m in modules):
        zmap = pd.Series(metadata["Z"].to_numpy(dtype=np.float64), index=metadata["object_id"].to_numpy())
        zmap = zmap[~zmap.index.duplicated(keep="last")]
        flat["z"] = zmap.reindex(object_ids).to_numpy(dtype=np.float64)
    with torch.cuda.device(dev):
        table = shard.featurise_shard(flat, len(object_ids), modules, device=int(device))
    torch.cuda.synchronize(dev)
    gpu_s = time.perf_counter() - t0
    info = {"files": len(paths), "groups": len(groups), "rows": int(n_rows), "objects": len(object_ids), "parse_seconds": parse_s,
            "stage_and_copy_issue_seconds": stage_s, "pack_and_passes_seconds": gpu_s,
            "wall_seconds": time.perf_counter() - t_start, "modules": list(modules)}
    return table, object_ids, info
