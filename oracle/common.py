"""Shared helpers of the CPU oracle (test infrastructure, see ``oracle/__init__.py``).

Objects are handed around as ``ObjectView``: the rows of one object in FILE order,
which is what ``lightcurves.groupby('object_id')`` yields in every reference extractor
(e.g. ``src/features/statistical.py:155``).  Per-band, time-sorted views replicate
``obj_lc[obj_lc['Filter'] == band].sort_values('Time (MJD)')``
(e.g. ``src/features/lightcurve_shape.py:192``); ties in time are broken by file order
(stable sort) -- the reference's default quicksort leaves them unspecified.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, Tuple

import numpy as np

NAN = float("nan")
N_BANDS = 6
U, G, R, I, Z, Y = range(6)


@dataclass
class ObjectView:
    t: np.ndarray      # float64, file order
    f: np.ndarray
    e: np.ndarray
    b: np.ndarray      # uint8 band code, 0..5 ugrizy, 6 unknown

    def band_file(self, band: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        m = self.b == band
        return self.t[m], self.f[m], self.e[m]

    def band_sorted(self, band: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        t, f, e = self.band_file(band)
        o = np.argsort(t, kind="stable")
        return t[o], f[o], e[o]


def group_objects(obj: np.ndarray, band: np.ndarray, t: np.ndarray, f: np.ndarray,
                  e: np.ndarray, n_obj: int) -> Iterator[Tuple[int, ObjectView]]:
    """Yield ``(object_index, ObjectView)`` for every object that has rows, in index order."""
    order = np.argsort(obj, kind="stable")          # keeps file order inside each object
    so = obj[order]
    starts = np.searchsorted(so, np.arange(n_obj), side="left")
    ends = np.searchsorted(so, np.arange(n_obj), side="right")
    bs, ts, fs, es = band[order], t[order], f[order], e[order]
    for k in range(n_obj):
        a, z = starts[k], ends[k]
        if z > a:
            yield k, ObjectView(ts[a:z], fs[a:z], es[a:z], bs[a:z])


def first_argmax(x: np.ndarray) -> int:
    return int(np.argmax(x))


def mag_color(f1: float, f2: float) -> float:
    """-2.5 log10(f1/f2), NaN unless both fluxes are finite and positive (colors.py:92-105)."""
    if np.isnan(f1) or np.isnan(f2) or f1 <= 0 or f2 <= 0:
        return NAN
    return -2.5 * np.log10(f1 / f2)
