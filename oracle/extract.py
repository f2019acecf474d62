"""Batch drivers of the CPU oracle (test infrastructure; see ``oracle/__init__.py``).

``run_module`` is the flat-array analogue of the reference's
``extract_*_features(lightcurves, [metadata,] object_ids)`` loops (e.g.
``src/features/statistical.py:135-226``): one row per object that has at least one
observation, in object-index order.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import bazin, colors, gp1d, gp2d, research, physics, powerlaw, shape, statistical, tde
from .common import group_objects

MODULES = ("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d", "gp1d")
_WIDTH = {"statistical": 123, "colors": 83, "shape": 65, "tde": 25, "physics": 32,
          "bazin": 52, "powerlaw": 27, "gp2d": 27, "gp1d": 21, "research": 40, "segments": 22}


def run_module(name: str, obj, band, t, f, e, n_obj: int,
               z: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Returns ``(present, rows)``: indices of objects with rows, and their feature rows."""
    present, rows = [], []
    for k, v in group_objects(obj, band, t, f, e, n_obj):
        if name == "statistical":
            r = statistical.object_row(v)
        elif name == "segments":
            r = statistical.segment_meta(v)
        elif name == "colors":
            r = colors.object_row(v)
        elif name == "shape":
            r = shape.object_row(v)
        elif name == "tde":
            r = tde.object_row(v)
        elif name == "physics":
            r = physics.object_row(v, float("nan") if z is None else float(z[k]))
        elif name == "bazin":
            r = bazin.object_row(v)
        elif name == "powerlaw":
            r = powerlaw.object_row(v)
        elif name == "gp2d":
            r = gp2d.object_row(v)
        elif name == "gp1d":
            r = gp1d.object_row(v)
        elif name == "research":
            r = research.object_row(v, float("nan") if z is None else float(z[k]))
        else:
            raise KeyError(name)
        present.append(k)
        rows.append(r)
    width = _WIDTH[name]
    dtype = np.int64 if name == "segments" else np.float64
    out = np.array(rows, dtype=dtype).reshape(len(rows), width)
    return np.array(present, dtype=np.int64), out


def run_table(name: str, table, z_from_table: bool = True):
    """Convenience wrapper for a ``mallorn_astrophysics_b200.synth.LightCurveTable``-like object."""
    z = table.z if (name in ("physics", "research") and z_from_table) else None
    return run_module(name, table.obj, table.band, table.t, table.flux, table.err, table.n_obj, z)
