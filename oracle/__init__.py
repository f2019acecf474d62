"""CPU oracle for the MALLORN light-curve featurisation path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy/scipy on flat arrays, the arithmetic of the
reference's per-object feature extractors so the CUDA path can be checked on the GPU
box (where ``/root/reference`` does not exist).  Nothing in the product package
``mallorn_astrophysics_b200`` imports it; the only callers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.

Parity pin
----------
* statistical, colors, shape, tde_physics, physics_based, bazin, power-law:
  **pinned** -- ``tests/golden/make_golden.py`` imports the unmodified reference
  modules from ``/root/reference/src`` (and exec's ``scripts/train_v55_powerlaw.py:106-202``)
  on a seeded synthetic set and commits their outputs under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` holds this oracle to those frames.
* 2-D multi-band GP (``oracle/gp2d.py``): **parity unpinned**.  The reference calls
  ``george>=0.4`` (requirements.txt:16), which is neither vendored under
  ``/root/reference`` nor installed here; the restatement follows george 0.4's
  published semantics (assumption A-george) and the reference's
  own call sites ``src/features/multiband_gp.py:34-344``.

Third-party numerics used exactly as the reference does: ``scipy.optimize.curve_fit``
(bounded -> ``least_squares(method='trf')``), ``scipy.optimize.minimize('L-BFGS-B')``,
``numpy.polyfit / percentile / median``.
"""
from .common import group_objects, ObjectView  # noqa: F401
