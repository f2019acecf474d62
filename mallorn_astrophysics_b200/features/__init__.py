"""Drop-in replacement of the reference's ``src/features`` package for the hot-path modules.

Put the directory that contains this package ahead of the reference's ``src`` on ``sys.path``
(``sys.path.insert(0, '<repo>/mallorn_astrophysics_b200')``) and the reference's drivers
(``scripts/precompute_features.py``, ``cache_bazin_features.py``, ``train_v34a_bazin.py`` ...)
import these extractors unchanged; see INTEGRATION.md.  Re-exports mirror
``src/features/__init__.py:1-4``.
"""
from .statistical import extract_statistical_features  # noqa: F401
from .colors import extract_color_features  # noqa: F401
from .lightcurve_shape import extract_shape_features  # noqa: F401
from .physics_based import extract_physics_features  # noqa: F401
