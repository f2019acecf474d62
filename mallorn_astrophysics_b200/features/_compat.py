"""Import glue so this package works under both names: ``mallorn_astrophysics_b200.features`` (normal use)
and top-level ``features`` (the drop-in mode of INTEGRATION.md, where ``<repo>/mallorn_astrophysics_b200`` is put
on ``sys.path`` ahead of the reference's ``src`` and its drivers do ``from features.x import ...``)."""
try:
    from .. import schema, store            # imported as mallorn_astrophysics_b200.features
except ImportError:                         # imported as top-level `features`
    import os
    import sys
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    from mallorn_astrophysics_b200 import schema, store   # noqa: F401

LightCurveStore = store.LightCurveStore
