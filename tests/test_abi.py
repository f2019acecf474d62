"""The C-ABI library must load and export every symbol include/mallorn_b200.h declares
(no compute: this runs without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g
    g.build()
    from mallorn_astrophysics_b200 import _native
    return _native


def test_header_symbols_are_exported(native):
    hdr = open(os.path.join(ROOT, "include", "mallorn_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mlrn_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 10
    L = native.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert L.mlrn_version() >= 100


def test_no_cpu_fallback(native):
    """Without a device the product path must fail loudly, not compute on the host."""
    import numpy as np
    if native.lib().mlrn_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from mallorn_astrophysics_b200.store import LightCurveStore
    with pytest.raises(native.MallornNativeError):
        LightCurveStore(np.zeros(1, np.int32), np.zeros(1, np.uint8), np.zeros(1), np.zeros(1), np.ones(1), 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mallorn_astrophysics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "hostsim" not in src, f


def test_schema_matches_header(native):
    from mallorn_astrophysics_b200 import schema
    hdr = open(os.path.join(ROOT, "include", "mallorn_b200.h")).read()
    for mod, macro in (("statistical", "STAT"), ("colors", "COLOR"), ("shape", "SHAPE"), ("tde", "TDE"),
                       ("physics", "PHYSICS"), ("bazin", "BAZIN"), ("powerlaw", "POWERLAW"), ("gp2d", "GP2D"), ("gp1d", "GP1D"), ("research", "RESEARCH")):
        n = int(re.search(rf"#define MLRN_{macro}_COLS (\d+)", hdr).group(1))
        assert n == len(schema.MODULE_COLUMNS[mod]) == native.COLS[mod]


def test_features_package_imports_under_both_names():
    """the reference's drivers import `features.*` after one sys.path line; the same modules are
    also importable as mallorn_astrophysics_b200.features.  (Import only -- no compute without a GPU.)"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); "
            "from features.statistical import extract_statistical_features, add_metadata_features, LSST_BANDS; "
            "from features.colors import extract_color_features; from features.lightcurve_shape import extract_shape_features; "
            "from features.physics_based import extract_physics_features; from features.tde_physics import extract_tde_physics_features; "
            "from features.bazin_fitting import extract_bazin_features, bazin_function; "
            "from features.multiband_gp import extract_multiband_gp_features; from features.gaussian_process import extract_gp_features; from features.research_features import extract_research_features; from features.powerlaw import extract_powerlaw_features; "
            "import features; assert features.extract_statistical_features is extract_statistical_features; "
            "assert LSST_BANDS == list('ugrizy'); print('ok')") % os.path.join(ROOT, "mallorn_astrophysics_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr
    import mallorn_astrophysics_b200.features as f2
    assert callable(f2.extract_color_features)
