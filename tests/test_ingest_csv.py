"""Native host ingest (libmallorn_ingest.so, csrc/ingest.cpp) against the reference's own loading path:
``pd.read_csv`` of every split file + ``pd.concat`` (src/utils/data_loader.py:36-62), then the factorisation
(store.frame_columns).  Bit-exact: ids and their order, dense object index, band codes, and every float -- which
means reproducing pandas' default number converter, not a correctly rounded strtod."""
import io
import os

import numpy as np
import pandas as pd
import pytest

import __graft_entry__ as entry
from mallorn_astrophysics_b200.store import frame_columns


@pytest.fixture(scope="module")
def ingest():
    entry.build()
    from mallorn_astrophysics_b200 import ingest as ing
    return ing


def _write_splits(tmp_path, n_obj=300, seed=5, n_files=3, extra_column=True):
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(n_obj, seed=seed)
    lc = tb.to_frame()
    rng = np.random.default_rng(seed)
    lc.loc[rng.choice(len(lc), 40, replace=False), "Flux"] = np.nan          # empty fields
    lc.loc[rng.choice(len(lc), 25, replace=False), "Filter"] = "w"           # a filter outside ugrizy
    lc.loc[rng.choice(len(lc), 7, replace=False), "Flux_err"] = 0.0
    lc.loc[rng.choice(len(lc), 5, replace=False), "Time (MJD)"] = 60000      # integer-looking values
    if extra_column:
        lc.insert(2, "detected", rng.integers(0, 2, len(lc)))
    edges = np.linspace(0, len(lc), n_files + 1).astype(int)
    paths = []
    for i in range(n_files):
        d = tmp_path / f"split_{i + 1:02d}"
        d.mkdir()
        p = d / "train_full_lightcurves.csv"
        lc.iloc[edges[i]:edges[i + 1]].to_csv(p, index=False)
        paths.append(p)
    return paths


def _reference_columns(paths):
    lc = pd.concat([pd.read_csv(p) for p in paths], ignore_index=True)       # data_loader.py:53-62
    return frame_columns(lc, None), lc


def test_csv_ingest_is_bit_identical_to_the_pandas_path(ingest, tmp_path):
    paths = _write_splits(tmp_path)
    (obj, band, t, f, e, ids), lc = _reference_columns(paths)
    for threads in (1, 4):
        got = ingest.read_lightcurve_csv(paths, threads=threads)
        assert list(got.object_ids) == list(ids)
        assert np.array_equal(got.obj, obj) and np.array_equal(got.band, band)
        for a, b in ((got.t, t), (got.flux, f), (got.err, e)):
            assert a.dtype == np.float64 and np.array_equal(a.view(np.uint64), b.view(np.uint64))
    via_loader = ingest.load_lightcurves("train", tmp_path)
    assert np.array_equal(via_loader.obj, obj) and via_loader.n_obj == len(ids)
    back = via_loader.to_frame()
    assert np.array_equal(back["Flux"].to_numpy(), lc["Flux"].to_numpy(), equal_nan=True)
    assert list(back["object_id"]) == list(lc["object_id"])


def test_number_conversion_follows_pandas_not_strtod(ingest, tmp_path):
    """random decimal spellings, including 17-digit reprs where pandas' converter is off by an ulp from strtod"""
    rng = np.random.default_rng(11)
    vals = np.concatenate([rng.normal(0, 5, 20000), 60000 + rng.uniform(0, 3000, 20000), rng.uniform(0, 1, 5000) * 10.0 ** rng.integers(-30, 30, 5000)])
    texts = [repr(float(v)) for v in vals]
    texts += ["%.3f" % v for v in vals[:5000]] + ["%.10e" % v for v in vals[:5000]] + ["%d" % int(v) for v in vals[20000:21000]]
    texts += ["0.30000000000000004", "3.141592653589793238462643383279", "1e23", "-.5", "5.", "+7.25", "1E5", "4.9e-324", "1e-400",
              "1e400", "123456789012345678901234567890.5", "0.000000000000000000000000123456789012345678", "inf", "-inf", "nan", "NA", ""]
    p = tmp_path / "numbers.csv"
    with open(p, "w") as fh:
        fh.write("object_id,Time (MJD),Flux,Flux_err,Filter\n")
        for s in texts:
            fh.write(f"a,1.5,{s},1.0,g\n")
    want = pd.read_csv(p)["Flux"].to_numpy(dtype=np.float64)
    got = ingest.read_lightcurve_csv(p).flux
    assert len(got) == len(texts)
    same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), [(texts[i], got[i], want[i]) for i in np.flatnonzero(~same)[:5]]
    off = sum(float(s) != w for s, w in zip(texts[:45000], want[:45000]))
    assert off > 0        # the reference's numbers are not the correctly rounded ones: this is why strtod would not do


def test_structural_errors_raise(ingest, tmp_path):
    p = tmp_path / "bad.csv"
    p.write_text("object_id,Time (MJD),Flux,Filter\nx,1,2,g\n")
    with pytest.raises(ingest.MallornIngestError, match="header"):
        ingest.read_lightcurve_csv(p)
    p.write_text("object_id,Time (MJD),Flux,Flux_err,Filter\nx,1,2,3\n")
    with pytest.raises(ingest.MallornIngestError, match="fields"):
        ingest.read_lightcurve_csv(p)
    p.write_text("object_id,Time (MJD),Flux,Flux_err,Filter\nx,1,abc,3,g\n")
    with pytest.raises(ingest.MallornIngestError, match="number"):
        ingest.read_lightcurve_csv(p)
    with pytest.raises(ingest.MallornIngestError, match="cannot open"):
        ingest.read_lightcurve_csv(tmp_path / "missing.csv")
    with pytest.raises(FileNotFoundError):
        ingest.load_lightcurves("test", tmp_path)


def test_header_order_crlf_and_missing_ids(ingest, tmp_path):
    p = tmp_path / "crlf.csv"
    p.write_bytes(b"Filter,Flux_err,Flux,Time (MJD),object_id\r\ng,0.5,1.25,60000.5,B\r\n\r\nr,0.25,-2.5,60001.5,\r\nz,1,3,60002,A\r\nq,1,3,60003,B")
    got = ingest.read_lightcurve_csv(p)
    want = frame_columns(pd.read_csv(p), None)
    assert list(got.object_ids) == list(want[5]) == ["B", "A"]
    assert np.array_equal(got.obj, want[0]) and got.obj.tolist() == [0, -1, 1, 0]
    assert np.array_equal(got.band, want[1]) and got.band.tolist() == [1, 2, 4, 6]
    assert np.array_equal(got.t, want[2]) and np.array_equal(got.flux, want[3]) and np.array_equal(got.err, want[4])


def test_ingest_header_symbols_are_exported(ingest):
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "mallorn_ingest.h")).read()
    names = sorted(set(re.findall(r"\b(mlrn_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 12
    L = ingest.lib()
    assert not [n for n in names if not hasattr(L, n)]


def test_number_grammar_adversarial_spellings(ingest, tmp_path):
    """digit strings built to hit every branch of the converter: more than 17 digits before / after the point, leading
    zeros, huge and tiny exponents, signs, bare points, exponent markers without digits"""
    rng = np.random.default_rng(23)
    digits = lambda n: "".join(rng.choice(list("0123456789"), n)) if n else ""
    texts = []
    for _ in range(3000):
        a, b = int(rng.integers(0, 25)), int(rng.integers(0, 25))
        if a == 0 and b == 0:
            a = 1
        s = rng.choice(["", "-", "+"]) + ("0" * int(rng.integers(0, 4)) if rng.random() < 0.3 else "") + digits(a)
        if b or rng.random() < 0.2:
            s += "." + digits(b)
        r = rng.random()
        if r < 0.35:
            s += rng.choice(["e", "E"]) + rng.choice(["", "-", "+"]) + str(int(rng.integers(0, 420)))
        elif r < 0.38:
            s += "e"                      # not a number for pandas either: the column would become strings
        texts.append(s)
    numeric = []
    for s in texts:
        col = pd.read_csv(io.StringIO(f"a,b\n{s},1\n"), dtype={"a": str})["a"].iloc[0]      # what pandas sees as text
        try:
            val = pd.read_csv(io.StringIO(f"a,b\n{s},1\n1.5,1\n"))["a"]
            if val.dtype == np.float64:
                numeric.append((s, float(val.iloc[0])))
        except Exception:
            pass
    assert len(numeric) > 2500
    p = tmp_path / "adv.csv"
    with open(p, "w") as fh:
        fh.write("object_id,Time (MJD),Flux,Flux_err,Filter\n")
        for s, _ in numeric:
            fh.write(f"a,1.5,{s},1.0,g\n")
    got = ingest.read_lightcurve_csv(p).flux
    want = np.array([v for _, v in numeric])
    same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), [(numeric[i][0], got[i], want[i]) for i in np.flatnonzero(~same)[:5]]


def test_malformed_files_are_rejected_or_parsed_never_crash(ingest, tmp_path):
    """random bytes, truncated lines, CRLF, blank lines, absurdly long numbers, NULs, wrong headers: the native reader either
    returns consistent columns or raises MallornIngestError"""
    rng = np.random.default_rng(5)
    hdr = b"object_id,Time (MJD),Flux,Flux_err,Filter\n"
    good = [b"a,60000.5,1.25,0.5,g", b"bb,6e4,-3,1,r", b",1,2,3,z", b"c,1,,3,", b"d, 1 ,2,3,u"]
    alphabet = list(b"0123456789.,-+eE \r\n\"abcNnaiIfF\t\x00\xff")
    parsed = rejected = 0
    for it in range(600):
        mode = rng.integers(0, 8)
        if mode == 0:
            body = bytes(rng.choice(alphabet, rng.integers(0, 400)).tolist())
        elif mode == 1:
            body = b"\n".join(good[i] for i in rng.integers(0, len(good), rng.integers(0, 30)))
        elif mode == 2:
            body = b"\n".join(good[i] for i in rng.integers(0, len(good), 10))[:rng.integers(0, 150)]
        elif mode == 3:
            body = b"\r\n".join(good[i] for i in rng.integers(0, len(good), 10)) + b"\r\n\r\n"
        elif mode == 4:
            body = b"x," + b"9" * int(rng.integers(1, 5000)) + b",1e" + b"9" * int(rng.integers(1, 40)) + b",0." + b"0" * 400 + b"1,g\n"
        elif mode == 5:
            body = b"\n" * int(rng.integers(0, 50))
        elif mode == 6:
            body = b",".join([b"q"] * int(rng.integers(0, 12))) + b"\n" + good[0]
        else:
            body = bytes(rng.integers(0, 256, rng.integers(0, 300), dtype=np.uint8).tolist())
        head = hdr if rng.random() < 0.85 else bytes(rng.choice(alphabet, rng.integers(0, 60)).tolist()) + b"\n"
        p = tmp_path / "f.csv"
        p.write_bytes(head + body)
        try:
            c = ingest.read_lightcurve_csv(p, threads=int(rng.integers(1, 5)))
        except ingest.MallornIngestError:
            rejected += 1
            continue
        parsed += 1
        assert len(c.obj) == len(c.band) == len(c.t) == len(c.flux) == len(c.err)
        assert (c.obj < len(c.object_ids)).all() and (c.obj >= -1).all() and (c.band <= 6).all()
    assert parsed > 100 and rejected > 100
