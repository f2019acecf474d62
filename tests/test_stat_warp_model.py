"""CPU model of the algorithmic pieces of the K2 fast path (``csrc/stat_warp.cuh``), checked against numpy: the
comparator schedule of the register sorting network, the 24-bit keys with the odd-even repair on the exact values, the
stable partition into band runs, the binary-search MAD and the all-band moments assembled from the band moments.  The
kernel itself is held to the reference on the GPU (``tests/test_gpu_parity.py``); this file keeps its *design* honest where
there is no GPU."""
import numpy as np
import pytest


def network_schedule(n=256):
    """(lower, upper) position pairs of every stage of sort256(): each merge of size k first pairs E with E ^ (k - 1),
    then E with E ^ j for j = k/4 ... 1; every comparator puts the smaller key at the lower position."""
    stages, k = [], 2
    while k <= n:
        masks = [k - 1] + [j for j in (k >> s for s in range(2, k.bit_length())) if j >= 1]
        for m in masks:
            stages.append([(e, e ^ m) for e in range(n) if (e ^ m) > e])
        k <<= 1
    return stages


def run_network(keys):
    v = np.array(keys, dtype=np.uint64)
    for stage in network_schedule(len(v)):
        lo = np.array([a for a, _ in stage]); hi = np.array([b for _, b in stage])
        a, b = np.minimum(v[lo], v[hi]), np.maximum(v[lo], v[hi])
        v[lo], v[hi] = a, b
    return v


def test_network_has_the_stage_counts_the_kernel_unrolls():
    st = network_schedule(256)
    assert len(st) == 36                                        # 21 in-lane stages + 15 cross-lane stages
    cross = [s for s in st if any((a >> 3) != (b >> 3) for a, b in s)]
    assert len(cross) == 15
    assert all(len(s) == 128 for s in st)


@pytest.mark.parametrize("seed", range(5))
def test_network_sorts(seed):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 2 ** 32, 256, dtype=np.uint64)
    keys[rng.integers(0, 256, 40)] = keys[0]                    # duplicates
    assert np.array_equal(run_network(keys), np.sort(keys))


def keys_and_repair(f):
    """The kernel's sort of one object's fluxes: 24-bit keys over a range good to float precision, index in the low byte,
    padding at the top, then odd-even transposition on the exact values until nothing is out of order."""
    n = len(f)
    lo_f, hi_f = np.float32(f.min()), np.float32(f.max())
    lo = float(lo_f) - 2.4e-7 * abs(float(lo_f)) - 1e-300
    hi = float(hi_f) + 2.4e-7 * abs(float(hi_f)) + 1e-300
    scale = 16777215.0 / (hi - lo) * (1.0 - 1e-9)
    q = ((f - lo) * scale).astype(np.uint64)
    assert (q < 2 ** 24).all() and (f >= lo).all() and (f <= hi).all()
    key = np.full(256, 0xFFFFFFFF, dtype=np.uint64)
    key[:n] = (q << np.uint64(8)) | np.arange(n, dtype=np.uint64)
    key = run_network(key)
    idx = (key[:n] & np.uint64(255)).astype(int)
    g = f[idx].copy()
    passes = 0
    while (g[:-1] > g[1:]).any():
        for start in (0, 1):
            a = np.arange(start, n - 1, 2)
            sw = g[a] > g[a + 1]
            g[a[sw]], g[a[sw] + 1] = g[a[sw] + 1].copy(), g[a[sw]].copy()
            idx[a[sw]], idx[a[sw] + 1] = idx[a[sw] + 1].copy(), idx[a[sw]].copy()
        passes += 1
        assert passes <= n
    return g, idx, passes


def test_quantised_keys_with_repair_give_the_exact_sort():
    rng = np.random.default_rng(3)
    cases = [rng.normal(50.0, 5.0, 150), rng.normal(0.0, 1.0, 255), np.full(60, 3.25),
             np.concatenate([1000.0 - 1e-6 * np.arange(199), [1e9]]),             # every flux but one in a single key, reversed
             np.repeat(rng.normal(10.0, 2.0, 8), 8), rng.normal(1e-3, 1e-6, 150) + np.where(np.arange(150) == 77, 5e3, 0.0),
             -np.abs(rng.normal(1e4, 1e3, 31))]
    repaired = 0
    for f in cases:
        g, idx, passes = keys_and_repair(f)
        assert np.array_equal(g, np.sort(f)) and np.array_equal(np.sort(idx), np.arange(len(f)))
        repaired += passes > 0
    assert repaired >= 2                                         # the tie cases really exercise the repair


def test_stable_partition_gives_the_band_runs():
    rng = np.random.default_rng(4)
    n = 200
    band = np.sort(rng.integers(0, 7, n))                       # store order: band-major
    f = rng.normal(0.0, 1.0, n)
    off = np.searchsorted(band, np.arange(8))
    g, idx, _ = keys_and_repair(f)
    run = np.empty(n)
    cnt = np.zeros(7, dtype=int)                                # the kernel keeps these as 8-bit fields of one 64-bit word per lane + a scan
    for k in range(n):
        b = band[idx[k]]
        run[off[b] + cnt[b]] = g[k]
        cnt[b] += 1
    for b in range(7):
        assert np.array_equal(run[off[b]:off[b + 1]], np.sort(f[off[b]:off[b + 1]]))


def mad_by_window(s):
    """median |s - med| of an ascending run: the h+1 values closest to the median are a contiguous window; its start is
    the first c at which the window need not move right (binary search), as in the kernel."""
    n = len(s)
    h = n >> 1
    med = s[h] if n & 1 else 0.5 * (s[h - 1] + s[h])
    w = h + 1
    c0, c1 = 0, n - w
    while c0 < c1:
        mid = (c0 + c1) >> 1
        if med - s[mid] > s[mid + w] - med:
            c0 = mid + 1
        else:
            c1 = mid
    dl, dr = med - s[c0], s[c0 + w - 1] - med
    dh = max(dl, dr)
    if n & 1:
        return dh
    if w == 1:
        dprev = dh
    elif dl >= dr:
        dprev = max(med - s[c0 + 1], dr)
    else:
        dprev = max(dl, s[c0 + w - 2] - med)
    return 0.5 * (dprev + dh)


def test_window_mad_equals_numpy():
    rng = np.random.default_rng(5)
    for n in list(range(1, 40)) + [64, 150, 255]:
        for _ in range(20):
            s = np.sort(rng.standard_cauchy(n) if n % 3 else np.round(rng.normal(0, 2, n)))      # heavy tails; many exact ties
            want = np.median(np.abs(s - np.median(s)))
            assert mad_by_window(s) == pytest.approx(want, rel=1e-14, abs=1e-300), n


def test_all_band_moments_from_band_moments():
    rng = np.random.default_rng(6)
    sizes = [10, 22, 38, 36, 28, 16, 3]
    parts = [rng.normal(loc, sc, k) for loc, sc, k in zip(rng.normal(100.0, 50.0, 7), rng.uniform(0.5, 20.0, 7), sizes)]
    allf = np.concatenate(parts)
    n, mean = len(allf), allf.mean()
    M2 = M3 = M4 = 0.0
    for p in parts:                                             # binomial shift of each band's central moments to the common mean
        nb, mb = len(p), p.mean()
        d = p - mb
        m2, m3, m4 = (d ** 2).sum(), (d ** 3).sum(), (d ** 4).sum()
        dl = mb - mean
        M2 += m2 + nb * dl ** 2
        M3 += m3 + 3 * dl * m2 + nb * dl ** 3
        M4 += m4 + 4 * dl * m3 + 6 * dl ** 2 * m2 + nb * dl ** 4
    d = allf - mean
    sd = np.sqrt((d ** 2).mean())
    assert np.sqrt(M2 / n) == pytest.approx(sd, rel=1e-12)
    assert (M3 / n) / sd ** 3 == pytest.approx(((d / sd) ** 3).mean(), rel=1e-9, abs=1e-12)
    assert (M4 / n) / sd ** 4 - 3.0 == pytest.approx(((d / sd) ** 4).mean() - 3.0, rel=1e-9, abs=1e-12)
