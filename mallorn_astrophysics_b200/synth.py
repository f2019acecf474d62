"""Seeded synthetic LSST-like light curves (the Kaggle MALLORN data is not available offline).

Follows SURVEY.md section 8(d): 6 bands, 5-9 seasonal clusters over a 1100-3000 d
survey, per-band Poisson counts (total ~150, clipped to [30, 400]), three source
classes (SN-like Bazin, TDE-like rise + t^-5/3 decay with constant blue colour,
AGN-like damped random walk), heteroscedastic noise with frequent negative fluxes.
Rows are emitted in *shuffled file order* (the real files are concatenated splits,
reference ``src/utils/data_loader.py:36-62``), because several reference features
depend on file order.

The table is produced as flat numpy columns (``LightCurveTable``) and, on demand,
as the reference's long-format ``pandas.DataFrame``
(``object_id, Time (MJD), Flux, Flux_err, Filter``; ``statistical.py:144``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .schema import LSST_BANDS

BASE_SEED = 20260118
_BAND_MEAN_COUNTS = np.array([10.0, 22.0, 38.0, 36.0, 28.0, 16.0])
_BAND_WAVE_A = np.array([3670.0, 4825.0, 6222.0, 7545.0, 8691.0, 9710.0])
_BAND_NOISE = np.array([1.8, 1.0, 1.0, 1.1, 1.5, 2.2])
CLASS_SN, CLASS_TDE, CLASS_AGN = 0, 1, 2


@dataclass
class LightCurveTable:
    """Flat long-format table; row i belongs to object ``obj[i]`` (dense index)."""
    obj: np.ndarray        # int32 [n_rows]
    band: np.ndarray       # uint8 [n_rows], 0..5 = ugrizy
    t: np.ndarray          # float64 [n_rows], MJD
    flux: np.ndarray       # float64 [n_rows]
    err: np.ndarray        # float64 [n_rows]
    n_obj: int
    z: np.ndarray          # float64 [n_obj]
    ebv: np.ndarray        # float64 [n_obj]
    label: np.ndarray      # int8 [n_obj] class id
    id_prefix: str = "obj"
    _ids: Optional[np.ndarray] = field(default=None, repr=False)

    @property
    def n_rows(self) -> int:
        return int(self.obj.shape[0])

    @property
    def object_ids(self) -> np.ndarray:
        if self._ids is None:
            self._ids = np.array([f"{self.id_prefix}_{i:08d}" for i in range(self.n_obj)], dtype=object)
        return self._ids

    def to_frame(self):
        import pandas as pd
        bands = np.array(LSST_BANDS + ["X"], dtype=object)   # code 6 = a Filter outside ugrizy
        return pd.DataFrame({
            "object_id": self.object_ids[self.obj],
            "Time (MJD)": self.t,
            "Flux": self.flux,
            "Flux_err": self.err,
            "Filter": bands[self.band],
        })

    def metadata_frame(self):
        import pandas as pd
        return pd.DataFrame({
            "object_id": self.object_ids,
            "Z": self.z,
            "EBV": self.ebv,
            "target": (self.label == CLASS_TDE).astype(np.int64),
        })

    def subset(self, n_first: int) -> "LightCurveTable":
        """First ``n_first`` objects (rows keep their relative file order)."""
        keep = self.obj < n_first
        return LightCurveTable(self.obj[keep], self.band[keep], self.t[keep], self.flux[keep],
                               self.err[keep], n_first, self.z[:n_first], self.ebv[:n_first],
                               self.label[:n_first], self.id_prefix)


def _logu(rng, lo, hi, size):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size))


def _chunk(rng: np.random.Generator, n: int, obj0: int, count_scale: float = 1.0):
    """Generate ``n`` objects; returns per-row arrays (grouped by object, then shuffled)."""
    counts = rng.poisson(_BAND_MEAN_COUNTS * count_scale, size=(n, 6)).astype(np.int64)
    tot = counts.sum(1)
    # clip total to [30, 400]: top up r, or thin proportionally
    low = tot < 30
    counts[low, 2] += 30 - tot[low]
    high = tot > 400
    if high.any():
        counts[high] = np.floor(counts[high] * (400.0 / tot[high])[:, None]).astype(np.int64)
    tot = counts.sum(1)
    nrow = int(tot.sum())
    obj_local = np.repeat(np.arange(n), tot)
    # band of each row: repeat band codes per (object, band)
    band = np.repeat(np.tile(np.arange(6, dtype=np.uint8), n), counts.reshape(-1))

    span = rng.uniform(1100.0, 3000.0, n)
    start = 60000.0 + rng.uniform(0.0, 365.0, n)
    n_season = rng.integers(5, 10, n)
    period = span / n_season
    season = np.floor(rng.uniform(0.0, 1.0, nrow) * n_season[obj_local])
    on_frac = 150.0 / 365.0
    t = start[obj_local] + season * period[obj_local] + rng.uniform(0.0, 1.0, nrow) * on_frac * period[obj_local]
    # nightly cadence granularity as in a real survey: observations of one night share ~0.01-0.3 d offsets
    t = np.floor(t) + rng.uniform(0.0, 0.35, nrow)

    cls = rng.choice(np.array([CLASS_TDE, CLASS_SN, CLASS_AGN]), size=n, p=[0.05, 0.65, 0.30]).astype(np.int8)
    z = rng.uniform(0.01, 1.0, n)
    ebv = rng.uniform(0.0, 0.15, n)

    # transient peak time inside an observing season
    pk_season = np.floor(rng.uniform(0.0, 1.0, n) * n_season)
    t0 = start + pk_season * period + rng.uniform(0.1, 0.7, n) * on_frac * period
    amp = _logu(rng, 2.0, 30.0, n)
    tau_r = rng.uniform(2.0, 12.0, n)
    tau_f = rng.uniform(15.0, 90.0, n)
    tde_t0 = rng.uniform(20.0, 60.0, n)
    beta0 = np.where(cls == CLASS_TDE, rng.uniform(1.0, 2.0, n), rng.uniform(-0.3, 1.0, n))
    redden = np.where(cls == CLASS_SN, rng.uniform(0.3, 1.5, n), 0.0)

    o = obj_local
    dt = t - t0[o]
    rcls = cls[o]
    shape = np.empty(nrow)
    tr = np.flatnonzero(rcls != CLASS_AGN)          # transients share the sigmoid rise
    ot = o[tr]
    rise = 1.0 / (1.0 + np.exp(-np.clip(dt[tr] / tau_r[ot], -60.0, 60.0)))
    is_sn = rcls[tr] == CLASS_SN
    decay_sn = np.exp(-np.clip(dt[tr] / tau_f[ot], -50.0, 700.0))
    decay_tde = np.power(1.0 + np.maximum(dt[tr], 0.0) / tde_t0[ot], -5.0 / 3.0)
    shape[tr] = rise * np.where(is_sn, decay_sn, decay_tde)
    # AGN: damped random walk approximated with random Fourier features of the OU spectrum
    ag = np.flatnonzero(rcls == CLASS_AGN)
    oa = o[ag]
    k_ff = 12
    tau_agn = _logu(rng, 30.0, 300.0, n)
    omega = np.abs(rng.standard_cauchy((n, k_ff))) / tau_agn[:, None]
    phase = rng.uniform(0.0, 2.0 * np.pi, (n, k_ff))
    lam_shift = 0.3 * (_BAND_WAVE_A[band[ag]] - 6222.0) / rng.uniform(2000.0, 8000.0, n)[oa]
    ta = t[ag] - start[oa]
    acc = np.zeros(ag.size)
    for k in range(k_ff):
        acc += np.cos(omega[oa, k] * ta + phase[oa, k] + lam_shift)
    shape[ag] = 0.6 + 0.35 * np.sqrt(2.0 / k_ff) * acc  # mostly positive, stochastic

    beta = beta0[o] + redden[o] * np.clip(np.maximum(dt, 0.0) / 60.0, 0.0, 2.5)
    colour = np.power(_BAND_WAVE_A[band] / 6222.0, -beta)
    signal = amp[o] * shape * colour

    sigma_obj = _logu(rng, 0.15, 2.5, n)
    sigma = sigma_obj[o] * _BAND_NOISE[band]
    err = sigma * rng.uniform(0.8, 1.2, nrow)
    flux = signal + rng.normal(0.0, 1.0, nrow) * sigma + rng.normal(0.0, 0.2, n)[o]

    perm = rng.permutation(nrow)
    return ((o[perm] + obj0).astype(np.int32), band[perm], t[perm], flux[perm], err[perm], z, ebv, cls)


def make_table(n_obj: int, config_index: int = 0, chunk: int = 32768, seed: Optional[int] = None,
               id_prefix: str = "obj", count_scale: float = 1.0) -> LightCurveTable:
    """Deterministic synthetic set; ``seed = 20260118 + config_index`` unless given.

    Rows are shuffled within chunks of ``chunk`` objects (file order inside an object is
    unrelated to time or band, as in the concatenated real files).  ``count_scale`` multiplies
    the per-band Poisson means (denser curves exercise the interpolation features harder).
    """
    seed = BASE_SEED + config_index if seed is None else seed
    parts = []
    obj0 = 0
    ci = 0
    while obj0 < n_obj:
        n = min(chunk, n_obj - obj0)
        rng = np.random.default_rng([seed, ci])
        parts.append(_chunk(rng, n, obj0, count_scale))
        obj0 += n
        ci += 1
    cat = lambda i: np.concatenate([p[i] for p in parts]) if len(parts) > 1 else parts[0][i]
    return LightCurveTable(cat(0), cat(1), cat(2), cat(3), cat(4), n_obj, cat(5), cat(6), cat(7), id_prefix)


def make_edge_table(seed: int = 7) -> LightCurveTable:
    """Small hand-thinned set covering the reference's small-n and degenerate branches:
    bands with 0/1/2/3/4 points, single-row objects, missing r / g / i, non-positive
    errors, all-negative fluxes, constant fluxes, duplicated times, <10 valid GP rows."""
    base = make_table(24, seed=BASE_SEED + 1000 + seed, count_scale=1.0)
    rng = np.random.default_rng(seed)
    keep = np.ones(base.n_rows, dtype=bool)
    flux, err, t, band = base.flux.copy(), base.err.copy(), base.t.copy(), base.band.copy()
    rows_of = [np.flatnonzero(base.obj == k) for k in range(base.n_obj)]

    def thin(k, b, n_keep):
        idx = rows_of[k][band[rows_of[k]] == b]
        keep[idx[n_keep:]] = False

    for k in range(0, 6):                      # band r with k points (0..5) in objects 0..5
        thin(k, 2, k)
    for k, b in ((6, 1), (7, 3), (8, 0)):      # no g / no i / no u
        thin(k, b, 0)
    thin(9, 1, 0); thin(9, 2, 0); thin(9, 3, 0)            # no g, r, i at all
    keep[rows_of[10][1:]] = False                             # single-row object
    keep[rows_of[11][7:]] = False                             # 7 rows: below the GP's 10-row floor
    err[rows_of[12][::3]] = 0.0                               # non-positive errors
    err[rows_of[12][1::7]] = -1.0
    flux[rows_of[13]] = -np.abs(flux[rows_of[13]]) - 0.1      # all negative
    flux[rows_of[14]] = 3.25                                  # constant flux
    r15 = rows_of[15]
    t[r15[1::2]] = t[r15[:-1:2]][: len(r15[1::2])]            # duplicated times (ties)
    for b in range(6):                                        # every band exactly 4 / 5 points
        thin(16, b, 4); thin(17, b, 5)
    flux[rows_of[18][band[rows_of[18]] == 2]] *= 0.0          # r band all exactly zero
    band[rows_of[19][::5]] = 6                                # unknown filters
    keep[rows_of[20]] = False                                 # object id with no rows at all
    sel = np.flatnonzero(keep)
    return LightCurveTable(base.obj[sel], band[sel], t[sel], flux[sel], err[sel], base.n_obj,
                           base.z, base.ebv, base.label, "edge")


def _device_chunk(n: int, gen_seed: int, dev):
    """One chunk of ``n`` objects of the generative model of ``_chunk``, evaluated with torch on ``dev``.  Returns
    ``(obj int64 local 0..n-1, band uint8, t, flux, err [rows]; z [n])`` in shuffled file order."""
    import torch
    f64 = torch.float64
    mean_counts = torch.tensor(_BAND_MEAN_COUNTS, dtype=f64, device=dev)
    wave = torch.tensor(_BAND_WAVE_A, dtype=f64, device=dev)
    noise = torch.tensor(_BAND_NOISE, dtype=f64, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(int(gen_seed))
    U = lambda lo, hi, size: lo + (hi - lo) * torch.rand(size, dtype=f64, device=dev, generator=g)
    logu = lambda lo, hi, size: torch.exp(U(np.log(lo), np.log(hi), size))
    counts = torch.poisson(mean_counts.expand(n, 6).contiguous(), generator=g).to(torch.int64)
    tot = counts.sum(1)
    counts[:, 2] += torch.clamp(30 - tot, min=0)
    tot = counts.sum(1)
    high = tot > 400
    if bool(high.any()):
        counts[high] = torch.floor(counts[high].to(f64) * (400.0 / tot[high].to(f64))[:, None]).to(torch.int64)
        tot = counts.sum(1)
    nrow = int(tot.sum())
    o = torch.repeat_interleave(torch.arange(n, device=dev), tot)
    band = torch.repeat_interleave(torch.arange(6, device=dev).repeat(n), counts.reshape(-1))
    span, start = U(1100.0, 3000.0, n), 60000.0 + U(0.0, 365.0, n)
    n_season = torch.randint(5, 10, (n,), device=dev, generator=g).to(f64)
    period = span / n_season
    season = torch.floor(U(0.0, 1.0, nrow) * n_season[o])
    on_frac = 150.0 / 365.0
    t = start[o] + season * period[o] + U(0.0, 1.0, nrow) * on_frac * period[o]
    t = torch.floor(t) + U(0.0, 0.35, nrow)
    u = U(0.0, 1.0, n)
    cls = torch.where(u < 0.05, CLASS_TDE, torch.where(u < 0.70, CLASS_SN, CLASS_AGN))
    z = U(0.01, 1.0, n)
    pk_season = torch.floor(U(0.0, 1.0, n) * n_season)
    t0 = start + pk_season * period + U(0.1, 0.7, n) * on_frac * period
    amp, tau_r, tau_f, tde_t0 = logu(2.0, 30.0, n), U(2.0, 12.0, n), U(15.0, 90.0, n), U(20.0, 60.0, n)
    beta0 = torch.where(cls == CLASS_TDE, U(1.0, 2.0, n), U(-0.3, 1.0, n))
    redden = torch.where(cls == CLASS_SN, U(0.3, 1.5, n), torch.zeros(n, dtype=f64, device=dev))
    dt = t - t0[o]
    rcls = cls[o]
    rise = torch.sigmoid(torch.clamp(dt / tau_r[o], -60.0, 60.0))
    decay_sn = torch.exp(-torch.clamp(dt / tau_f[o], -50.0, 700.0))
    decay_tde = torch.pow(1.0 + torch.clamp(dt, min=0.0) / tde_t0[o], -5.0 / 3.0)
    shape = rise * torch.where(rcls == CLASS_SN, decay_sn, decay_tde)
    k_ff = 12
    tau_agn = logu(30.0, 300.0, n)
    cauchy = torch.tan(np.pi * (U(0.0, 1.0, (n, k_ff)) - 0.5)).abs()
    omega = cauchy / tau_agn[:, None]
    phase = U(0.0, 2.0 * np.pi, (n, k_ff))
    lam_shift = 0.3 * (wave[band] - 6222.0) / U(2000.0, 8000.0, n)[o]
    ta = t - start[o]
    acc = torch.zeros(nrow, dtype=f64, device=dev)
    for k in range(k_ff):
        acc += torch.cos(omega[o, k] * ta + phase[o, k] + lam_shift)
    shape = torch.where(rcls == CLASS_AGN, 0.6 + 0.35 * np.sqrt(2.0 / k_ff) * acc, shape)
    beta = beta0[o] + redden[o] * torch.clamp(torch.clamp(dt, min=0.0) / 60.0, 0.0, 2.5)
    signal = amp[o] * shape * torch.pow(wave[band] / 6222.0, -beta)
    sigma = logu(0.15, 2.5, n)[o] * noise[band]
    err = sigma * U(0.8, 1.2, nrow)
    flux = signal + torch.randn(nrow, dtype=f64, device=dev, generator=g) * sigma \
        + (0.2 * torch.randn(n, dtype=f64, device=dev, generator=g))[o]
    perm = torch.randperm(nrow, device=dev, generator=g)
    return o[perm], band[perm].to(torch.uint8), t[perm], flux[perm], err[perm], z


def make_shard_columns_device(n_total: int, lo: int, hi: int, seed: int, device, chunk: int = 131072):
    """Objects ``[lo, hi)`` of the ``n_total``-object set of ``seed``, generated on ``device`` with ``obj`` re-based to
    ``0..hi-lo``.  The set is defined chunk by chunk (``chunk`` objects, one generator seed each), so every rank of a
    sharded run reproduces exactly its objects of the SAME set whatever the number of ranks: chunks that straddle a shard
    boundary are generated on both sides and cut.  Returns ``dict(obj int32, band uint8, t, flux, err float64 [n_rows];
    z float64 [hi-lo])`` in shuffled file order (shuffled inside chunks)."""
    import torch
    dev = torch.device(device)
    parts = []
    for ci in range(lo // chunk, (max(hi, lo + 1) - 1) // chunk + 1 if hi > lo else lo // chunk):
        c0 = ci * chunk
        n = min(chunk, n_total - c0)
        o, band, t, flux, err, z = _device_chunk(n, int(seed) * 1000003 + ci, dev)
        a, b = max(lo, c0) - c0, min(hi, c0 + n) - c0
        if a > 0 or b < n:
            keep = (o >= a) & (o < b)
            o, band, t, flux, err = o[keep], band[keep], t[keep], flux[keep], err[keep]
        parts.append(((o + (c0 - lo)).to(torch.int32), band, t, flux, err, z[a:b]))
    if not parts:
        e = torch.empty(0, dtype=torch.float64, device=dev)
        return {"obj": torch.empty(0, dtype=torch.int32, device=dev), "band": torch.empty(0, dtype=torch.uint8, device=dev),
                "t": e, "flux": e.clone(), "err": e.clone(), "z": e.clone()}
    cat = lambda i: torch.cat([p[i] for p in parts]) if len(parts) > 1 else parts[0][i]
    return {"obj": cat(0), "band": cat(1), "t": cat(2), "flux": cat(3), "err": cat(4), "z": cat(5)}


def make_columns_device(n_obj: int, seed: int, device, obj0: int = 0, chunk: int = 131072):
    """The same generative model as ``_chunk`` evaluated with torch on ``device`` (bench-scale sets:
    10^6-10^7 objects never exist on the host).  Not bit-identical to ``make_table`` -- parity tests
    use the numpy generator; this one only has to draw from the same distribution.

    Returns ``dict(obj int32, band uint8, t, flux, err float64 [n_rows]; z float64 [n_obj])`` in
    shuffled file order (shuffled inside chunks of ``chunk`` objects); ``obj`` starts at ``obj0``.
    """
    cols = make_shard_columns_device(n_obj, 0, n_obj, seed, device, chunk)
    if obj0:
        cols["obj"] = cols["obj"] + int(obj0)
    return cols
