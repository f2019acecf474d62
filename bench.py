#!/usr/bin/env python
"""bench.py -- objects/s of the MALLORN light-curve featurisation on B200 (contract in the task statement).

A *step* is one pass of the hot path over one batch of synthetic light curves: K1 pack (flat long-format rows -> CSR
(object, band) segments in HBM) followed by every feature pass of the 434-column module table the v34a/v92d 222-feature set
is selected from (statistical, colors, shape, tde_physics, physics_based, bazin, power-law, 2-D GP), through the product's
own entry points (``shard.featurise_shard`` / ``shard.featurise_sharded``).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c4|c5|c1|c2|c3]

* N = 1 (default ``--config c4``): BASELINE.json configs[3], 10^6 objects on one B200.
* N > 1 (default ``--config c5``, launched by torchrun): BASELINE.json configs[4], ONE fixed 10^7-object set.  Each rank
  owns a contiguous range of the set (10^7 / N objects) and works through it in batches of 1.25 x 10^6 objects: a step is
  one batch on every rank (N x 1.25 x 10^6 objects of the set, no collective during compute, ONE NCCL all-gather of the
  step's float64 rows into their place in the full 10^7-row table every rank holds), so 8 / N steps make one pass over
  the whole set and a step costs the same at every N -- the driver's ``--steps 20 --warmup 5`` stays inside its time
  limit at N = 2 (a step over the whole set would be 53 s there).  ``config.set_passes_timed`` says how many passes over
  the fixed set the timed steps made, ``config.seconds_per_set_pass`` is the strong-scaling reading of the same number.
* ``--config c1|c2|c3``: the 10,178-object configs through the DataFrame-in / DataFrame-out API with the host / device
  split (c1 statistics + Bazin, c2 the 2-D GP, c3 power law + the full table).
* ``value``  objects/s, inputs resident in HBM when the clock starts; ``e2e`` the same through the C-ABI with HOST buffers
  (H2D of the input columns and D2H of this rank's feature table inside the timed region).
* ``roofline`` the moment pass against measured HBM bandwidth; ``roofline_dominant`` the GP pass against the FP64 rate
  measured on the device in this run (``mlrn_measure_fp64``); ``passes`` every pass's share of the step.
* ``cpu_baseline`` the CPU oracle (``oracle/``: the reference's numpy/scipy arithmetic) on a bounded sample, all host cores.
* ``--impl reference`` times only that CPU path (the reference is a Python package without native code and
  ``/root/reference`` does not exist on the GPU box, so "the reference's own CPU implementation" is the oracle port); its
  ``config.workload`` names the sample it actually timed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODULE_WIDTH = {"statistical": 123, "colors": 83, "shape": 65, "tde": 25, "physics": 32, "bazin": 52, "powerlaw": 27,
                "gp2d": 27, "gp1d": 21}
MODULES = ("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d")
METRIC = "light curves/sec for the 222-feature set (full 434-column module table)"
UNIT = "objects/s"
BASE_SEED = 20260118           # seed = 20260118 + config index
C5_OBJECTS = 10_000_000
C5_BATCH_PER_GPU = 1_250_000     # one eighth of the C5 set: what one GPU takes per step (the whole shard at N = 8)
# ncu --set full captures of THIS build's kernels (profiles/r02_*.txt), per object; see profiles/README.md
NCU = {"stat_dram_bytes_per_object": None, "stat_source": None, "fits": None}
try:
    NCU.update(json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_constants.json"))))
except Exception:
    pass


# ----------------------------------------------------------------------------------------------
# clocks under load (B200_PROFILING.md "clocks line"), sampled through NVML while the timed region runs
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.wait(0.05):
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle over object chunks on all host cores
# ----------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    import warnings
    from oracle.extract import run_module
    cols, lo, hi, modules = args
    keep = (cols["obj"] >= lo) & (cols["obj"] < hi)
    obj = (cols["obj"][keep] - lo).astype(np.int32)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in modules:
            out[m] = run_module(m, obj, cols["band"][keep], cols["t"][keep], cols["flux"][keep], cols["err"][keep],
                                hi - lo, cols["z"][lo:hi])
    return lo, hi, out


def cpu_run(cols, n_objects, modules, cores, keep_rows=False):
    """The CPU oracle on the first ``n_objects`` objects, ``cores`` worker processes.  Returns (objects/s, seconds, rows)
    with rows = {module: [n_objects, cols]} when ``keep_rows``."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    per = max(1, (n_objects + cores * 2 - 1) // (cores * 2))
    jobs = [(cols, lo, min(lo + per, n_objects), modules) for lo in range(0, n_objects, per)]
    t0 = time.perf_counter()
    if cores == 1:
        parts = [_cpu_chunk(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            parts = pool.map(_cpu_chunk, jobs)
    dt = time.perf_counter() - t0
    rows = None
    if keep_rows:
        rows = {m: np.full((n_objects, MODULE_WIDTH[m]), np.nan) for m in modules}
        for lo, hi, out in parts:
            for m, (present, r) in out.items():
                rows[m][lo + present] = r
    return n_objects / dt, dt, rows


def cpu_rate(cols, n_objects, modules, cores):
    rate, dt, _ = cpu_run(cols, n_objects, modules, cores)
    return rate, dt


def host_sample(n_objects, config_index=3):
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(n_objects, seed=BASE_SEED + config_index)
    return {"obj": tb.obj, "band": tb.band, "t": tb.t, "flux": tb.flux, "err": tb.err, "z": tb.z}


def ingest_rates(n_objects=10178, device=0, n_files=20):
    """SURVEY 8(f) row 1 measured on this box: the reference's loading path (pd.read_csv of the split files + concat, then
    the factorisation) against the native ingest (libmallorn_ingest.so) on the same files, and files -> feature table
    through the chunked, overlapped pipeline."""
    import shutil
    import tempfile
    import pandas as pd
    from mallorn_astrophysics_b200 import ingest
    from mallorn_astrophysics_b200.store import frame_columns
    from mallorn_astrophysics_b200.synth import make_table
    root = tempfile.mkdtemp(prefix="mlrn_bench_ingest_")
    try:
        lc = make_table(n_objects, seed=BASE_SEED + 3).to_frame()
        edges = np.linspace(0, len(lc), n_files + 1).astype(int)
        paths = []
        for i in range(n_files):
            d = os.path.join(root, f"split_{i + 1:02d}")
            os.makedirs(d)
            paths.append(os.path.join(d, "train_full_lightcurves.csv"))
            lc.iloc[edges[i]:edges[i + 1]].to_csv(paths[-1], index=False)
        size = sum(os.path.getsize(p) for p in paths)
        t0 = time.perf_counter()
        ref = frame_columns(pd.concat([pd.read_csv(p) for p in paths], ignore_index=True), None)
        t_ref = time.perf_counter() - t0
        ingest.read_lightcurve_csv(paths)                                   # warm (page cache, thread pool)
        t0 = time.perf_counter()
        got = ingest.read_lightcurve_csv(paths)
        t_nat = time.perf_counter() - t0
        same = bool(np.array_equal(got.obj, ref[0]) and np.array_equal(got.band, ref[1]) and
                    all(np.array_equal(a.view(np.uint64), b.view(np.uint64))
                        for a, b in ((got.t, ref[2]), (got.flux, ref[3]), (got.err, ref[4]))))
        out = {"reference": "load_lightcurves (data_loader.py:36-62) + groupby factorisation", "objects": int(n_objects), "files": n_files, "rows": int(len(lc)),
               "csv_bytes": int(size), "pandas_mb_per_s": size / 1e6 / t_ref, "native_mb_per_s": size / 1e6 / t_nat,
               "native_rows_per_s": len(lc) / t_nat, "threads": os.cpu_count(), "bit_identical_to_pandas": same}
        if hasattr(ingest, "featurise_files"):
            try:
                mods = ("statistical", "colors", "shape", "tde", "bazin", "powerlaw", "gp2d")     # every pass that needs no metadata
                ingest.featurise_files(paths[:2], mods, device=device)       # warm
                t0 = time.perf_counter()
                _, _, info = ingest.featurise_files(paths, mods, device=device)
                out["files_to_table_seconds"] = time.perf_counter() - t0
                out["files_to_table"] = info
                t0 = time.perf_counter()                                          # the same without overlap: read everything, then the GPU
                cols = ingest.read_lightcurve_csv(paths)
                t_read = time.perf_counter() - t0
                out["files_to_table_serial_estimate_seconds"] = t_read + info["pack_and_passes_seconds"]
            except Exception as exc:
                out["files_to_table_error"] = f"{type(exc).__name__}: {exc}"
        return out
    except Exception as exc:                                                # never lose the bench line over the side measurement
        return {"error": f"{type(exc).__name__}: {exc}"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def frames_e2e(n_objects=10178, device=0, modules=MODULES, config_index=1):
    """The user-facing call itself on a 10,178-object set (BASELINE.json configs[0..2] shape): the reference's long
    DataFrame in, the reference's per-module DataFrames out (features.pipeline.extract_feature_tables = one K1 pack + the
    passes), host factorisation and DataFrame assembly included, with the host / device split."""
    try:
        import torch
        from mallorn_astrophysics_b200.features.pipeline import extract_feature_tables
        from mallorn_astrophysics_b200.store import LightCurveStore, frame_columns
        from mallorn_astrophysics_b200.synth import make_table
        tb = make_table(n_objects, config_index=config_index)
        lc, meta, ids = tb.to_frame(), tb.metadata_frame(), list(tb.object_ids)
        warm = ids[:512]
        extract_feature_tables(lc[lc["object_id"].isin(set(warm))], meta, warm, modules=modules, device=device)
        t0 = time.perf_counter()
        tables = extract_feature_tables(lc, meta, ids, modules=modules, device=device)
        dt = time.perf_counter() - t0
        # the split: host factorisation alone, and the device part alone (pack + passes with host buffers, no frames)
        t0 = time.perf_counter()
        codes, band, t, flux, err, _ = frame_columns(lc, ids)
        t_fact = time.perf_counter() - t0
        t0 = time.perf_counter()
        with LightCurveStore(codes, band, t, flux, err, len(ids), device=device) as s:
            for m in modules:
                s.features(m, z=tb.z if m == "physics" else None)
        torch.cuda.synchronize(device)
        t_dev = time.perf_counter() - t0
        return {"call": "features.pipeline.extract_feature_tables(lightcurves, metadata, object_ids)", "objects": int(n_objects),
                "rows": int(len(lc)), "seconds": dt, "objects_per_s": n_objects / dt,
                "host_factorise_seconds": t_fact, "device_pack_and_passes_seconds": t_dev,
                "frame_assembly_seconds": max(dt - t_fact - t_dev, 0.0),
                "columns": int(sum(len(t.columns) - 1 for t in tables.values())), "modules": list(modules)}
    except Exception as exc:                                                # never lose the bench line over the side measurement
        return {"error": f"{type(exc).__name__}: {exc}"}


def reference_workload(per_step, modules, n_rows):
    return {"workload": f"CPU reference arm: {per_step} objects per step of the C4 generator (seed {BASE_SEED + 3}, {n_rows} "
                        f"observations) -- a bounded SAMPLE of BASELINE.json configs[3], not the 10^6-object set; objects/s "
                        f"extrapolates linearly (objects are independent)",
            "sample_objects_per_step": per_step, "same_config_as_gpu_arm": False, "bands": 6, "mean_obs_per_object": 150,
            "modules": list(modules), "columns": sum(MODULE_WIDTH[m] for m in modules),
            "parallelism": "one process per host core over object chunks (the reference itself is single-process)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = args.cpu_objects or 32 * cores
    cols = host_sample(per_step)
    modules = list(MODULES)
    for _ in range(args.warmup):
        cpu_rate(cols, min(per_step, 2 * cores), modules, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_rate(cols, per_step, modules, cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} objects/step of the C4 generator (seed {BASE_SEED + 3}), {len(cols['obj'])} observations, {cores} host cores"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": reference_workload(per_step, modules, int(len(cols["obj"]))),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": cores,
    }))


def workload_config(name, world, n_total, n_local, modules, n_rows, n_batches=1, objects_per_step=None, steps=0, step_s=0.0):
    extra = {}
    if name == "c5":
        label = (f"C5: one fixed {n_total}-object LSST-scale synthetic set sharded over {world} B200 (BASELINE.json configs[4]); "
                 f"each rank owns a contiguous range of {n_local} objects and featurises it {n_batches} batch(es) of "
                 f"~{-(-n_local // n_batches)} objects at a time: a step = one batch on every rank + one all-gather of its float64 "
                 f"rows into the full table; {n_batches} step(s) = one pass over the whole set")
        extra = {"steps_per_set_pass": n_batches, "set_passes_timed": steps / n_batches,
                 "seconds_per_set_pass": step_s * n_batches, "objects_per_step": objects_per_step}
    elif n_total == 1_000_000:
        label = "C4: 1M-object LSST-like synthetic set, full feature table (BASELINE.json configs[3])"
    else:
        label = f"C4 generator at {n_total} objects (BASELINE.json configs[3] shape)"
    cfg = {"workload": label, "objects_total": n_total, "objects_this_rank": n_local, "bands": 6, "mean_obs_per_object": 150,
           "modules": list(modules), "columns": sum(MODULE_WIDTH[m] for m in modules),
           "l2": "inputs larger than L2 (no flush needed)" if n_rows * 25 > 2 * 126e6 else "L2 flushed between steps",
           "parallelism": f"objects sharded over {world} GPU(s) (shard.featurise_shard per rank + shard.allgather_table), one NCCL "
                          f"all-gather of the step's [objects, {sum(MODULE_WIDTH[m] for m in modules)}] float64 rows" if world > 1 else
                          "one GPU (shard.featurise_shard)"}
    cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def algorithmic_bytes(n_rows, n_obj):
    """Moment pass, SURVEY.md 8(d) with this build's layout: one read of every observation
    (t, flux, err float64 + uint16 time rank = 26 B), the 8 segment offsets of the object (int64) and one
    write of the 123-column float64 row."""
    return 26 * n_rows + n_obj * (8 * 8 + 123 * 8)


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from mallorn_astrophysics_b200 import _native, shard
    from mallorn_astrophysics_b200.synth import make_shard_columns_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _native.lib()
    symbol = {"statistical": "mlrn_stat", "colors": "mlrn_color", "shape": "mlrn_shape", "tde": "mlrn_tde",
              "physics": "mlrn_physics", "bazin": "mlrn_bazin", "powerlaw": "mlrn_powerlaw", "gp2d": "mlrn_gp2d",
              "gp1d": "mlrn_gp1d"}
    modules = [m for m in (args.modules.split(",") if args.modules else MODULES) if hasattr(L, symbol[m])]
    config = args.config or ("c5" if world > 1 else "c4")
    if config == "c5":
        n_total = args.objects or C5_OBJECTS
        seed = (BASE_SEED + 4) * 131
    else:
        n_total = (args.objects or 1_000_000) * world
        seed = (BASE_SEED + 3) * 131
    ranges = shard.equal_ranges(n_total, world)
    lo, hi = ranges[rank]
    # batches of this rank's range: C5 works through the range 1.25 x 10^6 objects at a time (module docstring), C4 is one batch
    sub = shard.batch_ranges(ranges, (args.batch_objects or C5_BATCH_PER_GPU) if config == "c5" else max(n_total, 1))   # sub[rank][batch]
    n_batches = len(sub[0])
    batch_rows = [[y - x for x, y in (sub[r][j] for r in range(world))] for j in range(n_batches)]   # rows_per_rank of step j
    batches = [make_shard_columns_device(n_total, x, y, seed, dev) for x, y in sub[rank]]            # this rank's objects of the ONE set, in HBM
    n_obj_max = max(y - x for x, y in sub[rank])
    n_rows = max(int(c["obj"].shape[0]) for c in batches)
    n_rows_all = sum(int(c["obj"].shape[0]) for c in batches)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    width = shard.table_width(modules)
    table_buf = torch.empty((n_obj_max, width), dtype=torch.float64, device=dev)
    # the full table of the set on every rank, batch-major: block j holds [rank 0's batch j | rank 1's batch j | ...]
    gathered = [torch.empty((world * max(batch_rows[j]), width), dtype=torch.float64, device=dev)
                for j in range(n_batches)] if world > 1 else None
    gp_status = [torch.zeros((y - x, 4), dtype=torch.int32, device=dev) for x, y in sub[rank]] if "gp2d" in modules else None
    flush = None
    if n_rows * 25 < 2 * 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    pass_ms = {m: 0.0 for m in modules}
    pass_ms["pack"] = 0.0
    pass_ms["allgather"] = 0.0
    timed_batches = [0] * n_batches                  # how often each batch was inside the timed region

    def step(k: int, timed: bool):
        j = k % n_batches
        cols, (x, y) = batches[j], sub[rank][j]
        marks = [("start", ev())]
        marks[0][1].record(stream)

        def mark(name):
            e = ev(); e.record(stream); marks.append((name, e))

        table = table_buf[: y - x]
        shard.featurise_shard(cols, y - x, modules, table=table, device=local, stream=sptr,
                              status={"gp2d": gp_status[j]} if gp_status is not None else {}, on_pass=mark)
        if world > 1:
            shard.allgather_table(table, batch_rows[j], out=gathered[j])
            mark("allgather")
        stream.synchronize()
        if timed:
            timed_batches[j] += 1
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                pass_ms[name] = pass_ms.get(name, 0.0) + a.elapsed_time(b)      # 'bazin+powerlaw': the two fit passes run as one overlapped call (mlrn_fits)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k, False)
    launches0 = L.mlrn_launch_count()
    barrier()
    with ClockSampler(local) as clk:
        t_start, t_end = ev(), ev()
        t_start.record(stream)
        for k in range(args.steps):
            if flush is not None:
                flush.zero_()
            step(args.warmup + k, True)
        t_end.record(stream)
        barrier()
    launches = L.mlrn_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    objects_timed = sum(timed_batches[j] * sum(batch_rows[j]) for j in range(n_batches))      # all ranks
    value = objects_timed / (ms * 1e-3)
    step_s = ms * 1e-3 / args.steps
    cols, n_obj = batches[0], sub[rank][0][1] - sub[rank][0][0]      # the legs beside the step use this rank's first batch

    # ---- section 8(f) row measured beside the step: the per-band 1-D GP pass (extract_gp_features), on a bounded slice
    widened = None
    if rank == 0 and hasattr(L, "mlrn_gp1d") and "gp1d" not in modules and not args.modules and args.gp1d_objects > 0:
        from mallorn_astrophysics_b200.store import LightCurveStore
        k = min(n_obj, args.gp1d_objects)
        keep = cols["obj"] < k
        sl = {c: cols[c][keep].contiguous() for c in ("obj", "band", "t", "flux", "err")}
        o1 = torch.empty((k, _native.COLS["gp1d"]), dtype=torch.float64, device=dev)
        s1 = torch.zeros((k, 12), dtype=torch.int32, device=dev)
        st1 = LightCurveStore(sl["obj"], sl["band"], sl["t"], sl["flux"], sl["err"], k, device=local, stream=sptr)
        st1.features("gp1d", out=o1, status=s1)                       # warm-up
        a, b = ev(), ev()
        a.record(stream)
        st1.features("gp1d", out=o1, status=s1)
        b.record(stream)
        stream.synchronize()
        st1.close()
        gms = a.elapsed_time(b)
        fitted = int((s1[:, 2::3] >= 0).sum())
        evals = int(s1[:, 0::3].sum())
        widened = {"gp1d": {"reference": "extract_gp_features (gaussian_process.py:251), 21 columns", "objects": k,
                            "ms": gms, "objects_per_s": k / (gms * 1e-3), "bands_fitted": fitted,
                            "objective_evaluations": evals, "evaluations_per_s": evals / (gms * 1e-3),
                            "observations": int(sl["obj"].shape[0]), "launches": 1}}
        del sl, o1, s1

    # ---- end to end: host buffers through the C-ABI, copies inside the timed region
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_warm = min(args.warmup, 3)
    if step_s > 8.0:                                   # long steps (big shards): one warm-up and one timed step bound the run
        e2e_steps, e2e_warm = 1, 1
    host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in cols.items()}
    hnp = {k: v.numpy() for k, v in host.items()}
    torch.cuda.synchronize()
    batch_shapes = [(int(c["obj"].shape[0]), y - x) for c, (x, y) in zip(batches, sub[rank])]       # (observations, objects) per batch
    cols = None                                        # the device-resident inputs are not needed any more (big shards: 20+ GB)
    batches.clear()
    table = table_buf[:n_obj]
    rows_per_rank = batch_rows[0]
    htable = torch.empty((n_obj, width), dtype=torch.float64, pin_memory=True)
    h2d = sum(hnp[k].nbytes for k in ("obj", "band", "t", "flux", "err")) + (hnp["z"].nbytes if "physics" in modules else 0)
    d2h = htable.numel() * 8

    def e2e_step():
        shard.featurise_shard(hnp, n_obj, modules, table=table, device=local, stream=sptr)
        if world > 1:
            shard.allgather_table(table, rows_per_rank, out=gathered[0])
        htable.copy_(table, non_blocking=True)
        stream.synchronize()

    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tms = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_s = float(tms.item())
    e2e_value = sum(rows_per_rank) * e2e_steps / e2e_s
    del host, hnp, htable

    if rank == 0:
        import ctypes
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        stat_ms = pass_ms["statistical"] / args.steps if "statistical" in modules else None
        roof = None
        if stat_ms:
            # per launch = averaged over the timed steps (C5: the batches differ by a few objects)
            alg = sum(timed_batches[j] * algorithmic_bytes(*batch_shapes[j]) for j in range(n_batches)) / args.steps
            obj_per_launch = sum(timed_batches[j] * batch_shapes[j][1] for j in range(n_batches)) / args.steps
            ach = alg / (stat_ms * 1e-3) / 1e9
            per_obj = NCU.get("stat_dram_bytes_per_object")
            roof = {"kernel": "k_stat_warp (moment pass, stat_warp.cuh / stat_fast.cu)", "bound": "hbm", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak,
                    "traffic": per_obj * obj_per_launch if per_obj else None,
                    "traffic_source": NCU.get("stat_source"),
                    "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch": alg, "ms_per_launch": stat_ms}
        roof_gp = None
        if gp_status is not None and pass_ms.get("gp2d"):
            # dominant kernel of the step: FP64 work of the GP = evaluations x N^3 flop
            # (N^3/2 FMA per -lnL + gradient evaluation) over its CUDA-event time, against the FP64
            # rate measured on this device right now (the DMMA loop: the instruction the sweep issues)
            dfma, dmma = ctypes.c_double(0.0), ctypes.c_double(0.0)
            rc = L.mlrn_measure_fp64(local, ctypes.byref(dfma), ctypes.byref(dmma))
            fp64_peak = max(dfma.value, dmma.value) if rc == 0 and max(dfma.value, dmma.value) > 0 else 37.1
            flop, stt, fitted = 0.0, None, None
            for j in range(n_batches):                                  # per launch = averaged over the timed steps
                if timed_batches[j]:
                    stt = gp_status[j].to(torch.float64)
                    fitted = stt[:, 3] >= 10
                    flop += timed_batches[j] * float(((stt[:, 1] * stt[:, 3] ** 3)[fitted]).sum())      # nfev evaluations (the one at the optimum is normally not repeated)
            flop /= args.steps
            gp_ms = pass_ms["gp2d"] / args.steps
            ach_tf = flop / (gp_ms * 1e-3) / 1e12
            roof_gp = {"kernel": "k_gp2d (2-D GP)", "bound": "tensor", "achieved": ach_tf, "peak": fp64_peak,
                       "unit": "TFLOP/s", "frac": ach_tf / fp64_peak, "traffic": None,
                       "peak_source": f"measured in this run on cuda:{local} (mlrn_measure_fp64: DFMA {dfma.value:.2f}, "
                                      f"DMMA m8n8k4 {dmma.value:.2f} TFLOP/s)" if rc == 0 else "fallback 37.1 TFLOP/s (round-1 measurement)",
                       "algorithmic_flop_per_launch": flop, "ms_per_launch": gp_ms, "share_of_step": pass_ms["gp2d"] / ms,
                       "mean_evaluations": float(stt[:, 1][fitted].mean()), "mean_points": float(stt[:, 3][fitted].mean())}
        cpu = None
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_objects or 48 * cores
            sample_cols = host_sample(n_cpu)
            rate, secs = cpu_rate(sample_cols, n_cpu, modules, cores)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_cpu} objects of the C4 generator (seed {BASE_SEED + 3}), modules {','.join(modules)}, "
                             f"{secs:.1f} s on {cores} processes"}
            if widened is not None:
                n_w = min(n_cpu, 8 * cores)
                rate_w, secs_w = cpu_rate(sample_cols, n_w, ["gp1d"], cores)
                widened["gp1d"]["cpu_baseline"] = {"value": rate_w, "unit": UNIT, "cores": cores, "kind": "port",
                                                   "sample": f"first {n_w} objects of the CPU sample, oracle/gp1d.py, {secs_w:.1f} s"}
                widened["ingest"] = ingest_rates(device=local)
                widened["frames_e2e"] = frames_e2e(device=local)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak",                     # per-GPU work per step is fixed as N grows (C4: 10^6 objects; C5: 1.25 x 10^6 of the set)
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(config, world, n_total, hi - lo, modules, n_rows, n_batches=n_batches,
                                      objects_per_step=objects_timed / args.steps, steps=args.steps, step_s=step_s),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "call": "shard.featurise_shard(host columns) + all-gather + D2H of this rank's table"},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_dominant": roof_gp, "roofline_fits": NCU.get("fits"), "cpu_baseline": cpu,
            "passes": {k: {"ms_per_step": v / args.steps, "share": v / ms} for k, v in pass_ms.items() if v > 0},
            "observations_this_rank": n_rows_all, "observations_this_rank_per_step": n_rows,
            "widened": widened,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# the 10,178-object configs (BASELINE.json configs[0..2]) through the DataFrame API
# ----------------------------------------------------------------------------------------------
def run_small_config(args):
    """One JSON line for c1 / c2 / c3: the reference-facing call (DataFrames in and out) on the 10,178-object set, its
    host / device split, the CPU oracle on the same objects and -- for the fit passes -- the parity classes on the set."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import warnings
    import torch
    from mallorn_astrophysics_b200.synth import make_table
    name = args.config
    modules = {"c1": ("statistical", "colors", "shape", "physics", "bazin"), "c2": ("gp2d",),
               "c3": MODULES}[name]
    ci = {"c1": 0, "c2": 1, "c3": 2}[name]
    label = {"c1": "C1: v34a base statistics + Bazin fits on the 10,178-object set (BASELINE.json configs[0])",
             "c2": "C2: v19 2-D multi-band GP features on the 10,178-object set (BASELINE.json configs[1])",
             "c3": "C3: v55 power law + the full v92d module table on the 10,178-object set (BASELINE.json configs[2])"}[name]
    n = args.objects or 10178
    cores = os.cpu_count() or 1
    tb = make_table(n, config_index=ci)
    cols = {"obj": tb.obj, "band": tb.band, "t": tb.t, "flux": tb.flux, "err": tb.err, "z": tb.z}
    n_cpu = n if args.full_cpu else min(n, args.cpu_objects or 64 * cores)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rate, secs, cpu_rows = cpu_run(cols, n_cpu, list(modules), cores, keep_rows=True)      # forks before CUDA is touched
    torch.cuda.set_device(0)
    runs = [frames_e2e(n, 0, modules, ci) for _ in range(max(1, args.steps))]
    best = min(runs, key=lambda r: r.get("seconds", 1e30))
    parity = {}
    if "error" not in best:
        from fitparity import classify_bazin, classify_r2
        from mallorn_astrophysics_b200.store import LightCurveStore
        with LightCurveStore.from_table(tb.subset(n_cpu)) as s:
            got = {m: s.features(m, z=tb.z[:n_cpu] if m == "physics" else None) for m in modules}
        for m in modules:
            g, w = got[m], cpu_rows[m]
            if m == "bazin":
                parity[m] = classify_bazin(g, w)
            elif m == "powerlaw":
                parity[m] = classify_r2(g, w)
            else:
                ok = ~np.isnan(w) & ~np.isnan(g)
                tol, atol = (1e-3, 1e-6) if m == "gp2d" else (1e-5, 1e-9)           # the rule of the parity tests
                bad = np.abs(g[ok] - w[ok]) > tol * np.abs(w[ok]) + atol
                parity[m] = {"cells": int(w.size), "nan_pattern_mismatch": int((np.isnan(g) != np.isnan(w)).sum()),
                             "beyond_tolerance": int(bad.sum()), "rtol": tol, "atol": atol}
    print(json.dumps({
        "metric": METRIC, "value": best.get("objects_per_s"), "unit": UNIT, "n_gpus": 1, "steps": len(runs), "warmup": 1,
        "ms_per_step": 1e3 * best.get("seconds", float("nan")), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "objects": n, "modules": list(modules), "columns": sum(MODULE_WIDTH[m] for m in modules),
                   "api": "features.pipeline.extract_feature_tables (DataFrames in, DataFrames out)"},
        "e2e": {"value": best.get("objects_per_s"), "unit": UNIT, "h2d_bytes_per_step": int(29 * tb.n_rows),
                "d2h_bytes_per_step": int(8 * n * sum(MODULE_WIDTH[m] for m in modules))},
        "frames": best,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {n_cpu} objects of the same set, modules {','.join(modules)}, {secs:.1f} s on {cores} processes"},
        "parity_vs_oracle": {"objects": n_cpu, "modules": parity},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="", choices=["", "c1", "c2", "c3", "c4", "c5"],
                    help="default: c4 on one GPU, c5 under torchrun")
    ap.add_argument("--objects", type=int, default=0,
                    help="c4: objects per GPU (default 10^6); c5: objects of the whole set (default 10^7); c1-c3: objects (default 10,178)")
    ap.add_argument("--batch-objects", type=int, default=0,
                    help="c5: objects one GPU takes per step (default 1.25 x 10^6 = an eighth of the set)")
    ap.add_argument("--modules", default="", help="comma list; default = every pass of the 434-column table")
    ap.add_argument("--cpu-objects", type=int, default=0, help="CPU sample size (default scales with host cores)")
    ap.add_argument("--full-cpu", action="store_true", help="c1-c3: run the CPU oracle on the whole set (minutes)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gp1d-objects", type=int, default=200_000,
                    help="objects of the batch the per-band GP pass (SURVEY 8f row 2) is timed on beside the step; 0 = skip")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config in ("c1", "c2", "c3"):
        run_small_config(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
