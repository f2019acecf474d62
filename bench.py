#!/usr/bin/env python
"""bench.py -- objects/s of the MALLORN light-curve featurisation on B200 (contract in the task statement).

A *step* is one pass of the hot path over one batch of synthetic light curves: K1 pack (flat
long-format rows -> CSR (object, band) segments in HBM) followed by every feature pass the
library exports (statistical, colors, shape, tde_physics, physics_based, bazin, power-law,
2-D GP), i.e. the full 434-column module table the v34a/v92d 222-feature set is selected from.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--objects M] [--impl reference]

* ``value``  objects/s, inputs resident in HBM (flat columns on the device) when the clock starts.
* ``e2e``    the same through the C-ABI with HOST buffers: H2D of the five input columns and D2H of
             every feature table inside the timed region.
* ``roofline`` the moment pass (north star: judged against HBM bandwidth), CUDA-event timed inside
             the step; ``passes`` lists every pass's share of the step.
* ``cpu_baseline`` the CPU oracle (``oracle/``: the reference's numpy/scipy arithmetic) on a bounded
             sample of the same workload, all host cores.
* ``widened``  SURVEY 8(f) rows built so far, measured beside the step and not part of ``value`` (the 434-column
             table): the per-band 1-D GP pass (``extract_gp_features``, 21 columns) on a slice of the same batch, and
             the native CSV ingest against the reference's pandas loading path on the box's host cores, and the
             DataFrame-in / DataFrame-out call (``extract_feature_tables``) on the 10,178-object C1-shaped set.
* ``--impl reference`` times only that CPU path (the reference has no GPU path and, being a Python
             package with no native code, "the reference's own CPU implementation" is the oracle
             restatement: ``/root/reference`` does not exist on the GPU box).

N > 1: launched by torchrun, one rank per GPU; objects are sharded (weak scaling, fixed objects per
GPU), no collective during compute, one NCCL all-gather of the feature table at the end of each step.
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
SEED = 20260118 + 3            # seed = 20260118 + config index (C4)
# DRAM traffic of the moment pass from `ncu --set full` (profiles/r01_stat_v2_full.txt: 793.78 MB read + 184.05 MB
# written for 200,000 objects of this generator) -- 0.99x the algorithmic bytes, i.e. no re-reads
NCU_STAT_DRAM_BYTES_PER_OBJECT = (793.778688e6 + 184.050432e6) / 200_000


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
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in modules:
            run_module(m, obj, cols["band"][keep], cols["t"][keep], cols["flux"][keep], cols["err"][keep],
                       hi - lo, cols["z"][lo:hi])
    return hi - lo


def cpu_rate(cols, n_objects, modules, cores):
    """objects/s of the CPU oracle on the first ``n_objects`` objects, ``cores`` worker processes."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    per = max(1, (n_objects + cores * 2 - 1) // (cores * 2))
    jobs = [(cols, lo, min(lo + per, n_objects), modules) for lo in range(0, n_objects, per)]
    t0 = time.perf_counter()
    if cores == 1:
        done = sum(_cpu_chunk(j) for j in jobs)
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            done = sum(pool.map(_cpu_chunk, jobs))
    dt = time.perf_counter() - t0
    return done / dt, dt


def host_sample(n_objects):
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(n_objects, seed=SEED)
    return {"obj": tb.obj, "band": tb.band, "t": tb.t, "flux": tb.flux, "err": tb.err, "z": tb.z}


def ingest_rates(n_objects=2000):
    """SURVEY 8(f) row 1 measured on this box's host cores: the reference's loading path (pd.read_csv of the split
    files + concat, then the factorisation) against the native ingest (libmallorn_ingest.so) on the same files."""
    import shutil
    import tempfile
    import pandas as pd
    from mallorn_astrophysics_b200 import ingest
    from mallorn_astrophysics_b200.store import frame_columns
    from mallorn_astrophysics_b200.synth import make_table
    root = tempfile.mkdtemp(prefix="mlrn_bench_ingest_")
    try:
        lc = make_table(n_objects, seed=SEED).to_frame()
        edges = np.linspace(0, len(lc), 5).astype(int)
        paths = []
        for i in range(4):
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
        return {"reference": "load_lightcurves (data_loader.py:36-62) + groupby factorisation", "rows": int(len(lc)),
                "csv_bytes": int(size), "pandas_mb_per_s": size / 1e6 / t_ref, "native_mb_per_s": size / 1e6 / t_nat,
                "native_rows_per_s": len(lc) / t_nat, "threads": os.cpu_count(), "bit_identical_to_pandas": same}
    except Exception as exc:                                                # never lose the bench line over the side measurement
        return {"error": f"{type(exc).__name__}: {exc}"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def frames_e2e(n_objects=10178, device=0):
    """The user-facing call itself on the C1-sized set (BASELINE.json configs[0] shape): the reference's long DataFrame in,
    the reference's per-module DataFrames out (features.pipeline.extract_feature_tables = one K1 pack, the eight passes of
    the 434-column table), host factorisation and DataFrame assembly included."""
    try:
        from mallorn_astrophysics_b200.features.pipeline import extract_feature_tables
        from mallorn_astrophysics_b200.synth import make_table
        tb = make_table(n_objects, config_index=1)
        lc, meta, ids = tb.to_frame(), tb.metadata_frame(), list(tb.object_ids)
        warm = ids[:512]
        extract_feature_tables(lc[lc["object_id"].isin(set(warm))], meta, warm, modules=MODULES, device=device)
        t0 = time.perf_counter()
        tables = extract_feature_tables(lc, meta, ids, modules=MODULES, device=device)
        dt = time.perf_counter() - t0
        return {"call": "features.pipeline.extract_feature_tables(lightcurves, metadata, object_ids)", "objects": int(n_objects),
                "rows": int(len(lc)), "seconds": dt, "objects_per_s": n_objects / dt,
                "columns": int(sum(len(t.columns) - 1 for t in tables.values()))}
    except Exception as exc:                                                # never lose the bench line over the side measurement
        return {"error": f"{type(exc).__name__}: {exc}"}


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
    sample = f"{per_step} objects/step of the C4 generator (seed {SEED}), {len(cols['obj'])} observations"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, args.objects, modules), sample_objects_per_step=per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, objects_per_gpu, modules):
    return {"workload": "C4: 1M-object LSST-like synthetic set, full feature table (BASELINE.json configs[3])"
            if objects_per_gpu == 1_000_000 else
            f"C4 generator at {objects_per_gpu} objects per GPU (BASELINE.json configs[3] shape)",
            "objects_per_gpu": objects_per_gpu, "bands": 6, "mean_obs_per_object": 150,
            "modules": list(modules), "columns": sum(MODULE_WIDTH[m] for m in modules),
            "l2": "inputs larger than L2 (no flush needed)" if objects_per_gpu * 150 * 25 > 2 * 126e6 else
                  "L2 flushed between steps",
            "parallelism": f"objects sharded over {args.gpus} GPU(s), all-gather of the feature table"}


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
    from mallorn_astrophysics_b200 import _native
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_columns_device

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
    n_obj = args.objects
    cols = make_columns_device(n_obj, SEED * 131 + rank, dev)
    n_rows = int(cols["obj"].shape[0])
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    outs = {m: torch.empty((n_obj, _native.COLS[m]), dtype=torch.float64, device=dev) for m in modules}
    width = sum(_native.COLS[m] for m in modules)
    table = torch.empty((n_obj, width), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * n_obj, width), dtype=torch.float64, device=dev) if world > 1 else None
    gp_status = torch.zeros((n_obj, 4), dtype=torch.int32, device=dev) if "gp2d" in modules else None
    gp1_status = torch.zeros((n_obj, 12), dtype=torch.int32, device=dev) if "gp1d" in modules else None
    flush = None
    if n_rows * 25 < 2 * 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    pass_ms = {m: 0.0 for m in modules}
    pass_ms["pack"] = 0.0
    pass_ms["allgather"] = 0.0

    def step(timed: bool):
        marks = [("start", ev())]
        marks[0][1].record(stream)
        store = LightCurveStore(cols["obj"], cols["band"], cols["t"], cols["flux"], cols["err"], n_obj,
                                device=local, stream=sptr)
        e = ev(); e.record(stream); marks.append(("pack", e))
        c0 = 0
        for m in modules:
            store.features(m, z=cols["z"] if m == "physics" else None, out=outs[m],
                           status=gp_status if m == "gp2d" else (gp1_status if m == "gp1d" else None))
            w = _native.COLS[m]
            table[:, c0:c0 + w].copy_(outs[m])
            c0 += w
            e = ev(); e.record(stream); marks.append((m, e))
        if world > 1:
            dist.all_gather_into_tensor(gathered, table)
            e = ev(); e.record(stream); marks.append(("allgather", e))
        stream.synchronize()
        store.close()
        if timed:
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                pass_ms[name] += a.elapsed_time(b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    launches0 = L.mlrn_launch_count()
    barrier()
    with ClockSampler(local) as clk:
        t_start, t_end = ev(), ev()
        t_start.record(stream)
        for _ in range(args.steps):
            if flush is not None:
                flush.zero_()
            step(True)
        t_end.record(stream)
        barrier()
    launches = L.mlrn_launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * n_obj * args.steps / (ms * 1e-3)

    # ---- end to end: host buffers through the C-ABI, copies inside the timed region
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in cols.items()}
    hnp = {k: v.numpy() for k, v in host.items()}
    houts = {m: torch.empty((n_obj, _native.COLS[m]), dtype=torch.float64, pin_memory=True).numpy() for m in modules}
    h2d = sum(hnp[k].nbytes for k in ("obj", "band", "t", "flux", "err")) + (hnp["z"].nbytes if "physics" in modules else 0)
    d2h = sum(a.nbytes for a in houts.values())

    def e2e_step():
        store = LightCurveStore(hnp["obj"], hnp["band"], hnp["t"], hnp["flux"], hnp["err"], n_obj,
                                device=local, stream=sptr)
        for m in modules:
            store.features(m, z=hnp["z"] if m == "physics" else None, out=houts[m])
        store.close()

    for _ in range(min(args.warmup, 3)):
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
    e2e_value = world * n_obj * e2e_steps / e2e_s

    # ---- section 8(f) row measured beside the step: the per-band 1-D GP pass (extract_gp_features), on a bounded slice
    widened = None
    if hasattr(L, "mlrn_gp1d") and "gp1d" not in modules and not args.modules and args.gp1d_objects > 0:
        k = min(n_obj, args.gp1d_objects)
        keep = cols["obj"] < k
        sub = {c: cols[c][keep].contiguous() for c in ("obj", "band", "t", "flux", "err")}
        o1 = torch.empty((k, _native.COLS["gp1d"]), dtype=torch.float64, device=dev)
        s1 = torch.zeros((k, 12), dtype=torch.int32, device=dev)
        st1 = LightCurveStore(sub["obj"], sub["band"], sub["t"], sub["flux"], sub["err"], k, device=local, stream=sptr)
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
                            "observations": int(sub["obj"].shape[0]), "launches": 1}}
        del sub, o1, s1

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        stat_ms = pass_ms["statistical"] / args.steps if "statistical" in modules else None
        roof = None
        if stat_ms:
            ach = algorithmic_bytes(n_rows, n_obj) / (stat_ms * 1e-3) / 1e9
            roof = {"kernel": "k_stat (moment pass, stat_fast.cu)", "bound": "hbm", "achieved": ach, "peak": peak,
                    "unit": "GB/s", "frac": ach / peak,
                    "traffic": NCU_STAT_DRAM_BYTES_PER_OBJECT * n_obj,
                    "traffic_source": "ncu --set full capture at 200k objects (profiles/r01_stat_v2_full.txt), scaled per object",
                    "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch": algorithmic_bytes(n_rows, n_obj), "ms_per_launch": stat_ms}
        roof_gp = None
        if gp_status is not None and pass_ms.get("gp2d"):
            # dominant kernel of the step: FP64 work of the GP = (evaluations + the one at the optimum) x N^3 flop
            # (N^3/2 FMA per -lnL + gradient evaluation) over its CUDA-event time
            stt = gp_status.to(torch.float64)
            fitted = stt[:, 3] >= 10
            flop = float((((stt[:, 1] + 1.0) * stt[:, 3] ** 3)[fitted]).sum())
            gp_ms = pass_ms["gp2d"] / args.steps
            ach_tf = flop / (gp_ms * 1e-3) / 1e12
            roof_gp = {"kernel": "k_gp2d (2-D GP, 61 % of the step)", "bound": "tensor", "achieved": ach_tf, "peak": 37.1,
                       "unit": "TFLOP/s", "frac": ach_tf / 37.1, "traffic": None,
                       "peak_source": "FP64 DMMA/DFMA rate measured on this pool's B200 with tests/micro/fp64_rates.cu "
                                      "(64 FMA/clk/SM); MEASURED_PEAKS.json has no FP64 figure",
                       "algorithmic_flop_per_launch": flop, "ms_per_launch": gp_ms,
                       "mean_evaluations": float(stt[:, 1][fitted].mean()), "mean_points": float(stt[:, 3][fitted].mean())}
        cpu = None
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_objects or 48 * cores
            sample_cols = host_sample(n_cpu)
            rate, secs = cpu_rate(sample_cols, n_cpu, modules, cores)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_cpu} objects of the same generator (seed {SEED}), modules {','.join(modules)}, "
                             f"{secs:.1f} s on {cores} processes"}
            if widened is not None:
                n_w = min(n_cpu, 8 * cores)
                rate_w, secs_w = cpu_rate(sample_cols, n_w, ["gp1d"], cores)
                widened["gp1d"]["cpu_baseline"] = {"value": rate_w, "unit": UNIT, "cores": cores, "kind": "port",
                                                   "sample": f"first {n_w} objects of the CPU sample, oracle/gp1d.py, {secs_w:.1f} s"}
                widened["ingest"] = ingest_rates()
                widened["frames_e2e"] = frames_e2e(device=local)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_obj, modules),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_dominant": roof_gp, "cpu_baseline": cpu,
            "passes": {k: {"ms_per_step": v / args.steps, "share": v / ms} for k, v in pass_ms.items() if v > 0},
            "observations_per_gpu": n_rows,
            "widened": widened,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--objects", type=int, default=1_000_000, help="objects per GPU per step")
    ap.add_argument("--modules", default="", help="comma list; default = every pass the library exports")
    ap.add_argument("--cpu-objects", type=int, default=0, help="CPU sample size (default scales with host cores)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gp1d-objects", type=int, default=200_000,
                    help="objects of the batch the per-band GP pass (SURVEY 8f row 2) is timed on beside the step; 0 = skip")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
