"""Downstream acceptance run on the GPU box:

    python tests/gpu_downstream_f1.py [seeds] > gpurun_out/r02_downstream_f1.json

Three feature arms on the 3,054 labelled objects of the 10,178-object C3 set, the 222-column v92d table assembled as the
reference's drivers assemble it (tests/downstream.py), 5-fold OOF F1 per seed:
  cpu       the CPU oracle (the reference's numpy / scipy arithmetic)
  gpu       this library on cuda:0 through the C-ABI
  cpu_libm  the CPU oracle with glibc's exp instead of numpy's in the Bazin model (<= 1 ulp apart): the reference
            against ITSELF -- the noise floor of the check
North star: |F1_gpu - F1_cpu| <= 0.005.  The JSON holds every per-seed number, the means and the floor; the committed
copy (profiles/r02_downstream_f1.json) is what tests/test_downstream_record.py asserts on.
"""
import os
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "1"          # one thread per worker process: the parallelism is across processes
import json
import multiprocessing as mp
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import downstream as D  # noqa: E402


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    classifier = sys.argv[2] if len(sys.argv) > 2 else "sklearn"          # "k10": the product's own GBDT (seconds instead of minutes)
    seeds = tuple(range(42, 42 + n_seeds))
    cores = os.cpu_count() or 1
    t0 = time.time()
    table = D.c3_train_table()
    pool = mp.get_context("fork").Pool(cores)              # forked before this process touches CUDA
    rows_cpu = D.oracle_rows(table, pool=pool, cores=cores)
    t_cpu = time.time() - t0
    print(f"cpu oracle: {t_cpu:.1f} s on {cores} processes", file=sys.stderr, flush=True)
    t0 = time.time()
    rows_libm = dict(rows_cpu)
    rows_libm["bazin"] = D.libm_bazin_rows(table, pool=pool, cores=cores)
    print(f"cpu oracle, glibc exp in the Bazin model: {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
    t0 = time.time()
    rows_gpu = D.gpu_rows(table)
    t_gpu = time.time() - t0
    print(f"gpu: {t_gpu:.2f} s", file=sys.stderr, flush=True)
    t0 = time.time()
    res = D.compare_arms(table, rows_cpu, rows_gpu, seeds=seeds, pool=pool, classifier=classifier)
    floor = D.compare_arms(table, rows_cpu, rows_libm, seeds=seeds, pool=pool, classifier=classifier)
    print(f"GBDT protocol, {2 * 2 * len(seeds) * 5} fits: {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
    pool.close()
    res["what"] = ("OOF F1 of the v34a/v92d-style GBDT on GPU features vs CPU (oracle) features, C3 set "
                   "(3,054 labelled objects of the 10,178-object set, seed 20260120), 222-column assembly of tests/downstream.py")
    res["cpu_seconds"], res["gpu_seconds"], res["cpu_cores"] = t_cpu, t_gpu, cores
    res["noise_floor"] = {"what": "the same comparison between the CPU oracle and the CPU oracle with glibc's exp in the Bazin model "
                                  "(the reference against itself, <= 1 ulp apart)",
                          **{k: floor[k] for k in ("mean_delta", "mean_abs_delta", "max_abs_delta", "std_delta", "cells_differing_beyond_1e-3")},
                          "deltas": {k: v["delta"] for k, v in floor["seeds"].items()}}
    res["north_star_bound"] = 0.005
    res["meets_north_star_per_seed"] = bool(res["max_abs_delta"] <= 0.005)
    res["mean_delta_standard_error"] = float(res["std_delta"] / np.sqrt(len(seeds)))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
