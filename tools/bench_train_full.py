#!/usr/bin/env python
"""BASELINE.json configs[2]: ridge-regression training of ALL 1152 W_out on one B200 -- Gram accumulation + solve
in region waves sized to HBM.

Every region is trained at the reference's default length: 6 interleaved phases (trainingdata(:, i::6),
src/mod_reservoir.f90:289-301) of `--cols` columns each (2000 -> 1960 kept states per phase after the 40-column
discard, batch 98), then fit_chunk_hybrid (beta_res = 1e-3, beta_model = 1, squared).  The synthetic series is one
standardised AR(1) draw per phase shared by the regions of a wave (the arithmetic does not depend on the values).
Reports wall time, CUDA-event time of the Gram and solve kernels, useful Gram TFLOP/s.  One JSON line.
"""
import argparse
import importlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wave", type=int, default=192)
    ap.add_argument("--regions", type=int, default=1152)
    ap.add_argument("--phases", type=int, default=6)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--discard", type=int, default=40)
    ap.add_argument("--batch", type=int, default=98)
    ap.add_argument("--global-series", action="store_true",
                    help="feed from the device-resident global series (sml_train_global_series) instead of per-region series")
    ap.add_argument("--overlap", action="store_true", help="state generation overlaps the previous slab's Gram (default: serial schedule)")
    ap.add_argument("--no-overlap", action="store_true", help="(default) serial schedule: state generation, then Gram, per slab")
    args = ap.parse_args()
    E = importlib.import_module("speedy-ml_b200.engine")
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=1152 // args.regions)
    regions = eng.region_indices
    dims = {}
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=16) as ex:
        for w in ex.map(bench.gen_region, regions):
            eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], None, w["mean"], w["std"],
                              win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"],
                              S=w["S"], P=w["P"])
            dims[w["region"]] = (w["D"], w["S"], w["n"], w["P"])
    eng.finalize()
    eng.train_set_overlap(args.overlap and not args.no_overlap)
    setup = time.perf_counter() - t0
    rng = np.random.default_rng(3)
    up_s = 0.0
    if args.global_series:
        # one resident series of cols + phases - 1 columns; phase i reads columns i, i+1, ... (the arithmetic does not
        # depend on the values, and this keeps the host array at ~5 GB instead of 30 GB)
        lay = E.global_layout()
        F0 = bench.initial_fields()
        Tt = args.cols + args.phases - 1
        base = np.concatenate([F0["clim4d"].ravel(order="F"), F0["clim2d"].ravel(order="F"), np.full(96 * 48, 0.3),
                               np.maximum(F0["base_sst"], 272.0).ravel(order="F"), F0["tisr"].ravel(order="F")])
        G = np.empty((lay["g_total"], Tt), order="F")
        Fs = np.empty((lay["f_total"], Tt), order="F")
        for t in range(Tt):
            G[:, t] = base * (1.0 + 0.01 * np.sin(0.37 * t + np.arange(base.size) * 1e-3))
            Fs[:, t] = 0.98 * G[:lay["f_total"], t]
        tu = time.perf_counter()
        eng.train_global_series(G, Fs)
        up_s = time.perf_counter() - tu
        del G, Fs
    D_max = max(d[0] for d in dims.values())
    S_max = max(d[1] for d in dims.values())
    phases = [(syn.ar1_series(D_max, args.cols, rng), np.asfortranarray(rng.standard_normal((S_max, args.cols))))
              for _ in range(args.phases)]
    tot = dict(gram_flops_useful=0.0, gram_ms=0.0, stategen_ms=0.0, solve_ms=0.0)
    by_chol = 0
    bad = 0
    t0 = time.perf_counter()
    host = dict(begin_s=0.0, feed_s=0.0, solve_s=0.0, end_s=0.0)   # wall clock of the calls, per kind
    for i0 in range(0, len(regions), args.wave):
        wave = regions[i0:i0 + args.wave]
        ta = time.perf_counter()
        eng.train_begin(wave, args.batch)
        tb = time.perf_counter()
        for ph, (td, im) in enumerate(phases):
            if args.global_series:
                eng.train_feed_global(ph, 1, args.cols, args.discard)
            else:
                eng.train_feed([td[:dims[r][0]] for r in wave], [im[:dims[r][1]] for r in wave], args.discard)
        tc = time.perf_counter()
        info = eng.train_solve(1e-3, 1.0, True, 0.0)
        td_ = time.perf_counter()
        bad += int(np.count_nonzero(info))
        by_chol += eng.train_solver_stats()
        st = eng.train_stats()
        for k in tot:
            tot[k] += st[k]
        eng.train_end()
        te = time.perf_counter()
        host["begin_s"] += tb - ta
        host["feed_s"] += tc - tb
        host["solve_s"] += td_ - tc
        host["end_s"] += te - td_
    wall = time.perf_counter() - t0
    w = eng.wout_get(regions[-1])
    out = {"workload": f"ridge training of {len(regions)} W_out (m=6000), {args.phases} phases x {args.cols} columns, "
                       f"waves of {args.wave}", "wall_s": wall, "gram_s": tot["gram_ms"] / 1e3,
           "stategen_s": tot["stategen_ms"] / 1e3, "solve_s": tot["solve_ms"] / 1e3,
           "gram_tflops_useful": tot["gram_flops_useful"] / (tot["gram_ms"] * 1e-3) / 1e12,
           "solve_ms_per_region": tot["solve_ms"] / len(regions), "solved_by_cholesky": by_chol, "dgesv_info_nonzero": bad,
           "wout_finite": bool(np.isfinite(w).all()), "setup_s": round(setup, 1),
           "schedule": "state generation overlaps the previous slab's Gram (stategen_s and gram_s overlap in time)" if (args.overlap and not args.no_overlap) else "serial",
           "host_wall": {k: round(v, 3) for k, v in host.items()},
           "feed": "device-resident global series" if args.global_series else "per-region host series",
           "global_series_upload_s": up_s}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
