#!/usr/bin/env python
"""Training-path measurement (BASELINE.json configs[2], second headline metric: Gram FP64 TFLOP/s).

A wave of full-size regions (m=6000) is trained for one phase of `--cols` columns; reports
  * Gram accumulation TFLOP/s on USEFUL flops N(N+1)K + 2PNK (the symmetric half + Y*R^T),
  * state-generation and solve times,
  * the box's cuBLAS DGEMM rate (torch.matmul float64 8192^3) as the measured FP64 roof.
One JSON line on stdout.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def cublas_dgemm_tflops(n=8192, reps=5):
    import torch
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=8)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--discard", type=int, default=40)
    ap.add_argument("--batch", type=int, default=98)
    ap.add_argument("--solve", action="store_true")
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--overlap", action="store_true",
                    help="default schedule of the engine (state generation overlaps the previous slab's Gram); "
                         "off here so that every kernel is timed alone")
    ap.add_argument("--phases", type=int, default=1)
    args = ap.parse_args()
    E = importlib.import_module("speedy-ml_b200.engine")
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    nreg = args.regions
    eng = E.Engine(number_of_regions=1152, irank=1, numprocs=1152 // nreg)
    regions = eng.region_indices
    ws = {}
    for r in regions:
        w = bench.gen_region(r)
        ws[r] = w
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], None, w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"], S=w["S"], P=w["P"])
    eng.finalize()
    rng = np.random.default_rng(1)
    tds = [syn.ar1_series(ws[r]["D"], args.cols, rng) for r in regions]
    ims = [np.asfortranarray(rng.standard_normal((ws[r]["S"], args.cols))) for r in regions]
    eng.train_set_overlap(args.overlap)
    eng.train_begin(regions, args.batch)
    t0 = time.perf_counter()
    for _ in range(args.phases):
        eng.train_feed(tds, ims, args.discard)
    st = eng.train_stats()                      # waits for both streams
    wall_feed = time.perf_counter() - t0
    out = {"workload": f"ridge training, {nreg} regions x {args.phases} phase(s) x {args.cols} columns, m=6000",
           "gram_tflops_useful": st["gram_flops_useful"] / (st["gram_ms"] * 1e-3) / 1e12,
           "gram_ms": st["gram_ms"], "stategen_ms": st["stategen_ms"], "feed_wall_s": wall_feed,
           "kept_columns": (args.cols - args.discard) // args.batch * args.batch, "phases": args.phases,
           "schedule": "overlap" if args.overlap else "serial"}
    if args.solve:
        info = eng.train_solve(1e-3, 1.0, True, 0.0)
        st = eng.train_stats()
        out["solve_ms_per_region"] = st["solve_ms"] / nreg
        out["solve_info"] = [int(i) for i in info]
    eng.train_end()
    eng.close()
    if not args.no_cublas:
        out["cublas_dgemm_tflops_8192"] = cublas_dgemm_tflops()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
