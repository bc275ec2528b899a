#!/usr/bin/env python
"""BASELINE.json configs[4]: reservoir-size and adjacency-degree sweep on one B200.

For each (m, degree): 1152 regions of the T30 tiling (true class mix, SST input on 70 %), synthetic weights;
  * update  : the state update alone (ELL SpMV + compact W_in + tanh + leak) -- the update-only launches of k_step
              that synchronize(ALL) issues, bracketed by CUDA events inside the engine;
  * step    : the fused update + readout kernel (CUDA events inside the engine, as bench.py does);
  * readout : step - update.
Bytes are the algorithmic ones of DESIGN.md section 4.1; the roof is MEASURED_PEAKS.json's copy bandwidth.
One JSON line per point on stdout; --md writes the table for profiles/.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def gen(region, m, deg):
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    E = importlib.import_module("speedy-ml_b200.engine")
    sst_in = bench.sst_input_mask(region)
    d = E.region_dims(1152, region, 1, m, float(deg), True, True, sst_in, False)
    rng = np.random.default_rng(20251018 + region)
    rows, cols, vals = syn.make_adjacency(d["n"], d["k"], rng, radius=0.7, power_iters=10)
    winc, wcol = syn.make_win_compact(d["n"], d["D"], rng, sigma=0.5)
    N = d["n"] + d["S"]
    wout = np.empty((d["P"], N), order="F")
    wout.reshape(-1, order="F")[:] = (rng.random(d["P"] * N) - 0.5) * (np.sqrt(12.0) / np.sqrt(N))
    mean, std = syn.make_mean_std(d["L"], rng)
    return dict(region=region, sst_bool_input=sst_in, rows=rows, cols=cols, vals=vals, winc=winc, wcol=wcol, wout=wout,
                mean=mean, std=std, **d)


def point(m, deg, nregions, steps):
    from concurrent.futures import ThreadPoolExecutor
    E = importlib.import_module("speedy-ml_b200.engine")
    world = 1152 // nregions
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=world, sst_prescribed=True)
    upd_bytes = rd_bytes = 0
    dims = []
    with ThreadPoolExecutor(max_workers=16) as ex:
        for w in ex.map(lambda r: gen(r, m, deg), eng.region_indices):
            eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                              win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
            width = int(np.bincount(w["rows"] - 1, minlength=w["n"]).max())
            upd_bytes += 12 * width * w["n"] + 16 * w["n"] + 12 * w["n"] + 8 * w["D"]
            rd_bytes += 8 * w["P"] * (w["n"] + w["S"]) + 8 * w["S"] + 8 * w["P"] + 16 * w["L"]
            dims.append((w["n"], w["D"]))
    eng.finalize()
    assert eng.predict_algorithmic_bytes() == upd_bytes + rd_bytes
    rng = np.random.default_rng(0)
    inputs = [np.asfortranarray(rng.standard_normal((D, steps))) for (_, D) in dims]
    os.environ["SML_SYNC_KERNEL"] = "steps"
    eng.synchronize_all(inputs, 3)               # warm-up
    eng.profile(True)
    eng.synchronize_all(inputs, steps)           # update-only launches, bracketed by CUDA events in the engine
    upd_ms, nsteps = eng.sync_times()
    eng.profile(False)
    upd_ms /= nsteps
    os.environ.pop("SML_SYNC_KERNEL")
    # the default synchronize: the whole time loop in one launch (k_sync_persist), `spin_steps` steps
    spin_steps = 100
    spin_in = [np.asfortranarray(np.tile(a, (1, (spin_steps + steps - 1) // steps))[:, :spin_steps]) for a in inputs]
    n0 = eng.kernel_launch_count()
    eng.synchronize_all(spin_in, 4)
    one_launch = eng.kernel_launch_count() - n0 <= 2
    eng.profile(True)
    eng.synchronize_all(spin_in, spin_steps)
    spin_ms, nspin = eng.sync_times()
    eng.profile(False)
    spin_ms /= nspin
    del spin_in
    for _ in range(3):
        eng.predict()
    eng.profile(True)
    for _ in range(steps):
        eng.predict()
    step_ms, _, cnt = eng.kernel_times()
    eng.profile(False)
    step_ms /= cnt
    eng.close()
    peak, _ = bench.measured_peak()
    rd_ms = max(step_ms - upd_ms, 1e-9)
    return {"m": m, "degree": deg, "regions": nregions, "n_typ": dims[len(dims) // 2][0],
            "update_ms": upd_ms, "update_GBs": upd_bytes / upd_ms / 1e6, "update_frac": upd_bytes / upd_ms / 1e6 / peak,
            "spin_ms_per_step": spin_ms, "spin_one_launch": one_launch, "spin_vs_roof": upd_bytes / spin_ms / 1e6 / peak,
            "step_ms": step_ms, "step_GBs": (upd_bytes + rd_bytes) / step_ms / 1e6,
            "step_frac": (upd_bytes + rd_bytes) / step_ms / 1e6 / peak,
            "readout_ms": rd_ms, "readout_GBs": rd_bytes / rd_ms / 1e6, "readout_frac": rd_bytes / rd_ms / 1e6 / peak,
            "update_MB": upd_bytes / 1e6, "readout_MB": rd_bytes / 1e6, "peak_GBs": peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, nargs="*", default=[2000, 4000, 6000, 8000, 12000])
    ap.add_argument("--deg", type=int, nargs="*", default=[3, 6, 12, 24])
    ap.add_argument("--regions", type=int, default=1152)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    rows = []
    for m in args.m:
        for deg in args.deg:
            if deg != 6 and m != 6000 and not (m in (2000, 12000) and deg in (3, 24)):
                continue  # full degree sweep at m=6000, size sweep at degree 6, plus the four corners
            r = point(m, deg, args.regions, args.steps)
            rows.append(r)
            print(json.dumps(r), flush=True)
    if args.md:
        with open(args.md, "w") as f:
            f.write("| m | degree | update MB | update ms (step launches) | update GB/s | frac | spin-up ms per step (one launch, 100 steps) | x HBM roof "
                    "| readout MB | readout ms | readout GB/s | frac | fused step ms | step GB/s | frac |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['m']} | {r['degree']} | {r['update_MB']:.0f} | {r['update_ms']:.3f} | {r['update_GBs']:.0f} | "
                        f"{r['update_frac']:.2f} | {r['spin_ms_per_step']:.3f}{'' if r['spin_one_launch'] else ' (step launches)'} | {r['spin_vs_roof']:.2f} | {r['readout_MB']:.0f} | {r['readout_ms']:.3f} | {r['readout_GBs']:.0f} | "
                        f"{r['readout_frac']:.2f} | {r['step_ms']:.3f} | {r['step_GBs']:.0f} | {r['step_frac']:.2f} |\n")


if __name__ == "__main__":
    main()
