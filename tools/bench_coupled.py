#!/usr/bin/env python
"""BASELINE.json configs[3]: coupled hybrid atmosphere + ML slab-ocean reservoirs, long run on one B200.

Config 2's model (1152 atmosphere reservoirs, m = 6000) plus an ocean reservoir (m = 4000: n = 3968, D = 128, P = 8)
on the 70 % of regions that carry an SST input; the ocean reservoirs step when mod(t*6, 168) == 0
(src/parallelmain.f90:238), their feedback is the 27-slot ring mean + the SST tile every step.  Device-resident
steps (host model excluded, as bench.py's `value`); one JSON line.
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


def gen_ocean(region):
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    E = importlib.import_module("speedy-ml_b200.engine")
    d = E.ocean_region_dims(1152, region)
    rng = np.random.default_rng(20251018 + 100000 + region)
    k = int((6.0 / 4000.0) * d["n"] * d["n"])
    rows, cols, vals = syn.make_adjacency(d["n"], k, rng, radius=0.9, power_iters=20)
    winc, wcol = syn.make_win_compact(d["n"], d["D"], rng, sigma=0.6)
    wout = np.asfortranarray((rng.random((d["P"], d["n"])) - 0.5) * (np.sqrt(12.0) / np.sqrt(d["n"])))
    return dict(region=region, rows=rows, cols=cols, vals=vals, winc=winc, wcol=wcol, wout=wout, **d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2920)   # 2 sim-years
    ap.add_argument("--warmup", type=int, default=56)
    args = ap.parse_args()
    import torch
    E = importlib.import_module("speedy-ml_b200.engine")
    H = importlib.import_module("speedy-ml_b200.hybrid")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = E.Engine(number_of_regions=1152, sst_prescribed=False, stream=stream)
    t0 = time.perf_counter()
    ocean_regions = [r for r in range(1152) if bench.sst_input_mask(r)]
    with ThreadPoolExecutor(max_workers=16) as ex:
        for w in ex.map(bench.gen_region, range(1152)):
            eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                              win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
            if w["region"] == 0:
                mean0, std0 = w["mean"], w["std"]
        for w in ex.map(gen_ocean, ocean_regions):
            # grid_special copies the atmosphere reservoir's mean/std; only the SST slot matters on this path
            eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], mean0, std0,
                              win_compact=w["winc"], win_col=w["wcol"], D=w["D"], kind=E.OCEAN,
                              sst_mean=290.0, sst_std=6.0)
    eng.finalize()
    setup = time.perf_counter() - t0
    F = bench.initial_fields()
    eng.set_sst_static(F["base_sst"], F["sea_mask"])
    rng = np.random.default_rng(1)
    for r in ocean_regions:
        eng.outvec_set(r, 285.0 + 5.0 * rng.random(8), kind=E.OCEAN)        # start_prediction_slab's seed
        eng.feedback_set(r, rng.standard_normal(eng.dims[(E.OCEAN, r)]["D"]), kind=E.OCEAN)
    shard = H.EngineShard(eng, torch, ocean=True)
    stepper = H.HybridStepper(shard)
    lay = E.global_layout()
    g0 = np.concatenate([F["clim4d"].ravel(order="F"), F["clim2d"].ravel(order="F"), np.zeros(96 * 48),
                         np.maximum(F["base_sst"], 272.0).ravel(order="F"), F["tisr"].ravel(order="F")])
    shard.G.copy_(torch.from_numpy(g0))
    f4, f2 = bench.host_stub(F["clim4d"], F["clim2d"], F["clim4d"], F["clim2d"])
    shard.F.copy_(torch.from_numpy(np.concatenate([f4.ravel(order="F"), f2.ravel(order="F")])))
    eng.step_unpack_device(1)
    for t in range(1, args.warmup + 1):
        stepper.device_step(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.kernel_launch_count()
    e0.record(stream)
    for t in range(args.warmup + 1, args.warmup + 1 + args.steps):
        stepper.device_step(t)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ocean_steps = sum(1 for t in range(args.warmup + 1, args.warmup + 1 + args.steps) if H.ocean_step_due(t))
    x = eng.state_get(ocean_regions[0], kind=E.OCEAN)
    out = {"workload": "coupled hybrid atmosphere + slab-ocean reservoirs (BASELINE configs[3]), 1152 + "
                       f"{len(ocean_regions)} reservoirs, device-resident", "steps": args.steps,
           "ocean_steps": ocean_steps, "ms_per_step": ms / args.steps,
           "sim_days_per_s": 0.25 / (ms / args.steps * 1e-3), "sim_years": args.steps * 0.25 / 365.0,
           "gpu_launches": eng.kernel_launch_count() - launches0, "ocean_state_finite": bool(np.isfinite(x).all()),
           "atmo_bytes_per_step": eng.predict_algorithmic_bytes(E.ATMO),
           "ocean_bytes_per_ocean_step": eng.predict_algorithmic_bytes(E.OCEAN), "setup_s": round(setup, 1)}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
