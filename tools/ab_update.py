#!/usr/bin/env python
"""A/B of the update-only kernels (synchronize path) on one engine: k_update<RPT> (x gathered from L2), k_update_sx<RPT>
(x staged in shared memory) at several row splits, and k_update_ring (round 2: the adjacency streamed by a TMA producer warp
through a shared-memory ring) at several ring depths / tile sizes.  The engine re-reads SML_UPDATE_KERNEL /
SML_UPDATE_RPT / SML_UPDATE_SPLIT at every launch, so one set-up serves all variants.  One JSON line per variant."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import sweep  # noqa: E402


def main():
    from concurrent.futures import ThreadPoolExecutor
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=6000)
    ap.add_argument("--deg", type=int, default=6)
    ap.add_argument("--regions", type=int, default=1152)
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    E = importlib.import_module("speedy-ml_b200.engine")
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=1152 // args.regions, sst_prescribed=True)
    upd_bytes = 0
    dims = []
    with ThreadPoolExecutor(max_workers=16) as ex:
        for w in ex.map(lambda r: sweep.gen(r, args.m, args.deg), eng.region_indices):
            eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                              win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
            width = int(np.bincount(w["rows"] - 1, minlength=w["n"]).max())
            upd_bytes += 12 * width * w["n"] + 16 * w["n"] + 12 * w["n"] + 8 * w["D"]
            dims.append((w["n"], w["D"]))
    eng.finalize()
    rng = np.random.default_rng(0)
    inputs = [np.asfortranarray(rng.standard_normal((D, args.steps))) for (_, D) in dims]
    peak, _ = bench.measured_peak()
    # (kernel, rows per thread, row splits, threads | ring: stages, tile rows)
    variants = [("global", 4, 0, 0), ("sx", 2, 0, 0), ("sx", 2, 2, 0), ("ring", 2, 0, 0), ("ring", 3, 0, 0), ("ring", 4, 0, 0),
                ("ring", 3, 2, 0), ("ring", 4, 0, 320), ("ring", 6, 0, 256), ("global", 4, 0, 0)]
    ref = None
    for kern, rpt, split, threads in variants:
        os.environ["SML_UPDATE_KERNEL"] = kern
        os.environ["SML_UPDATE_RPT"] = str(rpt)
        os.environ.pop("SML_UPDATE_STAGES", None)
        os.environ.pop("SML_UPDATE_TILE_ROWS", None)
        if kern == "ring":
            os.environ["SML_UPDATE_STAGES"] = str(rpt)
            os.environ.pop("SML_UPDATE_RPT", None)
            if threads:
                os.environ["SML_UPDATE_TILE_ROWS"] = str(threads)
                threads = 0
        for key, val in (("SML_UPDATE_SPLIT", split), ("SML_UPDATE_THREADS", threads)):
            if val:
                os.environ[key] = str(val)
            else:
                os.environ.pop(key, None)      # 0: the engine's own choice
        probe = [eng.region_indices[i] for i in (0, 7, len(dims) - 1)]
        for r in probe:
            eng.state_set(r, np.zeros(eng.dims[(E.ATMO, r)]["n"]))
        eng.synchronize_all(inputs, 3)
        eng.profile(True)
        eng.synchronize_all(inputs, args.steps)
        ms, nsteps = eng.sync_times()
        eng.profile(False)
        ms /= nsteps
        x = np.concatenate([eng.state_get(r) for r in probe])
        if ref is None:
            ref = x
        print(json.dumps({"kernel": kern, "rpt": rpt, "split": split, "threads": threads, "update_ms": ms, "GBs": upd_bytes / ms / 1e6,
                          "frac": upd_bytes / ms / 1e6 / peak, "m": args.m, "deg": args.deg, "regions": args.regions,
                          "finite": bool(np.isfinite(x).all()), "bit_identical_to_first": bool(np.array_equal(x, ref))}),
              flush=True)


if __name__ == "__main__":
    main()
