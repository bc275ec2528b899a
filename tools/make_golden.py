#!/usr/bin/env python
"""Writes tests/golden/hotpath_v1.npz: regression vectors of the hot path.

PROVENANCE -- read before trusting: the reference (Fortran 90 + MPI + MKL + ARPACK + NetCDF) cannot be built
or imported here, so these vectors are NOT outputs of the reference.  They are outputs of the CPU oracle
(oracle/speedyml_oracle.c, cross-checked against oracle/oracle_np.py) on seeded inputs, frozen so that (a) an
accidental change of the oracle shows up as a diff against history and (b) the CUDA engine is also compared
with a committed artefact, not only with a checker built in the same run.  The only reference-held known
answers (tests/mod_unit_test.f90:16-47, :63-96) are stored alongside under ref_* keys.

    python tools/make_golden.py        # regenerates the file; commit the result
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import c_ocean, c_region, ocean_weights, oc, region_weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "hotpath_v1.npz")
PREDICT_REGIONS = (0, 555, 24 * 47 + 23)
M = 300


def predict_case(region):
    """2 synchronize columns + 3 predict steps of one small region; inputs come from the seeded generator"""
    w = region_weights(1152, region, m=M)
    rc = c_region(w)
    rng = np.random.default_rng(9000 + region)
    series = np.asfortranarray(rng.standard_normal((w["D"], 5)))
    model = np.asfortranarray(rng.standard_normal((w["S"], 5)))
    rc.synchronize(series[:, :2], 2)
    outs = []
    for t in range(2, 5):
        rc.feedback[:] = series[:, t]
        rc.local_model[:] = model[:, t]
        rc.predict()
        outs.append(rc.outvec.copy())
    return rc.x.copy(), np.array(outs)


def ocean_case(region):
    wa = region_weights(1152, region, m=M, sst_bool_input=True)
    wo = ocean_weights(1152, region, m=M, mean=wa["mean"], std=wa["std"])
    co = c_ocean(wo)
    rng = np.random.default_rng(9500 + region)
    outs = []
    for _ in range(3):
        co.feedback[:] = rng.standard_normal(wo["D"])
        co.predict()
        outs.append(co.outvec.copy())
    return co.x.copy(), np.array(outs)


def main():
    d = {}
    # reference-held known answers
    d["ref_unit_test_288_region145_x"] = np.array([49, 52])          # tests/mod_unit_test.f90:75-87
    d["ref_unit_test_288_tile"] = np.array([4, 4])
    d["ref_pinv_diag_1_10"] = np.diag(1.0 / np.arange(1, 11))        # tests/mod_unit_test.f90:31-44
    # index arithmetic for every region of three tilings
    for R in (1152, 288, 4608):
        ext = np.array([oc.getxyresextent(R, r) for r in range(R)], dtype=np.int32)
        ov = np.array([[int(v) for v in oc.getoverlapindices(R, r, 1)] for r in range(R)], dtype=np.int32)
        td = np.array([oc.get_trainingdataindices(R, r, 1) for r in range(R)], dtype=np.int32)
        d[f"extent_{R}"], d[f"overlap_{R}"], d[f"tdata_{R}"] = ext, ov, td
    for world in (3, 5, 8):
        for rank in (0, 1, world - 1):
            d[f"procdecomp_{world}_{rank}"] = np.array(oc.processor_decomposition(rank, world, 1152), dtype=np.int32)
    for r in PREDICT_REGIONS:
        x, outs = predict_case(r)
        d[f"predict_x_{r}"], d[f"predict_out_{r}"] = x, outs
    x, outs = ocean_case(555)
    d["ocean_x_555"], d["ocean_out_555"] = x, outs
    np.savez_compressed(OUT, **d)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
