"""The compiled C++ host program (speedy-ml_b200/drivers/replay_main.cpp on host/speedyml_host.hpp) drives the engine
through the C ABI without Python: load -> mklsparse -> synchronize -> [predict; sendrecievegrid] x N.  Its grids must
match the CPU oracle's closed loop (1e-10) and its overlapped mode the sequential one (1e-11).  -m gpu."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from case_io import read_output, write_case
from helpers import c_region, initial_grids, oc, region_weights, rel_inf, syn

pytestmark = pytest.mark.gpu

NSTEPS, SYNC = 4, 3


def test_cpp_host_replay_matches_oracle(tmp_path):
    B = importlib.import_module("speedy-ml_b200.build")
    exe = B.build_host_driver()
    ws = [region_weights(1152, r, m=300, with_dense_win=False) for r in range(1152)]
    G = initial_grids()
    rng = np.random.default_rng(2024)
    x0 = [0.1 * rng.standard_normal(w["n"]) for w in ws]
    fb0 = [rng.standard_normal(w["D"]) for w in ws]
    lm0 = [rng.standard_normal(w["S"]) for w in ws]
    sync = [syn.ar1_series(w["D"], SYNC, rng) for w in ws]
    case, out_seq, out_ovl = str(tmp_path / "case.bin"), str(tmp_path / "seq.bin"), str(tmp_path / "ovl.bin")
    write_case(case, ws, x0, fb0, lm0, sync, G, NSTEPS)
    for out, extra in ((out_seq, []), (out_ovl, ["--overlap"])):
        p = subprocess.run([exe, case, out] + extra, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and "program finished correctly" in p.stdout, p.stdout + p.stderr
    seq, ov_seq, fb_seq = read_output(out_seq, ws, NSTEPS)
    ovl, ov_ovl, fb_ovl = read_output(out_ovl, ws, NSTEPS)

    # oracle: the same start, synchronize, closed loop
    rcs = [c_region(w) for w in ws]
    for i, rc in enumerate(rcs):
        rc.x[:] = x0[i]
        rc.synchronize(sync[i], SYNC)
        rc.feedback[:] = fb0[i]
        rc.local_model[:] = lm0[i]
    sst_mean = np.array([w["mean"][-1] for w in ws])
    sst_std = np.array([w["std"][-1] for w in ws])
    has = np.ones(len(rcs), dtype=np.int32)
    oo = np.zeros((len(rcs), 4))
    for i, w in enumerate(ws):
        xs, xe, ys, ye, *_ = oc.getxyresextent(1152, w["region"])
        oo[i] = G["base_sst"][xs - 1:xe, ys - 1:ye].ravel(order="F")
    for t in range(NSTEPS):
        oc.predict_all(rcs, nthreads=8)
        gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
        for a, b in zip(seq[t], gc):
            assert rel_inf(a, b) < 1e-10
        for a, b in zip(ovl[t], seq[t]):
            assert rel_inf(a, b) < 1e-11
        f4, f2 = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
        oc.run_model_clamp(f4)
        oc.step_scatter(rcs, True, True, False, *gc, f4, f2, G["tisr"], sst_mean, sst_std, nthreads=8)
    for i in (0, 23, 555, 1128, 1151):
        assert rel_inf(ov_seq[i], rcs[i].outvec) < 1e-10
        assert rel_inf(fb_seq[i], rcs[i].feedback) < 1e-10
