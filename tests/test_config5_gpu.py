"""BASELINE config 5 corners against the CPU oracle (-m gpu): reservoir sizes m = 2000 / 8000 / 12000 and adjacency
degrees 3 / 12 / 24, where the kernels' launch geometry changes (k_update_sx drops to one CTA per SM or falls back to
k_update at m = 12000, the step kernels' x~ tile and stage sizing move with n, wide ELL rows at degree 24).

For every corner: the update-only path (synchronize), one fused step, a 20-step open loop -- each with BOTH fused step
kernels (k_step_persist, the default, and the classic k_step) -- and, at m = 12000, the Gram accumulation + ridge solve.
Tolerances as everywhere: one step <= 1e-13, loops <= 1e-10, Gram <= 1e-12, solve residual <= 1e-13.
"""
import importlib
import os

import numpy as np
import pytest

from helpers import c_region, region_weights, rel_inf, syn

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-13
TOL_LOOP = 1e-10

CORNERS = [(2000, 3.0), (2000, 24.0), (6000, 3.0), (6000, 12.0), (6000, 24.0), (8000, 6.0), (8000, 12.0), (12000, 3.0),
           (12000, 6.0), (12000, 24.0)]


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


class step_kernel:
    """SML_STEP_KERNEL is read when the step plan is built (sml_finalize)"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.old = os.environ.get("SML_STEP_KERNEL")
        os.environ["SML_STEP_KERNEL"] = self.name

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("SML_STEP_KERNEL", None)
        else:
            os.environ["SML_STEP_KERNEL"] = self.old


def make_engine(E, w, kernel):
    with step_kernel(kernel):
        eng = E.Engine(number_of_regions=w["num_regions"], irank=w["region"], numprocs=w["num_regions"])
        eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                          win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
        eng.finalize()
    assert eng.step_plan()["kernel"] == ("k_step_persist" if kernel == "persist" else "k_step")
    return eng


@pytest.mark.parametrize("m,deg", CORNERS)
def test_corner_update_step_and_open_loop(E, m, deg):
    region = 555
    w = region_weights(1152, region, m=m, deg=deg, with_dense_win=False)
    assert abs(w["k"] - (deg / m) * w["n"] * w["n"]) < 2.0    # k = int(density * n * n), density = deg / m (:172)
    rng = np.random.default_rng(int(m + deg))
    T_SYNC, T_LOOP = 7, 20
    series = syn.ar1_series(w["D"], T_SYNC + T_LOOP, rng)
    model = np.asfortranarray(rng.standard_normal((w["S"], T_SYNC + T_LOOP)))
    x0 = 0.2 * rng.standard_normal(w["n"])

    # oracle trajectory
    rc = c_region(w)
    rc.x[:] = x0
    rc.synchronize(np.asfortranarray(series[:, :T_SYNC]), T_SYNC)
    x_sync = rc.x.copy()
    outs, states = [], []
    for t in range(T_SYNC, T_SYNC + T_LOOP):
        rc.feedback[:] = series[:, t]
        rc.local_model[:] = model[:, t]
        rc.predict()
        outs.append(rc.outvec.copy())
        states.append(rc.x.copy())

    results = {}
    for kernel in ("persist", "classic"):
        eng = make_engine(E, w, kernel)
        eng.state_set(region, x0)
        eng.synchronize(region, np.asfortranarray(series[:, :T_SYNC]))      # update-only kernels
        assert rel_inf(eng.state_get(region), x_sync) < TOL_STEP * 10, (kernel, "synchronize")
        worst = 0.0
        got = []
        for i, t in enumerate(range(T_SYNC, T_SYNC + T_LOOP)):
            eng.feedback_set(region, series[:, t])
            eng.local_model_set(region, model[:, t])
            eng.predict()
            ov = eng.outvec_get(region)
            got.append(ov)
            err = rel_inf(ov, outs[i])
            if i == 0:
                assert err < TOL_STEP, (kernel, "first fused step", err)
                assert rel_inf(eng.state_get(region), states[0]) < TOL_STEP, (kernel, "state after one step")
            worst = max(worst, err)
        assert worst < TOL_LOOP, (kernel, "open loop", worst)
        assert rel_inf(eng.state_get(region), states[-1]) < TOL_LOOP
        results[kernel] = got
        eng.close()
    # the two kernels order the readout sum differently: equal within rounding, and each within tolerance of the oracle
    for a, b in zip(results["persist"], results["classic"]):
        assert rel_inf(a, b) < 1e-12


def test_corner_m12000_all_regions_synchronize(E):
    """the batched update-only path at m = 12000 (shared-memory staging no longer fits two CTAs per SM)"""
    regions = [0, 23, 555, 1151]
    ws = {r: region_weights(1152, r, m=12000, with_dense_win=False) for r in regions}
    rng = np.random.default_rng(12)
    T = 6
    for r in regions:
        w = ws[r]
        eng = E.Engine(number_of_regions=1152, irank=r, numprocs=1152)
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
        eng.finalize()
        series = syn.ar1_series(w["D"], T, rng)
        rc = c_region(w)
        rc.synchronize(series, T)
        eng.synchronize_all([series], T)
        assert rel_inf(eng.state_get(r), rc.x) < TOL_STEP * 10
        eng.close()


def test_corner_m12000_gram_and_solve(E):
    """training at the largest reservoir: Gram <= 1e-12 of the oracle's accumulators, ridge solve residual <= 1e-13"""
    region, m = 555, 12000
    w = region_weights(1152, region, m=m, with_dense_win=False)
    N, P = w["n"] + w["S"], w["P"]
    rng = np.random.default_rng(77)
    cols, discard, batch = 140, 40, 50
    td = syn.ar1_series(w["D"], cols, rng)
    im = np.asfortranarray(rng.standard_normal((w["S"], cols)))

    rc = c_region(w)
    rc.train_init(batch)
    rc.train_phase(td, im, discard)
    sxs_ref, sxt_ref = rc.sxs.copy(), rc.sxt.copy()

    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    eng.region_upload(region, w["rows"], w["cols"], w["vals"], None, w["mean"], w["std"], win_compact=w["winc"],
                      win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"], S=w["S"], P=w["P"])
    eng.finalize()
    eng.train_begin([region], batch)
    eng.train_feed([td], [im], discard)
    sxs, sxt = eng.train_gram_get(region)
    scale = np.max(np.abs(sxs_ref))
    assert np.max(np.abs(sxs - sxs_ref)) / scale < 1e-12
    assert np.max(np.abs(sxt - sxt_ref)) / max(np.max(np.abs(sxt_ref)), 1e-300) < 1e-12
    beta_res, beta_model = 1e-3, 1.0
    info = eng.train_solve(beta_res, beta_model, True, 0.0)
    assert int(info[0]) == 0
    wout = eng.wout_get(region)
    eng.train_end()
    eng.close()
    # residual of A^T X = B^T with A = sxs + ridge (beta^2 with using_prior), B = sxt  (src/mod_reservoir.f90:1275-1316);
    # Frobenius norms (a 2-norm of an 11652^2 matrix would take minutes)
    A = sxs_ref.copy()
    idx = np.arange(N)
    A[idx[:w["S"]], idx[:w["S"]]] += beta_model ** 2
    A[idx[w["S"]:], idx[w["S"]:]] += beta_res ** 2
    X = wout.T
    res = A.T @ X - sxt_ref.T
    assert np.linalg.norm(res) / (np.linalg.norm(A) * np.linalg.norm(X) + np.linalg.norm(sxt_ref)) < 1e-13
    assert P == 136 and N == w["n"] + 132
