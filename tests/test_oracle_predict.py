"""C oracle vs NumPy oracle: synchronize, predict, and the closed hybrid loop (small reservoirs)."""
import numpy as np
import pytest

from helpers import c_region, initial_grids, np_region, oc, on, region_weights, rel_inf, syn

REGIONS = [0, 23, 555, 556, 24 * 47 + 3, 1151, 24 * 20]  # polar+periodic, polar, interior(sst/land), east edge ...


@pytest.mark.parametrize("region", REGIONS)
def test_sync_and_predict(region):
    w = region_weights(1152, region, m=1100)
    rc, rn = c_region(w), np_region(w)
    rng = np.random.default_rng(11 + region)
    T = 12
    inputs = syn.ar1_series(w["D"], T, rng)
    rc.synchronize(inputs, T)
    rn.x = on.synchronize(rn, inputs, np.zeros(rn.n), T)
    assert rel_inf(rc.x, rn.x) < 1e-13
    fb = rng.standard_normal(w["D"])
    lm = rng.standard_normal(w["S"])
    rc.feedback[:] = fb
    rc.local_model[:] = lm
    rn.feedback, rn.local_model = fb.copy(), lm.copy()
    for _ in range(3):
        rc.predict()
        rn.x, rn.outvec = on.predict(rn, rn.x)
    assert rel_inf(rc.x, rn.x) < 1e-13
    assert rel_inf(rc.outvec, rn.outvec) < 1e-12


def test_leakage_general_form():
    w = region_weights(1152, 555, m=600)
    rc, rn = c_region(w), np_region(w)
    rc.set_leakage(0.3)
    rn.leakage = 0.3
    rng = np.random.default_rng(5)
    inputs = syn.ar1_series(w["D"], 6, rng)
    rc.synchronize(inputs, 6)
    rn.x = on.synchronize(rn, inputs, np.zeros(rn.n), 6)
    assert rel_inf(rc.x, rn.x) < 1e-13


def test_squared_feature_is_even_one_based_in_place():
    # src/mod_reservoir.f90:1450-1451: x_temp(2:n:2) squared; feature order [local_model ; x_temp]
    w = region_weights(1152, 555, m=600)
    w["wout"] = np.zeros_like(w["wout"])
    w["wout"][0, w["S"] + 0] = 1.0   # picks x(1) (odd, 1-based) -> linear
    w["wout"][1, w["S"] + 1] = 1.0   # picks x(2) (even) -> squared
    w["wout"][2, 0] = 1.0            # picks local_model(1)
    w["mean"][:] = 0.0
    w["std"][:] = 1.0
    rc = c_region(w)
    rc.feedback[:] = 0.3
    rc.local_model[:] = 0.25
    rc.predict()
    x = rc.x.copy()
    assert rc.outvec[0] == x[0]
    assert rc.outvec[1] == x[1] * x[1]
    assert rc.outvec[2] == 0.25


def test_closed_hybrid_loop_c_vs_numpy():
    regions = REGIONS
    ws = [region_weights(1152, r, m=600) for r in regions]
    rcs = [c_region(w) for w in ws]
    rns = [np_region(w) for w in ws]
    G = initial_grids()
    sst_mean = np.array([w["mean"][-1] for w in ws])
    sst_std = np.array([w["std"][-1] for w in ws])
    rng = np.random.default_rng(2)
    for w, rc, rn in zip(ws, rcs, rns):
        fb = rng.standard_normal(w["D"])
        lm = rng.standard_normal(w["S"])
        x0 = 0.1 * rng.standard_normal(w["n"])
        rc.feedback[:] = fb
        rc.local_model[:] = lm
        rc.x[:] = x0
        rn.feedback, rn.local_model, rn.x = fb.copy(), lm.copy(), x0.copy()
    for step in range(4):
        oc.predict_all(rcs, nthreads=2)
        for rn in rns:
            rn.x, rn.outvec = on.predict(rn, rn.x)
        gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"])
        gn = on.step_gather(rns, True, True, G["base_sst"], G["sea_mask"])
        for a, b in zip(gc, gn):
            assert rel_inf(a, b) < 1e-11
        f4c, f2c = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
        f4n, f2n = on.host_stub(gn[0], gn[1], G["clim4d"], G["clim2d"])
        oc.step_scatter(rcs, True, True, False, *gc, f4c, f2c, G["tisr"], sst_mean, sst_std, nthreads=2)
        on.step_scatter(rns, True, True, *gn, f4n, f2n, G["tisr"], sst_mean, sst_std)
        for rc, rn in zip(rcs, rns):
            assert rel_inf(rc.feedback, rn.feedback) < 1e-10
            assert rel_inf(rc.local_model, rn.local_model) < 1e-10


def test_gather_clamps():
    # src/mpires.f90:460-490: q floor 1e-6, precip < 1e-5 -> 0, sst < 272 -> 272, land mask -> base sst,
    # regions without an ocean reservoir report 272.0 (:323-326)
    ws = [region_weights(1152, r, m=600) for r in (100, 101)]
    rcs = [c_region(w) for w in ws]
    for rc in rcs:
        rc.outvec[:] = -1.0
    G = initial_grids()
    w4d, w2d, wp, wsst = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"])
    assert w4d[3].min() == 0.000001 and w4d[0].min() == -1.0
    assert wp.min() == 0.0 and wp.max() == 0.0
    xs, xe, ys, ye, *_ = oc.getxyresextent(1152, 100)
    tile = wsst[xs - 1:xe, ys - 1:ye]
    mask = G["sea_mask"][xs - 1:xe, ys - 1:ye] > 0
    assert np.all(tile[~mask] == 272.0)
    assert np.all(tile[mask] == np.maximum(G["base_sst"][xs - 1:xe, ys - 1:ye][mask], 272.0))
    assert wsst.min() >= 272.0


def test_compact_win_is_bit_identical_to_dense():
    # matmul(win,u)(j) == win(j,c_j)*u(c_j) exactly when every other entry of the row is zero
    w = region_weights(1152, 555, m=1100)
    rd = c_region(w)
    w2 = dict(w)
    w2["win"] = None
    rk = c_region(w2)
    rng = np.random.default_rng(8)
    inputs = syn.ar1_series(w["D"], 9, rng)
    rd.synchronize(inputs, 9)
    rk.synchronize(inputs, 9)
    assert np.array_equal(rd.x, rk.x)
    for r in (rd, rk):
        r.feedback[:] = inputs[:, 3]
        r.local_model[:] = 0.5
        r.predict()
    assert np.array_equal(rd.outvec, rk.outvec)
