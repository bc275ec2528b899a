"""CUDA engine vs the CPU oracle, through the C ABI (ctypes).  All tests here need a B200: -m gpu.

Tolerances: single predict step <= 1e-13 relative (inf-norm, scaled by the vector's
inf-norm); open/closed loop over N steps <= 1e-10; index work exact.
"""
import importlib

import numpy as np
import pytest

from helpers import c_region, initial_grids, oc, region_weights, rel_inf, sst_input_mask, syn

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-13
TOL_LOOP = 1e-10


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


def upload(eng, w, dense=True, **kw):
    if dense and w.get("win") is not None:
        eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win=w["win"],
                          sst_bool_input=w["sst_bool_input"], **kw)
    else:
        eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                          win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"], **kw)


def single_region_engine(E, w, **kw):
    """numprocs = number_of_regions makes rank r own exactly region r"""
    eng = E.Engine(number_of_regions=w["num_regions"], irank=w["region"], numprocs=w["num_regions"],
                   precip_bool=w["precip_bool"], slab_ocean_model_bool=w["sst_bool"], ml_only=w["ml_only"])
    assert eng.region_indices == [w["region"]]
    upload(eng, w, **kw)
    eng.finalize()
    return eng


REGIONS = [0, 23, 555, 556, 557, 24 * 47 + 3, 1151, 24 * 20]


@pytest.mark.parametrize("region", REGIONS)
@pytest.mark.parametrize("m", [600, 2000])
def test_predict_single_step(E, region, m):
    w = region_weights(1152, region, m=m)
    rc = c_region(w)
    eng = single_region_engine(E, w)
    rng = np.random.default_rng(100 + region)
    x0 = 0.3 * rng.standard_normal(w["n"])
    fb = rng.standard_normal(w["D"])
    lm = rng.standard_normal(w["S"])
    rc.x[:] = x0
    rc.feedback[:] = fb
    rc.local_model[:] = lm
    eng.state_set(region, x0)
    eng.feedback_set(region, fb)
    eng.local_model_set(region, lm)
    rc.predict()
    eng.predict()
    assert rel_inf(eng.state_get(region), rc.x) < TOL_STEP
    assert rel_inf(eng.outvec_get(region), rc.outvec) < TOL_STEP
    eng.close()


def test_predict_full_size_region_open_loop(E):
    # config 1: region 555, m=6000 (n=5760, D=576, P=136, S=132, k=33177); sync 55 then 100 open-loop steps
    w = region_weights(1152, 555, m=6000, with_dense_win=True)
    assert (w["n"], w["D"], w["P"], w["S"], w["k"]) == (5760, 576, 136, 132, 33177)
    rc = c_region(w)
    eng = single_region_engine(E, w)
    rng = np.random.default_rng(555)
    T = 55 + 100
    series = syn.ar1_series(w["D"], T, rng)
    model = np.asfortranarray(rng.standard_normal((w["S"], T)))
    rc.synchronize(series[:, :55], 55)
    eng.synchronize(555, series[:, :55])
    assert rel_inf(eng.state_get(555), rc.x) < TOL_LOOP
    worst = 0.0
    for t in range(55, T):
        rc.feedback[:] = series[:, t]
        rc.local_model[:] = model[:, t]
        eng.feedback_set(555, series[:, t])
        eng.local_model_set(555, model[:, t])
        rc.predict()
        eng.predict()
        worst = max(worst, rel_inf(eng.outvec_get(555), rc.outvec))
    assert worst < TOL_LOOP
    assert rel_inf(eng.state_get(555), rc.x) < TOL_LOOP
    eng.close()


def test_synchronize_leakage_and_ld(E):
    w = region_weights(1152, 24 * 5 + 7, m=900)
    rc = c_region(w)
    rc.set_leakage(0.35)
    eng = single_region_engine(E, w, leakage=0.35)
    rng = np.random.default_rng(3)
    padded = np.asfortranarray(rng.standard_normal((w["D"] + 5, 9)))   # ld > D
    rc.synchronize(np.asfortranarray(padded[:w["D"], :]), 9)
    eng.synchronize(w["region"], padded, 9)
    assert rel_inf(eng.state_get(w["region"]), rc.x) < TOL_STEP * 10
    eng.close()


def test_coo_duplicates_sum_and_wide_rows(E):
    # duplicate (row,col) pairs must sum (COO semantics, src/mod_linalg.f90:17); rows may exceed 6 entries
    w = region_weights(1152, 555, m=600)
    rows, cols, vals = w["rows"].copy(), w["cols"].copy(), w["vals"].copy()
    rows[10:20] = rows[0]
    cols[10:20] = cols[0]
    w["rows"], w["cols"], w["vals"] = rows, cols, vals
    rc = c_region(w)
    eng = single_region_engine(E, w)
    rng = np.random.default_rng(4)
    x0 = rng.standard_normal(w["n"])
    fb = rng.standard_normal(w["D"])
    rc.x[:] = x0
    rc.feedback[:] = fb
    eng.state_set(555, x0)
    eng.feedback_set(555, fb)
    rc.predict()
    eng.predict()
    assert rel_inf(eng.state_get(555), rc.x) < TOL_STEP
    eng.close()


def test_dense_win_fallback(E):
    # a W_in that is NOT one-non-zero-per-row must still be honoured (dense path)
    w = region_weights(1152, 555, m=600)
    rng = np.random.default_rng(6)
    win = w["win"].copy(order="F")
    win[::7, 3] += rng.standard_normal(win[::7, 3].shape)
    win[5, :] = rng.standard_normal(w["D"])
    w["win"] = win
    rc = c_region(w)
    eng = single_region_engine(E, w)
    x0 = rng.standard_normal(w["n"]) * 0.2
    fb = rng.standard_normal(w["D"])
    lm = rng.standard_normal(w["S"])
    rc.x[:] = x0
    rc.feedback[:] = fb
    rc.local_model[:] = lm
    eng.state_set(555, x0)
    eng.feedback_set(555, fb)
    eng.local_model_set(555, lm)
    for _ in range(3):
        rc.predict()
        eng.predict()
    assert rel_inf(eng.state_get(555), rc.x) < 1e-12
    assert rel_inf(eng.outvec_get(555), rc.outvec) < 1e-12
    eng.close()


def test_dense_and_compact_upload_agree_bitwise(E):
    w = region_weights(1152, 777, m=600)
    outs = []
    for dense in (True, False):
        eng = E.Engine(number_of_regions=1152, irank=777, numprocs=1152)
        upload(eng, w, dense=dense)
        eng.finalize()
        eng.feedback_set(777, np.linspace(-1, 1, w["D"]))
        eng.local_model_set(777, np.linspace(1, -1, w["S"]))
        eng.predict()
        eng.predict()
        outs.append((eng.state_get(777), eng.outvec_get(777)))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_ml_only_predict(E):
    w = region_weights(1152, 555, m=600, ml_only=True)
    assert w["S"] == 0
    rc = c_region(w)
    eng = single_region_engine(E, w)
    rng = np.random.default_rng(12)
    fb = rng.standard_normal(w["D"])
    rc.feedback[:] = fb
    eng.feedback_set(555, fb)
    for _ in range(2):
        rc.predict()
        eng.predict()
    assert rel_inf(eng.outvec_get(555), rc.outvec) < TOL_STEP * 10
    eng.close()


def test_upload_errors_are_reported_not_fatal(E):
    w = region_weights(1152, 555, m=600)
    eng = E.Engine(number_of_regions=1152, irank=555, numprocs=1152)
    with pytest.raises(E.EngineError):      # predict before finalize
        eng.predict()
    bad = dict(w)
    bad["rows"] = w["rows"].copy()
    bad["rows"][3] = w["n"] + 1              # mklsparse would stop (src/mod_linalg.f90:18-22)
    with pytest.raises(E.EngineError):
        upload(eng, bad)
    other = region_weights(1152, 556, m=600)
    with pytest.raises(E.EngineError):      # region not owned by this rank
        upload(eng, other)
    wrong = dict(w)
    wrong["mean"], wrong["std"] = w["mean"][:-1], w["std"][:-1]
    with pytest.raises(E.EngineError):      # L does not match the tiling
        upload(eng, wrong)
    upload(eng, w)
    eng.finalize()
    eng.predict()
    eng.close()


@pytest.fixture(scope="module")
def small_model(E):
    """full 1152-region tiling with minimal reservoirs (n = D): engine + oracle twins"""
    ws = [region_weights(1152, r, m=300, with_dense_win=False) for r in range(1152)]
    eng = E.Engine(number_of_regions=1152, sst_prescribed=True)
    for w in ws:
        upload(eng, w, dense=False)
    eng.finalize()
    rcs = [c_region(w) for w in ws]
    return ws, eng, rcs


def test_closed_hybrid_loop_full_tiling(E, small_model):
    ws, eng, rcs = small_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    sst_mean = np.array([w["mean"][-1] for w in ws])
    sst_std = np.array([w["std"][-1] for w in ws])
    rng = np.random.default_rng(21)
    for w, rc in zip(ws, rcs):
        fb = rng.standard_normal(w["D"])
        lm = rng.standard_normal(w["S"])
        rc.feedback[:] = fb
        rc.local_model[:] = lm
        rc.x[:] = 0.0
        eng.feedback_set(w["region"], fb)
        eng.local_model_set(w["region"], lm)
        eng.state_set(w["region"], np.zeros(w["n"]))
    nthreads = 8
    for step in range(1, 6):
        oc.predict_all(rcs, nthreads=nthreads)
        eng.predict()
        # prescribed-SST mode: every region "has an ocean reservoir" whose output is the prescribed field
        has = np.ones(len(rcs), dtype=np.int32)
        oo = np.zeros((len(rcs), 4))
        for i, w in enumerate(ws):
            xs, xe, ys, ye, *_ = oc.getxyresextent(1152, w["region"])
            oo[i] = G["base_sst"][xs - 1:xe, ys - 1:ye].ravel(order="F")
        gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
        ge = eng.step_exchange_begin(step)
        tol = TOL_STEP * 10 if step == 1 else TOL_LOOP
        for a, b in zip(ge, gc):
            assert rel_inf(a, b) < tol
        assert ge[0][3].min() >= 0.000001 and ge[3].min() >= 272.0
        f4c, f2c = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
        f4e, f2e = oc.host_stub(ge[0], ge[1], G["clim4d"], G["clim2d"])
        oc.step_scatter(rcs, True, True, False, *gc, f4c, f2c, G["tisr"], sst_mean, sst_std, nthreads=nthreads)
        eng.step_exchange_end(step, f4e, f2e, G["tisr"])
        for r in (0, 23, 555, 556, 1151, 1128, 700):
            assert rel_inf(eng.feedback_get(r), rcs[r].feedback) < tol
            assert rel_inf(eng.local_model_get(r), rcs[r].local_model) < tol
            assert rel_inf(eng.outvec_get(r), rcs[r].outvec) < tol


def test_exchange_is_exact_given_identical_outvecs(E, small_model):
    # scatter -> clamp -> gather -> standardise is pure index work + two roundings per element: feed the
    # engine's own outvecs to the oracle and require BIT equality of grids, feedback and local_model.
    ws, eng, rcs = small_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    eng.predict()
    for w, rc in zip(ws, rcs):
        rc.outvec[:] = eng.outvec_get(w["region"])
    has = np.ones(len(rcs), dtype=np.int32)
    oo = np.zeros((len(rcs), 4))
    for i, w in enumerate(ws):
        xs, xe, ys, ye, *_ = oc.getxyresextent(1152, w["region"])
        oo[i] = G["base_sst"][xs - 1:xe, ys - 1:ye].ravel(order="F")
    gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
    ge = eng.step_exchange_begin(1)
    for a, b in zip(ge, gc):
        assert np.array_equal(a, b)
    f4, f2 = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
    sst_mean = np.array([w["mean"][-1] for w in ws])
    sst_std = np.array([w["std"][-1] for w in ws])
    oc.step_scatter(rcs, True, True, False, *gc, f4, f2, G["tisr"], sst_mean, sst_std, nthreads=8)
    eng.step_exchange_end(1, f4, f2, G["tisr"])
    for w, rc in zip(ws, rcs):
        assert np.array_equal(eng.feedback_get(w["region"]), rc.feedback)
        assert np.array_equal(eng.local_model_get(w["region"]), rc.local_model)


def test_readout_properties_full_tiling(E, small_model):
    """size-independent properties: W_out = 0 -> outvec is the mean vector; readout is linear in W_out"""
    ws, eng, rcs = small_model
    r = 600
    w = ws[r]
    saved = eng.wout_get(r)
    assert np.array_equal(saved, w["wout"])
    x = eng.state_get(r)
    fb = eng.feedback_get(r)
    lm = eng.local_model_get(r)

    def run(wout):
        eng.wout_set(r, wout)
        eng.state_set(r, x)
        eng.feedback_set(r, fb)
        eng.local_model_set(r, lm)
        eng.predict()
        return eng.outvec_get(r)

    m = E.region_maps(1152, r, 1, True, w["sst_bool_input"])
    mean_vec = w["mean"][m["output_ms"]]
    std_vec = w["std"][m["output_ms"]]
    o0 = run(np.zeros_like(saved))
    assert np.array_equal(o0, mean_vec)
    o1 = run(saved)
    o2 = run(2.0 * saved)
    # (o - mean)/std is linear in W_out
    assert rel_inf((o2 - mean_vec) / std_vec, 2.0 * (o1 - mean_vec) / std_vec) < 1e-12
    eng.wout_set(r, saved)


def _reset_small_model(ws, eng, rcs, seed):
    rng = np.random.default_rng(seed)
    for w, rc in zip(ws, rcs):
        fb = rng.standard_normal(w["D"])
        lm = rng.standard_normal(w["S"])
        x0 = 0.1 * rng.standard_normal(w["n"])
        rc.feedback[:], rc.local_model[:], rc.x[:] = fb, lm, x0
        eng.feedback_set(w["region"], fb)
        eng.local_model_set(w["region"], lm)
        eng.state_set(w["region"], x0)


def test_overlapped_step_matches_sequential_and_oracle(E, small_model):
    """the next predict's state update + x~ readout runs while the host model works; only the
    summation order of the readout changes (v_p + v_ml).  Grids must agree with the sequential engine loop to
    1e-12 and with the oracle to the loop tolerance; feedback vectors bit-identical given identical grids."""
    ws, eng, rcs = small_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    nsteps = 6

    def tisr_of(t):
        return np.asfortranarray(G["tisr"] * (1.0 + 0.01 * t))

    def run(overlap):
        _reset_small_model(ws, eng, rcs, 33)
        eng.set_overlap(overlap)
        grids = []
        for t in range(1, nsteps + 1):
            eng.predict()
            if overlap:
                eng.set_tisr(tisr_of(t))
            g = eng.step_exchange_begin(t)
            f4, f2 = oc.host_stub(g[0], g[1], G["clim4d"], G["clim2d"])
            eng.step_exchange_end(t, f4, f2, None if overlap else tisr_of(t))
            grids.append(g)
        fb = {r: eng.feedback_get(r) for r in (0, 555, 1151)}
        lm = {r: eng.local_model_get(r) for r in (0, 555, 1151)}
        eng.set_overlap(False)
        return grids, fb, lm

    seq, fb_s, lm_s = run(False)
    ovl, fb_o, lm_o = run(True)
    for t in range(nsteps):
        for a, b in zip(ovl[t], seq[t]):
            assert rel_inf(a, b) < (1e-14 if t == 0 else 1e-11)
    for r in fb_s:
        assert rel_inf(fb_o[r], fb_s[r]) < 1e-11 and rel_inf(lm_o[r], lm_s[r]) < 1e-11
    # oracle closed loop from the same start
    _reset_small_model(ws, eng, rcs, 33)
    sst_mean = np.array([w["mean"][-1] for w in ws])
    sst_std = np.array([w["std"][-1] for w in ws])
    has = np.ones(len(rcs), dtype=np.int32)
    oo = np.zeros((len(rcs), 4))
    for i, w in enumerate(ws):
        xs, xe, ys, ye, *_ = oc.getxyresextent(1152, w["region"])
        oo[i] = G["base_sst"][xs - 1:xe, ys - 1:ye].ravel(order="F")
    for t in range(1, nsteps + 1):
        oc.predict_all(rcs, nthreads=8)
        gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
        for a, b in zip(ovl[t - 1], gc):
            assert rel_inf(a, b) < TOL_LOOP
        f4, f2 = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
        oc.step_scatter(rcs, True, True, False, *gc, f4, f2, tisr_of(t), sst_mean, sst_std, nthreads=8)


def test_overlapped_step_protocol_errors(E, small_model):
    ws, eng, rcs = small_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    eng.set_overlap(True)
    eng.predict()
    with pytest.raises(E.EngineError):          # TISR must be on the device before the exchange begins
        eng.step_exchange_begin(1)
    eng.set_tisr(G["tisr"])
    g = eng.step_exchange_begin(1)
    with pytest.raises(E.EngineError):          # predict before the step was closed
        eng.predict()
    f4, f2 = oc.host_stub(g[0], g[1], G["clim4d"], G["clim2d"])
    eng.step_exchange_end(1, f4, f2)
    eng.predict()                               # consumes the look-ahead result
    eng.set_overlap(False)


def test_outvec_component_contribs(E):
    """reservoir%v_p / reservoir%v_ml (src/mod_reservoir.f90:1458-1461): the two halves of the readout"""
    region = 556
    w = region_weights(1152, region, m=1300)
    rc = c_region(w)
    eng = single_region_engine(E, w)
    rng = np.random.default_rng(8)
    x0, fb, lm = 0.3 * rng.standard_normal(w["n"]), rng.standard_normal(w["D"]), rng.standard_normal(w["S"])
    rc.x[:], rc.feedback[:], rc.local_model[:] = x0, fb, lm
    eng.state_set(region, x0)
    eng.feedback_set(region, fb)
    eng.local_model_set(region, lm)
    with pytest.raises(E.EngineError):
        eng.contribs_get(region)
    eng.set_contribs(True)
    rc.predict()
    eng.predict()
    xt = rc.x.copy()
    xt[1::2] **= 2                                    # x_temp(2:n:2) squared
    S = w["S"]
    v_p = w["wout"][:, :S] @ lm
    v_ml = w["wout"][:, S:] @ xt
    vp_e, vml_e = eng.contribs_get(region)
    assert rel_inf(vp_e, v_p) < 1e-13 and rel_inf(vml_e, v_ml) < 1e-13
    assert rel_inf(eng.outvec_get(region), rc.outvec) < TOL_STEP * 10     # split order vs the oracle's fused order
    eng.set_contribs(False)
    eng.close()


@pytest.mark.parametrize("R,region", [(288, 145), (288, 0), (576, 300)])
def test_other_tilings_predict_and_exchange(E, R, region):
    """288 regions -> 4x4 tiles (P = 544, the tiling of the reference's unit test), 576 -> 4x2: the engine takes its
    sizes from the tiling, nothing is specialised to 1152"""
    m = 2600 if R == 288 else 1500
    w = region_weights(R, region, m=m, with_dense_win=False)
    assert w["n"] > 0
    rc = c_region(w)
    eng = E.Engine(number_of_regions=R, irank=region, numprocs=R, sst_prescribed=True)
    eng.region_upload(region, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                      win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    rng = np.random.default_rng(R + region)
    x0, fb, lm = 0.3 * rng.standard_normal(w["n"]), rng.standard_normal(w["D"]), rng.standard_normal(w["S"])
    rc.x[:], rc.feedback[:], rc.local_model[:] = x0, fb, lm
    eng.state_set(region, x0)
    eng.feedback_set(region, fb)
    eng.local_model_set(region, lm)
    rc.predict()
    eng.predict()
    assert rel_inf(eng.state_get(region), rc.x) < TOL_STEP
    assert rel_inf(eng.outvec_get(region), rc.outvec) < TOL_STEP * 10
    # overlapped (split-order) finish with this P as well
    eng.state_set(region, x0)
    eng.set_contribs(True)
    eng.predict()
    assert rel_inf(eng.outvec_get(region), rc.outvec) < TOL_STEP * 10
    eng.close()


def test_zero_copy_staging_views(E, small_model):
    """sml_step_exchange_begin_view / sml_forecast_staging: same grids and same feedback as the copying entry points"""
    ws, eng, rcs = small_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    _reset_small_model(ws, eng, rcs, 44)
    eng.predict()
    copies = eng.step_exchange_begin(1)
    views = eng.step_exchange_begin_view(1)
    for a, b in zip(views, copies):
        assert np.array_equal(a, b) and not a.flags.writeable
    f4, f2 = oc.host_stub(copies[0], copies[1], G["clim4d"], G["clim2d"])
    eng.step_exchange_end(1, f4, f2, G["tisr"])
    fb_copy = {r: (eng.feedback_get(r), eng.local_model_get(r)) for r in (0, 555, 1151)}
    s4, s2, st = eng.forecast_staging()
    s4[...] = f4
    s2[...] = f2
    st[...] = G["tisr"]
    for r in fb_copy:                                   # scramble, then rebuild from the staging written in place
        eng.feedback_set(r, np.zeros(ws[r]["D"]))
        eng.local_model_set(r, np.zeros(ws[r]["S"]))
    eng.step_exchange_end(1, s4, s2, st)
    for r, (fb, lm) in fb_copy.items():
        assert np.array_equal(eng.feedback_get(r), fb) and np.array_equal(eng.local_model_get(r), lm)


def test_fused_exchange_kernel_equals_pack_then_unpack(E, small_model):
    """sml_step_exchange_device (one cooperative launch: scatter, grid barrier, gather) against the two separate kernels:
    grids, feedback and local_model bit for bit"""
    ws, eng, rcs = small_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    _reset_small_model(ws, eng, rcs, 77)
    eng.predict()
    probe = (0, 23, 555, 700, 1128, 1151)
    eng.step_pack_device(1)
    eng.step_unpack_device(1)
    ref_grids = eng.grids_get()
    ref = {r: (eng.feedback_get(r), eng.local_model_get(r)) for r in probe}
    for r in probe:                                   # scramble what the exchange rebuilds
        eng.feedback_set(r, np.zeros(ws[r]["D"]))
        eng.local_model_set(r, np.zeros(ws[r]["S"]))
    for _ in range(3):                                # the arrival counter of the grid barrier is monotonic: several calls
        eng.step_exchange_device(1)
    for a, b in zip(eng.grids_get(), ref_grids):
        assert np.array_equal(a, b)
    for r, (fb, lm) in ref.items():
        assert np.array_equal(eng.feedback_get(r), fb) and np.array_equal(eng.local_model_get(r), lm)
