"""Slab-ocean reservoirs (BASELINE config 4) on the CUDA engine vs the CPU oracle, through the
C ABI.  Needs a B200: -m gpu.  Tolerances as in test_engine_gpu.py; the exchange/feedback assembly is exact."""
import importlib

import numpy as np
import pytest

from helpers import (c_ocean, c_region, initial_grids, oc, ocean_weights, region_weights, rel_inf, sst_input_mask, syn)

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-13
TOL_LOOP = 1e-10


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


def upload_atmo(eng, w):
    eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                      win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])


def upload_ocean(eng, E, w):
    eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                      win_compact=w["winc"], win_col=w["wcol"], D=w["D"], kind=E.OCEAN,
                      sst_mean=w["mean"][w["sst_idx"] - 1], sst_std=w["std"][w["sst_idx"] - 1])


def test_ocean_dims_and_maps_match_oracle(E):
    for reg in (0, 23, 555, 24 * 47 + 23, 1151):
        d = E.ocean_region_dims(1152, reg)
        c = oc.OceanRegion(1152, reg)
        assert (d["n"], d["k"], d["D"], d["P"], d["A"]) == (c.n, c.k, c.D, c.P, c.A)
        m = E.ocean_region_maps(1152, reg)
        ca = oc.Region(1152, reg)
        assert m["atmo_slice0"] == ca.g.atmo3d_end - 4 * (c.D // 8)
        # target rows: push row indices through the oracle tiler
        sv = np.asfortranarray(np.arange(c.D, dtype=np.float64).reshape(-1, 1))
        assert np.array_equal(c.target(sv).ravel().astype(np.int32), m["target_map"])
        # SST tile offsets: tile a grid that stores its own offset
        lay = E.global_layout()
        grid = np.asfortranarray((lay["sst"] + np.arange(96 * 48, dtype=np.float64)).reshape((96, 48), order="F"))
        tile = oc.tileoverlapgrid2d(grid, 1152, reg, 1).ravel(order="F")
        assert np.array_equal(tile.astype(np.int32), m["sst_map"])


@pytest.mark.parametrize("region,m", [(555, 500), (0, 500), (24 * 47 + 23, 900), (555, 4000)])
def test_predict_slab_ml_single_region(E, region, m):
    wa = region_weights(1152, region, m=300, sst_bool_input=True, with_dense_win=False)
    wo = ocean_weights(1152, region, m=m, mean=wa["mean"], std=wa["std"], with_dense_win=False)
    if m == 4000 and region == 555:
        assert (wo["n"], wo["D"], wo["P"]) == (3968, 128, 8)
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    upload_atmo(eng, wa)
    upload_ocean(eng, E, wo)
    eng.finalize()
    co = c_ocean(wo)
    rng = np.random.default_rng(region + m)
    x0 = 0.3 * rng.standard_normal(wo["n"])
    co.x[:] = x0
    eng.state_set(region, x0, kind=E.OCEAN)
    worst = 0.0
    for step in range(4):
        fb = rng.standard_normal(wo["D"])
        co.feedback[:] = fb
        eng.feedback_set(region, fb, kind=E.OCEAN)
        co.predict()
        eng.predict(kind=E.OCEAN)
        worst = max(worst, rel_inf(eng.outvec_get(region, kind=E.OCEAN), co.outvec),
                    rel_inf(eng.state_get(region, kind=E.OCEAN), co.x))
    assert worst < TOL_STEP * 10
    # the atmosphere reservoir of the same region is untouched by ocean steps
    assert np.array_equal(eng.state_get(region), np.zeros(wa["n"]))
    eng.close()


@pytest.mark.parametrize("region,m", [(555, 500), (23, 900)])
def test_predict_slab_hybrid_single_region(E, region, m):
    """predict_slab (src/mod_slab_ocean_reservoir.f90:1268-1316), the ocean reservoir with ml_only_ocean = .False.:
    no leak term, chunk_size_prediction model columns, and local_model <- the standardised outvec after every step"""
    wa = region_weights(1152, region, m=300, sst_bool_input=True, with_dense_win=False)
    wo = ocean_weights(1152, region, m=m, mean=wa["mean"], std=wa["std"], with_dense_win=False, hybrid=True)
    assert wo["S"] == wo["P"] == 8
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    upload_atmo(eng, wa)
    eng.region_upload(region, wo["rows"], wo["cols"], wo["vals"], wo["wout"], wo["mean"], wo["std"],
                      win_compact=wo["winc"], win_col=wo["wcol"], D=wo["D"], kind=E.OCEAN, leakage=0.4,   # ignored: no leak term
                      sst_mean=wo["mean"][wo["sst_idx"] - 1], sst_std=wo["std"][wo["sst_idx"] - 1])
    eng.finalize()
    co = c_ocean(wo)
    rng = np.random.default_rng(region + m)
    x0 = 0.3 * rng.standard_normal(wo["n"])
    lm0 = rng.standard_normal(wo["S"])           # start_prediction_slab: the last observed SST / OHTC tile (:861-864)
    co.x[:] = x0
    co.local_model[:] = lm0
    eng.state_set(region, x0, kind=E.OCEAN)
    eng.local_model_set(region, lm0, kind=E.OCEAN)
    worst = 0.0
    for step in range(5):
        fb = rng.standard_normal(wo["D"])
        co.feedback[:] = fb
        eng.feedback_set(region, fb, kind=E.OCEAN)
        co.predict()
        eng.predict(kind=E.OCEAN)
        worst = max(worst, rel_inf(eng.outvec_get(region, kind=E.OCEAN), co.outvec),
                    rel_inf(eng.state_get(region, kind=E.OCEAN), co.x),
                    rel_inf(eng.local_model_get(region, kind=E.OCEAN), co.local_model))
    assert worst < TOL_STEP * 100           # five closed steps: the prediction feeds back through local_model
    eng.close()


def test_ocean_synchronize(E):
    region = 700
    wa = region_weights(1152, region, m=300, sst_bool_input=True, with_dense_win=False)
    wo = ocean_weights(1152, region, m=900, mean=wa["mean"], std=wa["std"], with_dense_win=False)
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    upload_atmo(eng, wa)
    upload_ocean(eng, E, wo)
    eng.finalize()
    co = c_ocean(wo)
    series = syn.ar1_series(wo["D"], 12, np.random.default_rng(1))
    co.synchronize(series, 12)                                  # slab synchronize, T = 11/2 in the reference
    eng.synchronize(region, series, kind=E.OCEAN)
    assert rel_inf(eng.state_get(region, kind=E.OCEAN), co.x) < TOL_LOOP
    eng.close()


@pytest.fixture(scope="module")
def coupled_model(E):
    """full 1152-region tiling, minimal atmosphere reservoirs + ocean reservoirs on the 70 % 'ocean' regions"""
    ws = [region_weights(1152, r, m=300, with_dense_win=False) for r in range(1152)]
    wos = {r: ocean_weights(1152, r, m=300, mean=ws[r]["mean"], std=ws[r]["std"], with_dense_win=False)
           for r in range(1152) if sst_input_mask(r)}
    eng = E.Engine(number_of_regions=1152, sst_prescribed=False)
    for w in ws:
        upload_atmo(eng, w)
    for w in wos.values():
        upload_ocean(eng, E, w)
    eng.finalize()
    rcs = [c_region(w) for w in ws]
    cos = {r: c_ocean(w) for r, w in wos.items()}
    return ws, wos, eng, rcs, cos


def test_coupled_loop_with_ocean_steps(E, coupled_model):
    ws, wos, eng, rcs, cos = coupled_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.ocean_ring_reset()
    sst_mean = np.array([w["mean"][-1] for w in ws])
    sst_std = np.array([w["std"][-1] for w in ws])
    rng = np.random.default_rng(77)
    for w, rc in zip(ws, rcs):
        fb, lm = rng.standard_normal(w["D"]), rng.standard_normal(w["S"])
        rc.feedback[:], rc.local_model[:], rc.x[:] = fb, lm, 0.0
        eng.feedback_set(w["region"], fb)
        eng.local_model_set(w["region"], lm)
        eng.state_set(w["region"], np.zeros(w["n"]))
    for r, co in cos.items():
        # start_prediction_slab: feedback = last observed input, outvec = its SST/OHTC interior in K
        fb = rng.standard_normal(wos[r]["D"])
        ov = 285.0 + 5.0 * rng.random(8)
        co.feedback[:], co.outvec[:], co.x[:] = fb, ov, 0.0
        co.ring[:] = 0.0
        eng.feedback_set(r, fb, kind=E.OCEAN)
        eng.outvec_set(r, ov, kind=E.OCEAN)
        eng.state_set(r, np.zeros(wos[r]["n"]), kind=E.OCEAN)
    has = np.array([1 if r in cos else 0 for r in range(1152)], dtype=np.int32)
    nthreads = 8
    probe = [0, 3, 23, 555, 556, 557, 700, 1128, 1151]
    for t in range(1, 58):                                    # ocean steps at t = 28 and 56, ring wraps at 28
        oc.predict_all(rcs, nthreads=nthreads)
        eng.predict()
        if (t * 6) % 168 == 0:                                # src/parallelmain.f90:238
            for co in cos.values():
                co.predict()
            eng.predict(kind=E.OCEAN)
        oo = np.zeros((1152, 4))
        for r, co in cos.items():
            oo[r] = co.outvec[:4]
        gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
        ge = eng.step_exchange_begin(t)
        for a, b in zip(ge, gc):
            assert rel_inf(a, b) < TOL_LOOP
        f4c, f2c = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
        f4e, f2e = oc.host_stub(ge[0], ge[1], G["clim4d"], G["clim2d"])
        oc.step_scatter(rcs, True, True, False, *gc, f4c, f2c, G["tisr"], sst_mean, sst_std, nthreads=nthreads)
        for r, co in cos.items():
            co.build_feedback(rcs[r], t, gc[3])
        eng.step_exchange_end(t, f4e, f2e, G["tisr"])
        if t in (1, 2, 27, 28, 29, 55, 56, 57):
            for r in probe:
                assert rel_inf(eng.feedback_get(r), rcs[r].feedback) < TOL_LOOP
                if r in cos:
                    assert rel_inf(eng.feedback_get(r, kind=E.OCEAN), cos[r].feedback) < TOL_LOOP
                    assert rel_inf(eng.outvec_get(r, kind=E.OCEAN), cos[r].outvec) < TOL_LOOP
    # regions without an ocean reservoir report 272.0 (or the base grid where the land mask says so)
    wsst = ge[3]
    r = next(r for r in range(1152) if r not in cos)
    xs, xe, ys, ye, *_ = oc.getxyresextent(1152, r)
    tile, base, mask = wsst[xs - 1:xe, ys - 1:ye], G["base_sst"][xs - 1:xe, ys - 1:ye], G["sea_mask"][xs - 1:xe, ys - 1:ye]
    assert np.array_equal(tile, np.where(mask > 0.0, np.maximum(base, 272.0), 272.0))


def test_ocean_feedback_assembly_is_exact(E, coupled_model):
    # ring + mean + SST standardise is copies and fixed-order FP64 arithmetic: feed the oracle the engine's own
    # atmosphere feedback and SST grid and require BIT equality of the ocean feedback over a ring wrap
    ws, wos, eng, rcs, cos = coupled_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.ocean_ring_reset()
    for co in cos.values():
        co.ring[:] = 0.0
    probe = [r for r in (0, 3, 555, 556, 1151, 24 * 47 + 23) if r in cos]
    for r in probe:
        cos[r].feedback[:] = eng.feedback_get(r, kind=E.OCEAN)
    for t in range(1, 31):
        eng.predict()
        g = eng.step_exchange_begin(t)
        f4, f2 = oc.host_stub(g[0], g[1], G["clim4d"], G["clim2d"])
        eng.step_exchange_end(t, f4, f2, G["tisr"])
        for r in probe:
            rcs[r].feedback[:] = eng.feedback_get(r)
            cos[r].build_feedback(rcs[r], t, g[3])
            assert np.array_equal(eng.feedback_get(r, kind=E.OCEAN), cos[r].feedback)


def test_ocean_training_gram_and_fit(E):
    region = 555
    wa = region_weights(1152, region, m=300, sst_bool_input=True, with_dense_win=False)
    wo = ocean_weights(1152, region, m=500, mean=wa["mean"], std=wa["std"], with_dense_win=False)
    wo["wout"] = np.zeros_like(wo["wout"])
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    upload_atmo(eng, wa)
    upload_ocean(eng, E, wo)
    eng.finalize()
    co = c_ocean(wo)
    bs, discard = 11, 4
    phases = [syn.ar1_series(wo["D"], discard + 4 * bs + 3, np.random.default_rng(90 + p)) for p in range(2)]
    co.train_init(bs)
    eng.train_begin([region], bs, kind=E.OCEAN)
    for td in phases:
        co.train_phase(td, None, discard)
        eng.train_feed([td], None, discard)
    sxs, sxt = eng.train_gram_get(region)
    assert rel_inf(sxs, co.sxs) < 1e-12
    assert rel_inf(sxt, co.sxt) < 1e-12
    A = co.sxs.copy()
    B = co.sxt.copy()
    assert co.fit(beta_res=1e-4) == 0
    info = eng.train_solve(1e-4)
    assert info[0] == 0
    eng.train_end()
    d = np.arange(wo["n"])
    A[d, d] += 1e-4
    for wout in (eng.wout_get(region, kind=E.OCEAN), co.wout):
        X = wout.T
        res = np.linalg.norm(A.T @ X - B.T) / (np.linalg.norm(A) * np.linalg.norm(X) + np.linalg.norm(B))
        assert res < 1e-13
    eng.close()


def test_coupled_loop_overlapped_mode_matches_sequential(E, coupled_model):
    """the overlapped step with ocean reservoirs: the ocean ring/SST feedback is rebuilt inside exchange_begin, the
    ocean reservoirs still step on the caller's schedule; grids must match the sequential engine loop"""
    ws, wos, eng, rcs, cos = coupled_model
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    rng = np.random.default_rng(5)
    start = {w["region"]: (0.1 * rng.standard_normal(w["n"]), rng.standard_normal(w["D"]), rng.standard_normal(w["S"])) for w in ws}
    ostart = {r: (0.1 * rng.standard_normal(wos[r]["n"]), rng.standard_normal(wos[r]["D"]), 285.0 + 5.0 * rng.random(8)) for r in wos}

    def run(overlap):
        for r, (x0, fb, lm) in start.items():
            eng.state_set(r, x0)
            eng.feedback_set(r, fb)
            eng.local_model_set(r, lm)
        for r, (x0, fb, ov) in ostart.items():
            eng.state_set(r, x0, kind=E.OCEAN)
            eng.feedback_set(r, fb, kind=E.OCEAN)
            eng.outvec_set(r, ov, kind=E.OCEAN)
        eng.ocean_ring_reset()
        eng.set_overlap(overlap)
        out = []
        for t in range(1, 31):                                 # one ocean step at t = 28, ring wraps
            eng.predict()
            if (t * 6) % 168 == 0:
                eng.predict(kind=E.OCEAN)
            if overlap:
                eng.set_tisr(G["tisr"])
            g = eng.step_exchange_begin(t)
            f4, f2 = oc.host_stub(g[0], g[1], G["clim4d"], G["clim2d"])
            eng.step_exchange_end(t, f4, f2, None if overlap else G["tisr"])
            out.append([a.copy() for a in g])
        ofb = {r: eng.feedback_get(r, kind=E.OCEAN) for r in (0, 3, 555)}
        eng.set_overlap(False)
        return out, ofb

    seq, ofb_seq = run(False)
    ovl, ofb_ovl = run(True)
    for t in range(30):
        for a, b in zip(ovl[t], seq[t]):
            assert rel_inf(a, b) < 1e-10
    for r in ofb_seq:
        assert rel_inf(ofb_ovl[r], ofb_seq[r]) < 1e-10
