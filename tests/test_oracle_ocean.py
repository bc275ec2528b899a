"""Slab-ocean reservoir: C oracle vs NumPy oracle -- sizes, predict_slab_ml, the ocean
feedback assembly (Appendix C intended semantics), the ocean target tiler and ML-only training."""
import numpy as np

from helpers import (c_ocean, c_region, initial_grids, np_ocean, np_region, oc, ocean_weights, on, region_weights,
                     rel_inf, syn)


def test_ocean_sizes_match_initialize_slab_ocean_model():
    # interior tile: 4x4 halo -> D = 4*16+16 + 3*16 = 128, P = 8, n = NINT(4000/128)*128 = 3968
    r = oc.OceanRegion(1152, 555)
    assert (r.D, r.P, r.S, r.n, r.A) == (128, 8, 0, 3968, 80)
    assert r.k == int((6.0 / 4000.0) * 3968 * 3968)
    assert (r.g.sst_start, r.g.sst_end, r.g.tisr_start, r.g.tisr_end, r.g.ohtc_start, r.g.ohtc_end) == (81, 96, 97, 112, 113, 128)
    assert r.g.sst_mean_std_idx == 36 and r.g.ohtc_mean_std_idx == 1 and r.d.leakage == 1.0
    # polar tile: 4x3 halo -> D = 96, n = NINT(4000/96)*96 = 42*96
    p = oc.OceanRegion(1152, 24 * 5)
    assert (p.D, p.P, p.n, p.A) == (96, 8, 4032, 60)
    for reg in (555, 24 * 5, 47, 1151):
        c, n = oc.OceanRegion(1152, reg), on.OceanRegion(1152, reg)
        assert (c.n, c.D, c.P, c.k, c.A, c.L) == (n.n, n.D, n.P, n.k, n.A, n.mean_std_length)


def test_predict_slab_ml_c_vs_numpy():
    w = ocean_weights(1152, 555, m=500)
    rc, rn = c_ocean(w), np_ocean(w)
    rng = np.random.default_rng(3)
    x = np.zeros(w["n"])
    for step in range(5):
        fb = rng.standard_normal(w["D"])
        rc.feedback[:] = fb
        rn.feedback = fb.copy()
        rc.predict()
        x, out = on.predict_slab_ml(rn, x)
        assert rel_inf(rc.x, x) < 1e-13
        assert rel_inf(rc.outvec, out) < 1e-13
    # leakage 1: the state is exactly tanh(...) -- no memory of the previous x beyond A x
    assert np.all(np.abs(rc.x) <= 1.0)


def test_ocean_feedback_ring_and_sst_slot():
    for reg in (555, 0, 24 * 47 + 23):          # interior, south-pole + periodic west edge, north-pole east edge
        wa = region_weights(1152, reg, m=450, sst_bool_input=True)
        wo = ocean_weights(1152, reg, m=500, mean=wa["mean"], std=wa["std"])
        ca, na = c_region(wa), np_region(wa)
        co, no = c_ocean(wo), np_ocean(wo)
        G = initial_grids()
        rng = np.random.default_rng(11 + reg)
        held = rng.standard_normal(wo["D"])
        co.feedback[:] = held
        no.feedback = held.copy()
        for t in range(1, 31):                  # wraps the 27-slot ring
            fb = rng.standard_normal(wa["D"])
            ca.feedback[:] = fb
            na.feedback = fb.copy()
            wsst = np.asfortranarray(G["base_sst"] + 0.1 * t)
            co.build_feedback(ca, t, wsst)
            on.ocean_feedback(no, na, t, wsst)
            assert np.array_equal(co.feedback, no.feedback)
            assert np.array_equal(co.ring, no.ring)
        A, ixy = wo["A"], wo["A"] // 5
        # TISR and OHTC slots keep their start_prediction_slab values
        assert np.array_equal(co.feedback[A + ixy:], held[A + ixy:])
        # slot 30 -> ring index (30-1)%27 = 2 holds the last atmosphere slice
        a0 = ca.g.atmo3d_end - 4 * ixy
        assert np.array_equal(co.ring[:, 2], fb[a0:a0 + A])
        # the mean divides by 27 even before the ring is full
        co2 = c_ocean(wo)
        ca.feedback[:] = 1.0
        co2.build_feedback(ca, 1, wsst)
        assert np.allclose(co2.feedback[:A], 1.0 / 27.0, rtol=0, atol=1e-16)


def test_ocean_target_rows():
    for reg in (555, 0, 24 * 47 + 23):
        wo = ocean_weights(1152, reg, m=500)
        co, no = c_ocean(wo), np_ocean(wo)
        rng = np.random.default_rng(5)
        sv = np.asfortranarray(rng.standard_normal((wo["D"], 4)))
        tc = co.target(sv)
        tn = on.tile_full_input_to_target_data_ocean(no, sv)
        assert tc.shape == (8, 4) and np.array_equal(tc, tn)


def test_ocean_training_c_vs_numpy():
    wo = ocean_weights(1152, 555, m=500)         # n = 512
    co, no = c_ocean(wo), np_ocean(wo)
    bs, discard = 6, 4
    phases = [syn.ar1_series(wo["D"], discard + 3 * bs, np.random.default_rng(70 + p)) for p in range(2)]
    co.train_init(bs)
    for td in phases:
        co.train_phase(td, None, discard)
    sxs_c, sxt_c = co.sxs.copy(), co.sxt.copy()
    assert co.fit(beta_res=1e-4) == 0
    wout_n, sxs_n, sxt_n, info = on.train_ml(no, phases, bs, discard, 1e-4, on.tile_full_input_to_target_data_ocean)
    assert info == 0
    d = np.arange(wo["n"])
    sxs_n_noreg = sxs_n.copy()
    sxs_n_noreg[d, d] -= 1e-4
    assert rel_inf(sxs_c, sxs_n_noreg) < 1e-12
    assert rel_inf(sxt_c, sxt_n) < 1e-12
    A = sxs_n.T
    for wout in (co.wout, wout_n):
        X = wout.T
        res = np.linalg.norm(A @ X - sxt_n.T) / (np.linalg.norm(A) * np.linalg.norm(X) + np.linalg.norm(sxt_n))
        assert res < 1e-13
