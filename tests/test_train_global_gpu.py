"""Training from the device-resident GLOBAL series (sml_train_global_series / sml_train_feed_global):
the engine tiles and standardises every region's input, imperfect-model and target columns on the fly.  The oracle
builds the same per-region series column by column with the reference's tilers (the feedback construction of
sendrecievegrid), trains on them, and the Gram must agree to 1e-12; feeding those per-region series through
sml_train_feed must give the BIT-identical Gram.  -m gpu."""
import importlib

import numpy as np
import pytest

from helpers import c_region, initial_grids, oc, region_weights, rel_inf, syn

pytestmark = pytest.mark.gpu


def test_global_series_feed_matches_oracle_tiling_and_per_region_feed():
    E = importlib.import_module("speedy-ml_b200.engine")
    regions = [0, 1, 2, 3]                                   # south-pole + interior tiles of the first x column
    ws = {r: region_weights(1152, r, m=450, with_dense_win=False) for r in regions}
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)
    for r in regions:
        w = ws[r]
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], None, w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"], S=w["S"], P=w["P"])
    eng.finalize()
    lay = E.global_layout()
    F0 = initial_grids()
    rng = np.random.default_rng(17)
    T = 64
    G = np.zeros((lay["g_total"], T), order="F")
    F = np.zeros((lay["f_total"], T), order="F")
    for t in range(T):
        w4d = F0["clim4d"] * (1.0 + 0.02 * rng.standard_normal(F0["clim4d"].shape))
        w4d[3] = np.maximum(w4d[3], 0.000001)
        w2d = F0["clim2d"] + 0.01 * rng.standard_normal((96, 48))
        wp = np.log1p(np.abs(rng.standard_normal((96, 48))))        # log(1 + p/eps)-like, >= 0
        wsst = np.maximum(F0["base_sst"] + rng.standard_normal((96, 48)), 272.0)
        tisr = np.abs(F0["tisr"] * (1.0 + 0.1 * rng.standard_normal((96, 48))))
        G[:, t] = np.concatenate([a.ravel(order="F") for a in (w4d, w2d, wp, wsst, tisr)])
        f4 = 0.97 * w4d + 0.03 * F0["clim4d"]
        f2 = 0.97 * w2d
        F[:, t] = np.concatenate([f4.ravel(order="F"), f2.ravel(order="F")])

    # oracle: per-region series through the reference's tilers + standardisation, column by column
    rcs = [c_region(ws[r]) for r in regions]
    sst_mean = np.array([ws[r]["mean"][-1] for r in regions])
    sst_std = np.array([ws[r]["std"][-1] for r in regions])
    td = {r: np.zeros((ws[r]["D"], T), order="F") for r in regions}
    im = {r: np.zeros((ws[r]["S"], T), order="F") for r in regions}
    o = lay
    for t in range(T):
        g = G[:, t]
        grids = (g[:o["w2d"]].reshape((4, 96, 48, 8), order="F"), g[o["w2d"]:o["precip"]].reshape((96, 48), order="F"),
                 g[o["precip"]:o["sst"]].reshape((96, 48), order="F"), g[o["sst"]:o["tisr"]].reshape((96, 48), order="F"))
        f4 = F[:o["w2d"], t].reshape((4, 96, 48, 8), order="F")
        f2 = F[o["w2d"]:, t].reshape((96, 48), order="F")
        tisr = g[o["tisr"]:].reshape((96, 48), order="F")
        oc.step_scatter(rcs, True, True, False, *grids, f4, f2, tisr, sst_mean, sst_std, nthreads=2)
        for r, rc in zip(regions, rcs):
            td[r][:, t] = rc.feedback
            im[r][:, t] = rc.local_model

    first, stride, ncols, discard, bs = 1, 2, 27, 3, 6       # phase = columns 1, 3, 5, ...
    sel = first + stride * np.arange(ncols)
    for rc, r in zip(rcs, regions):
        rc.train_init(bs)
        rc.train_phase(np.asfortranarray(td[r][:, sel]), np.asfortranarray(im[r][:, sel]), discard)

    eng.train_global_series(G, F)
    eng.train_begin(regions, bs)
    eng.train_feed_global(first, stride, ncols, discard)
    gram_global = {r: eng.train_gram_get(r) for r in regions}
    with pytest.raises(E.EngineError):                        # runs past the resident columns
        eng.train_feed_global(first, stride, T, discard)
    eng.train_end()
    eng.train_begin(regions, bs)
    eng.train_feed([np.asfortranarray(td[r][:, sel]) for r in regions], [np.asfortranarray(im[r][:, sel]) for r in regions], discard)
    gram_local = {r: eng.train_gram_get(r) for r in regions}
    eng.train_end()
    eng.train_global_release()
    for rc, r in zip(rcs, regions):
        assert rel_inf(gram_global[r][0], rc.sxs) < 1e-12 and rel_inf(gram_global[r][1], rc.sxt) < 1e-12
        assert np.array_equal(gram_global[r][0], gram_local[r][0]) and np.array_equal(gram_global[r][1], gram_local[r][1])
    eng.close()


def test_conditioning_stats_from_resident_series():
    """grid%mean / grid%std of every local region computed on the device from the resident global series vs the
    reference's formulas restated in NumPy (two-pass population std; one-pass form for precip; SST gate std > 0.2)"""
    from helpers import on
    E = importlib.import_module("speedy-ml_b200.engine")
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)      # regions 0..3; no region is uploaded
    regions = eng.region_indices
    lay = E.global_layout()
    F0 = initial_grids()
    rng = np.random.default_rng(23)
    T = 40
    w4d_t = np.stack([F0["clim4d"] * (1.0 + 0.03 * rng.standard_normal(F0["clim4d"].shape)) for _ in range(T)], axis=-1)
    logp_t = np.stack([F0["clim2d"] + 0.01 * rng.standard_normal((96, 48)) for _ in range(T)], axis=-1)
    precip_t = np.stack([np.log1p(np.abs(rng.standard_normal((96, 48)))) for _ in range(T)], axis=-1)
    sst_t = np.stack([np.maximum(F0["base_sst"] + rng.standard_normal((96, 48)), 272.0) for _ in range(T)], axis=-1)
    sst_t[0:4, 0:3, :] = 272.0          # region 0's whole halo block (x 96,1,2,3 wraps; rows 1..3): constant -> no SST input
    sst_t[95, 0:3, :] = 272.0
    tisr_t = np.stack([np.abs(F0["tisr"] * (1.0 + 0.1 * rng.standard_normal((96, 48)))) for _ in range(T)], axis=-1)
    G = np.zeros((lay["g_total"], T), order="F")
    for t in range(T):
        G[:, t] = np.concatenate([a.ravel(order="F") for a in (w4d_t[..., t], logp_t[..., t], precip_t[..., t], sst_t[..., t], tisr_t[..., t])])
    eng.train_global_series(G, np.zeros((lay["f_total"], T), order="F"))
    first, stride, ncols = 2, 3, 12
    sel = first + stride * np.arange(ncols)
    mean, std, sst_in = eng.conditioning_stats(first, stride, ncols)
    assert mean.shape == (4, 36)
    for i, r in enumerate(regions):
        m, s, any_change = on.conditioning_stats(w4d_t[..., sel], logp_t[..., sel], tisr_t[..., sel], precip_t[..., sel],
                                                 sst_t[..., sel], 1152, r, 1)
        assert rel_inf(mean[i], m) < 1e-12 and rel_inf(std[i], s) < 1e-11
        assert bool(sst_in[i]) == any_change
    assert not sst_in[0] and mean[0, 35] == 0.0 and std[0, 35] == 0.0 and sst_in[1]
    eng.train_global_release()
    eng.close()


def test_condition_raw_series_on_device():
    """unit conversion, floors, precip accumulation over the time step and log transform (get_training_data,
    src/mod_reservoir.f90:362-395) applied in place to the resident raw series"""
    from helpers import on
    E = importlib.import_module("speedy-ml_b200.engine")
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)
    lay = E.global_layout()
    rng = np.random.default_rng(31)
    T, period, eps = 30, 6, 0.001
    w4d_t = rng.standard_normal((4, 96, 48, 8, T))
    w4d_t[3] = 0.004 * rng.standard_normal((96, 48, 8, T))            # kg/kg, some negative
    logp_t = 0.05 * rng.standard_normal((96, 48, T))
    precip_t = np.where(rng.random((96, 48, T)) < 0.6, 0.0, 0.002 * rng.standard_normal((96, 48, T)))
    sst_t = 280.0 + 10.0 * rng.standard_normal((96, 48, T))
    tisr_t = 1.0e5 * rng.standard_normal((96, 48, T))
    G = np.zeros((lay["g_total"], T), order="F")
    for t in range(T):
        G[:, t] = np.concatenate([a.ravel(order="F") for a in (w4d_t[..., t], logp_t[..., t], precip_t[..., t], sst_t[..., t], tisr_t[..., t])])
    eng.train_global_series(G, np.zeros((lay["f_total"], T), order="F"))
    eng.condition_series(period, eps)
    with pytest.raises(E.EngineError):
        eng.condition_series(period, eps)                              # once per upload
    w4c, tisrc, pc, sstc = on.condition_raw_series(w4d_t, tisr_t, precip_t, sst_t, period, eps)
    # read the conditioned series back through the statistics of a region whose halo we can slice: compare all slots
    mean, std, _ = eng.conditioning_stats(0, 1, T)
    for i, r in enumerate(eng.region_indices):
        m, s, _ = on.conditioning_stats(w4c, logp_t, tisrc, pc, sstc, 1152, r, 1)
        assert rel_inf(mean[i], m) < 1e-12 and rel_inf(std[i], s) < 1e-11
    assert pc.min() >= 0.0 and w4c[3].min() >= 0.000001 and sstc.min() >= 272.0
    eng.close()


def test_device_input_noise_for_global_feed(monkeypatch):
    """u*(1 + noisemag*g) with precip noised in linear space (gaussian_noise_1d_function_precip,
    src/mod_utilities.f90:1410-1464): the arithmetic is checked exactly against the NumPy restatement given the engine's
    own N(0,1) draws; the draws are checked statistically and for reproducibility (counter-based generator)."""
    from helpers import on
    E = importlib.import_module("speedy-ml_b200.engine")
    regions = [0, 1, 2, 3]
    ws = {r: region_weights(1152, r, m=450, with_dense_win=False) for r in regions}
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)
    for r in regions:
        w = ws[r]
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], None, w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"], S=w["S"], P=w["P"])
    eng.finalize()
    lay = E.global_layout()
    F0 = initial_grids()
    rng = np.random.default_rng(41)
    T = 48
    G = np.zeros((lay["g_total"], T), order="F")
    for t in range(T):
        w4d = F0["clim4d"] * (1.0 + 0.02 * rng.standard_normal(F0["clim4d"].shape))
        G[:, t] = np.concatenate([w4d.ravel(order="F"), (F0["clim2d"] + 0.01 * rng.standard_normal((96, 48))).ravel(order="F"),
                                  np.log1p(np.abs(rng.standard_normal(96 * 48))), np.maximum(F0["base_sst"], 272.0).ravel(order="F"),
                                  np.abs(F0["tisr"]).ravel(order="F")])
    Fz = np.asfortranarray(np.tile(G[:lay["f_total"], :1], (1, T)))
    eng.train_global_series(G, Fz)
    bs = 6
    eng.train_begin(regions, bs)
    eng.train_set_noise(0.2, seed=7, precip_epsilon=0.001)
    draws = []
    for r in regions:
        ca = oc.Region(1152, r, sst_bool_input=ws[r]["sst_bool_input"])
        for col in (0, 5, 17):
            clean, g, noisy = eng.train_noise_sample(r, 1, 2, col)
            clean2, g2, noisy2 = eng.train_noise_sample(r, 1, 2, col)
            assert np.array_equal(g, g2) and np.array_equal(noisy, noisy2)          # reproducible
            pi = ca.g.precip_mean_std_idx - 1
            want = on.gaussian_noise_1d_function_precip(clean, g, 0.2, ca.g.precip_start, ca.g.precip_end,
                                                        ws[r]["mean"][pi], ws[r]["std"][pi], 0.001)
            assert rel_inf(noisy, want) < 1e-13
            draws.append(g)
    g_all = np.concatenate(draws)
    assert len(np.unique(g_all)) == g_all.size                                   # no repeated draws across regions / columns
    for r in regions:                                                            # more samples for the moments
        for col in range(20):
            draws.append(eng.train_noise_sample(r, 0, 1, col)[1])
    g_all = np.concatenate(draws)
    assert abs(g_all.mean()) < 0.02 and abs(g_all.var() - 1.0) < 0.03 and np.abs(g_all).max() < 6.5
    # a noisy phase changes the Gram, reproducibly; noise off restores the clean one
    eng.train_feed_global(1, 2, 20, 2)
    noisy_gram = eng.train_gram_get(1)[0]
    eng.train_end()
    eng.train_begin(regions, bs)
    eng.train_feed_global(1, 2, 20, 2)
    assert np.array_equal(eng.train_gram_get(1)[0], noisy_gram)
    eng.train_end()
    # the per-time-step launches (k_train_update) and the in-kernel time loop (k_train_stategen) see the same noised inputs
    for route in ("steps", "kernel", "ring"):
        monkeypatch.setenv("SML_TRAIN_STATEGEN", route)
        eng.train_begin(regions, bs)
        eng.train_feed_global(1, 2, 20, 2)
        assert eng.train_stategen_route() == route
        assert np.array_equal(eng.train_gram_get(1)[0], noisy_gram)
        eng.train_end()
    monkeypatch.delenv("SML_TRAIN_STATEGEN", raising=False)
    eng.train_set_noise(0.0)
    eng.train_begin(regions, bs)
    eng.train_feed_global(1, 2, 20, 2)
    clean_gram = eng.train_gram_get(1)[0]
    eng.train_end()
    assert rel_inf(noisy_gram, clean_gram) > 1e-4
    eng.close()
