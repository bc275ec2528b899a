"""Training path on the GPU (state generation + DMMA Gram + ridge solve) vs the CPU oracle.  -m gpu."""
import importlib

import numpy as np
import pytest
from scipy.linalg import lapack

from helpers import c_region, oc, region_weights, rel_inf, syn
from test_engine_gpu import single_region_engine, upload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


def _series(w, T, seed):
    rng = np.random.default_rng(seed)
    td = syn.ar1_series(w["D"], T, rng)
    im = np.asfortranarray(rng.standard_normal((w["S"], T))) if w["S"] else None
    return td, im


def _residual(A, X, B):
    return np.linalg.norm(A @ X - B) / (np.linalg.norm(A) * np.linalg.norm(X) + np.linalg.norm(B))


@pytest.mark.parametrize("region,m,bs,nb,discard", [(555, 450, 7, 4, 5), (24 * 3, 450, 16, 3, 2), (556, 1300, 98, 7, 40)])
def test_hybrid_training_parity(E, region, m, bs, nb, discard):
    w = region_weights(1152, region, m=m)
    rc = c_region(w)
    eng = single_region_engine(E, w)
    phases = [_series(w, discard + nb * bs + 3, 50 + p) for p in range(2)]   # +3: trailing states are dropped
    rc.train_init(bs)
    eng.train_begin([region], bs)
    for td, im in phases:
        rc.train_phase(td, im, discard)
        eng.train_feed([td], [im], discard)
    sxs_e, sxt_e = eng.train_gram_get(region)
    assert rel_inf(sxs_e, rc.sxs) < 1e-12
    assert rel_inf(sxt_e, rc.sxt) < 1e-12
    assert np.array_equal(sxs_e, sxs_e.T)
    sxs0, sxt0 = rc.sxs.copy(), rc.sxt.copy()
    assert rc.fit(beta_res=1e-3, beta_model=1.0, using_prior=True, prior_val=0.0) == 0
    info = eng.train_solve(1e-3, 1.0, True, 0.0)
    assert info[0] == 0
    wout_e = eng.wout_get(region)
    # ridge system: (sxs + diag)^T X = sxt^T ; judge by relative residual
    N, S = w["n"] + w["S"], w["S"]
    A = sxs0.copy()
    d = np.arange(N)
    A[d[:S], d[:S]] += 1.0
    A[d[S:], d[S:]] += 1e-6
    for wout in (wout_e, rc.wout):
        assert _residual(A.T, wout.T, sxt0.T) < 1e-13
    # downstream forecast parity: engine-trained vs oracle-trained W_out, 100 open-loop steps
    rng = np.random.default_rng(77)
    series = syn.ar1_series(w["D"], 100, rng)
    model = rng.standard_normal((w["S"], 100))
    rc.x[:] = 0.0
    eng.state_set(region, np.zeros(w["n"]))
    worst = 0.0
    for t in range(100):
        rc.feedback[:] = series[:, t]
        rc.local_model[:] = model[:, t]
        eng.feedback_set(region, series[:, t])
        eng.local_model_set(region, model[:, t])
        rc.predict()
        eng.predict()
        worst = max(worst, rel_inf(eng.outvec_get(region), rc.outvec))
    assert worst < 1e-8
    eng.train_end()
    eng.close()


def test_prior_and_plain_beta_variants(E):
    w = region_weights(1152, 555, m=450)
    td, im = _series(w, 3 + 3 * 8, 5)
    for using_prior, prior_val in ((True, 0.7), (False, 0.0)):
        rc = c_region(w)
        eng = single_region_engine(E, w)
        rc.train_init(8)
        rc.train_phase(td, im, 3)
        eng.train_begin([555], 8)
        eng.train_feed([td], [im], 3)
        sxs0, sxt0 = rc.sxs.copy(), rc.sxt.copy()
        assert rc.fit(beta_res=0.05, beta_model=0.5, using_prior=using_prior, prior_val=prior_val) == 0
        assert eng.train_solve(0.05, 0.5, using_prior, prior_val)[0] == 0
        N, S = w["n"] + w["S"], w["S"]
        d = np.arange(N)
        A, B = sxs0.copy(), sxt0.copy()
        if using_prior:
            A[d[:S], d[:S]] += 0.25
            A[d[S:], d[S:]] += 0.0025
            B[d[:S], d[:S]] += prior_val * 0.25
        else:
            A[d[:S], d[:S]] += 0.5
            A[d[S:], d[S:]] += 0.05
        assert _residual(A.T, eng.wout_get(555).T, B.T) < 1e-13
        assert rel_inf(eng.wout_get(555), rc.wout) < 1e-7   # well conditioned here (beta large)
        eng.train_end()
        eng.close()


def test_ml_only_training_with_restart_quirk(E):
    w = region_weights(1152, 555, m=450, ml_only=True)
    rc = c_region(w)
    eng = single_region_engine(E, w)
    bs, discard = 5, 3
    td, _ = _series(w, discard + 4 * bs, 9)
    rc.train_init(bs)
    rc.train_phase(td, None, discard)
    eng.train_begin([555], bs)
    eng.train_feed([td], None, discard)
    sxs_e, sxt_e = eng.train_gram_get(555)
    assert rel_inf(sxs_e, rc.sxs) < 1e-12
    assert rel_inf(sxt_e, rc.sxt) < 1e-12
    sxs0, sxt0 = rc.sxs.copy(), rc.sxt.copy()
    assert rc.fit(beta_res=1e-2) == 0
    assert eng.train_solve(1e-2)[0] == 0
    A = sxs0 + 1e-2 * np.eye(w["n"])
    assert _residual(A.T, eng.wout_get(555).T, sxt0.T) < 1e-13
    eng.train_end()
    eng.close()


def test_wave_of_regions_with_different_shapes_and_slab_boundaries(E, monkeypatch):
    monkeypatch.setenv("SML_TRAIN_SLAB", "16")   # force several K slabs per phase
    regions = [0, 1, 2, 3]
    ws = {r: region_weights(1152, r, m=450) for r in regions}
    assert len({ws[r]["n"] for r in regions}) > 1
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)   # rank 0 owns regions 0..3
    assert eng.region_indices == regions
    for r in regions:
        upload(eng, ws[r])
    eng.finalize()
    bs, discard = 10, 4
    series = {r: _series(ws[r], discard + 5 * bs, 30 + r) for r in regions}
    eng.train_begin(regions, bs)
    eng.train_feed([series[r][0] for r in regions], [series[r][1] for r in regions], discard)
    for r in regions:
        rc = c_region(ws[r])
        rc.train_init(bs)
        rc.train_phase(series[r][0], series[r][1], discard)
        sxs_e, sxt_e = eng.train_gram_get(r)
        assert rel_inf(sxs_e, rc.sxs) < 1e-12
        assert rel_inf(sxt_e, rc.sxt) < 1e-12
    st = eng.train_stats()
    assert st["gram_flops_useful"] > 0 and st["gram_ms"] > 0
    eng.train_end()
    eng.close()


def test_mldivide_matches_dgesv(E, monkeypatch):
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=1152)
    rng = np.random.default_rng(0)
    for n, k in ((1, 1), (7, 3), (333, 136), (1200, 40)):
        A = rng.standard_normal((n, n)) + 0.1 * n * np.eye(n)
        B = rng.standard_normal((n, k))
        X, info = eng.mldivide(A, B)
        _, _, Xl, infol = lapack.dgesv(A, B)
        assert info == 0 and infol == 0
        assert rel_inf(X, Xl) < 1e-10
        assert _residual(A, X, B) < 1e-14
    # no diagonal dominance: every column needs its interchange, pivots come from any CTA of the panel launch
    for n, k, gmax in ((500, 5, None), (777, 3, 3), (130, 2, 1)):
        if gmax is None:
            monkeypatch.delenv("SML_LU_GMAX", raising=False)
        else:
            monkeypatch.setenv("SML_LU_GMAX", str(gmax))     # few CTAs: many rows per CTA
        A = rng.standard_normal((n, n))
        A[:, 7] = np.sign(A[:, 7])                            # ties in |a|: the first maximal row must win (idamax)
        B = rng.standard_normal((n, k))
        X, info = eng.mldivide(A, B)
        _, _, Xl, infol = lapack.dgesv(A, B)
        assert info == 0 and infol == 0
        assert rel_inf(X, Xl) < 1e-8
        assert _residual(A, X, B) < 1e-13
    monkeypatch.delenv("SML_LU_GMAX", raising=False)
    # a permutation matrix: pure interchanges
    perm = rng.permutation(200)
    A = np.zeros((200, 200))
    A[np.arange(200), perm] = 2.0
    B = rng.standard_normal((200, 4))
    X, info = eng.mldivide(A, B)
    assert info == 0 and np.allclose(A @ X, B, rtol=0, atol=1e-14)
    # singular: LAPACK info > 0, B is not the solution (src/mod_linalg.f90:147-150 prints and continues)
    A = np.array([[1.0, 2.0], [2.0, 4.0]])
    _, info = eng.mldivide(A, np.ones((2, 1)))
    assert info == 2
    # shape mismatch: 'returning A and B unchanged' (:134-137)
    B = np.ones((2, 1))
    X, info = eng.mldivide(np.eye(3), B)
    assert info == -1 and np.array_equal(X, B)
    eng.close()


def _train_small(E, eng, rc, w, region, bs=8, discard=3, nbatch=3, seed=5):
    td, im = _series(w, discard + nbatch * bs, seed)
    rc.train_init(bs)
    rc.train_phase(td, im, discard)
    eng.train_begin([region], bs)
    eng.train_feed([td], [im], discard)
    return rc.sxs.copy(), rc.sxt.copy()


def test_solver_paths_cholesky_lu_and_fallback(E, monkeypatch):
    """fit_chunk's system is SPD for ridge > 0: the engine factorises it by batched Cholesky (chol.cuh) and falls
    back to dgesv-style LU when a pivot is not positive.  All three routes must solve the reference's system."""
    region = 555
    w = region_weights(1152, region, m=1300)          # N = 1284: 11 panels of 128, last one partial
    N, S = w["n"] + w["S"], w["S"]
    d = np.arange(N)
    results = {}
    for route in ("cholesky", "lu", "fallback"):
        if route == "lu":
            monkeypatch.setenv("SML_SOLVER", "lu")
        else:
            monkeypatch.delenv("SML_SOLVER", raising=False)
        rc = c_region(w)
        eng = single_region_engine(E, w)
        sxs0, sxt0 = _train_small(E, eng, rc, w, region)
        A, B = sxs0.copy(), sxt0.copy()
        if route == "fallback":
            # plain-beta variant with a negative "ridge": A is indefinite, Cholesky must bail out and LU must solve it
            beta_res, beta_model, using_prior = -50.0, -50.0, False
            A[d, d] += -50.0
        else:
            beta_res, beta_model, using_prior = 1e-3, 1.0, True
            A[d[:S], d[:S]] += 1.0
            A[d[S:], d[S:]] += 1e-6
        info = eng.train_solve(beta_res, beta_model, using_prior, 0.0)
        assert info[0] == 0
        assert eng.train_solver_stats() == (1 if route == "cholesky" else 0)
        wout = eng.wout_get(region)
        assert _residual(A.T, wout.T, B.T) < 1e-13
        assert rc.fit(beta_res=beta_res, beta_model=beta_model, using_prior=using_prior, prior_val=0.0) == 0
        assert _residual(A.T, rc.wout.T, B.T) < 1e-13
        results[route] = wout
        eng.train_end()
        eng.close()
    # same system, two factorizations: agreement is limited by cond(A) * eps, the residuals above are the criterion
    assert rel_inf(results["cholesky"] @ np.ones(N), results["lu"] @ np.ones(N)) < 1e-6


def test_solver_wave_mixed_shapes(E):
    """a wave whose regions have different N (different panel counts, different partial last panels)"""
    regions = [0, 1, 2, 3]
    ws = {r: region_weights(1152, r, m=450) for r in regions}
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)
    for r in regions:
        upload(eng, ws[r])
    eng.finalize()
    bs, discard = 10, 4
    series = {r: _series(ws[r], discard + 4 * bs, 60 + r) for r in regions}
    eng.train_begin(regions, bs)
    eng.train_feed([series[r][0] for r in regions], [series[r][1] for r in regions], discard)
    grams = {r: eng.train_gram_get(r) for r in regions}
    info = eng.train_solve(1e-3, 1.0, True, 0.0)
    assert list(info) == [0, 0, 0, 0] and eng.train_solver_stats() == 4
    for r in regions:
        w = ws[r]
        N, S = w["n"] + w["S"], w["S"]
        d = np.arange(N)
        A, B = grams[r][0].copy(), grams[r][1].copy()
        A[d[:S], d[:S]] += 1.0
        A[d[S:], d[S:]] += 1e-6
        assert _residual(A.T, eng.wout_get(r).T, B.T) < 1e-13
    eng.train_end()
    eng.close()


@pytest.mark.parametrize("ml_only", [False, True])
def test_overlapped_schedule_is_bit_identical(E, monkeypatch, ml_only):
    """sml_train_set_overlap: state generation of slab k+1 runs while the Gram of slab k is accumulated on its own
    stream (double-buffered slab), across phases too.  Same arithmetic in the same order: the accumulators of the
    overlapped and the serial schedule must agree bit for bit -- and with the oracle.  The slab is shrunk to 32
    columns so that a phase spans many slabs; batch 16 puts the ML-only restart (which reads the previous slab's last
    column, src/mod_reservoir.f90:1034) on every second slab boundary."""
    monkeypatch.setenv("SML_TRAIN_SLAB", "32")
    region, bs, discard = 555, 16, 4
    w = region_weights(1152, region, m=450, ml_only=ml_only)
    phases = [_series(w, discard + 11 * bs + 5, 300 + p) for p in range(3)]
    rc = c_region(w)
    rc.train_init(bs)
    for td, im in phases:
        rc.train_phase(td, im, discard)
    grams = []
    # (overlap, state generation): the time loop inside one kernel per slab (k_train_stategen: the default for waves of
    # 64 regions and more, SML_TRAIN_STATEGEN=kernel forces it) or one launch per time step (k_train_update, =steps);
    # SML_TRAIN_SG_GROUP=6 selects the wide instantiation of the in-kernel loop
    # =ring: the same loop on the TMA ring of the spin-up kernel (k_train_stategen_ring, the default for waves >= 24)
    for on, stategen, group in ((True, "kernel", None), (False, "kernel", None), (False, "kernel", "6"), (True, "steps", None),
                                (False, "steps", None), (False, None, None), (False, "ring", None)):
        if group:
            monkeypatch.setenv("SML_TRAIN_SG_GROUP", group)
        else:
            monkeypatch.delenv("SML_TRAIN_SG_GROUP", raising=False)
        if stategen:
            monkeypatch.setenv("SML_TRAIN_STATEGEN", stategen)
        else:
            monkeypatch.delenv("SML_TRAIN_STATEGEN", raising=False)
        eng = single_region_engine(E, w)
        eng.train_set_overlap(on)
        eng.train_begin([region], bs)
        for td, im in phases:
            eng.train_feed([td], [im] if im is not None else None, discard)
        grams.append(eng.train_gram_get(region))
        st = eng.train_stats()
        if stategen and not on:
            assert eng.train_stategen_route() == stategen
        assert st["gram_ms"] > 0.0 and st["stategen_ms"] > 0.0
        assert eng.train_solve(1e-2, 1.0, True, 0.0)[0] == 0 if not ml_only else eng.train_solve(1e-2)[0] == 0
        eng.train_end()
        eng.close()
    monkeypatch.delenv("SML_TRAIN_STATEGEN", raising=False)
    for g in grams[1:]:
        assert np.array_equal(grams[0][0], g[0])
        assert np.array_equal(grams[0][1], g[1])
    assert np.array_equal(grams[0][0], grams[1][0])
    assert np.array_equal(grams[0][1], grams[1][1])
    assert rel_inf(grams[0][0], rc.sxs) < 1e-12
    assert rel_inf(grams[0][1], rc.sxt) < 1e-12


def test_training_with_dense_win_on_both_state_generation_routes(E, monkeypatch):
    """A W_in that is not one-non-zero-per-row takes the dense matmul(win, u) branch (src/mod_reservoir.f90:1113) in
    training too; both state-generation routes must agree bit for bit with each other and with the oracle to 1e-12."""
    region, bs, discard = 555, 6, 3
    w = region_weights(1152, region, m=450)
    rng = np.random.default_rng(8)
    win = w["win"].copy(order="F")
    win[::5, 2] += rng.standard_normal(win[::5, 2].shape)
    win[7, :] = rng.standard_normal(w["D"])
    w["win"] = win
    td, im = _series(w, discard + 4 * bs + 2, 21)
    rc = c_region(w)
    rc.train_init(bs)
    rc.train_phase(td, im, discard)
    grams = []
    for route in ("kernel", "steps"):
        monkeypatch.setenv("SML_TRAIN_STATEGEN", route)
        eng = single_region_engine(E, w)
        eng.train_begin([region], bs)
        eng.train_feed([td], [im], discard)
        grams.append(eng.train_gram_get(region))
        eng.train_end()
        eng.close()
    assert np.array_equal(grams[0][0], grams[1][0]) and np.array_equal(grams[0][1], grams[1][1])
    assert rel_inf(grams[0][0], rc.sxs) < 1e-12
    assert rel_inf(grams[0][1], rc.sxt) < 1e-12
