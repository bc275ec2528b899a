"""gen_res on the device: batched power iteration for the spectral radius of every local
adjacency (replaces ARPACK sparse_eigen) and the rescale to the target radius.  -m gpu."""
import importlib

import numpy as np
import pytest

from helpers import c_region, on, region_weights, rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


def _unscaled(w, seed):
    rng = np.random.default_rng(seed)
    w = dict(w)
    w["vals"] = rng.random(w["vals"].size)           # makesparse: vals ~ U[0,1)  (src/mod_linalg.f90:191)
    return w


def test_sparse_eigen_matches_dense_eigenvalues_and_rescale(E):
    regions = [0, 1, 2, 3]                              # one rank of a 288-rank layout owns regions 0..3
    ws = {r: _unscaled(region_weights(1152, r, m=450, with_dense_win=False), 900 + r) for r in regions}
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=288)
    for r in regions:
        w = ws[r]
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    eigs, iters, ok = eng.sparse_eigen(maxit=600, tol=1e-14)
    assert ok and iters <= 600
    want = np.array([on.sparse_eigen(ws[r]["n"], ws[r]["rows"], ws[r]["cols"], ws[r]["vals"]) for r in regions])
    assert rel_inf(eigs, want) < 1e-10
    # deterministic: fixed-order reductions
    eigs2, _, _ = eng.sparse_eigen(maxit=600, tol=1e-14)
    assert np.array_equal(eigs, eigs2)
    # gen_res: vals <- vals/eig*radius on the device == the oracle stepping with host-scaled vals
    radius = 0.7
    factor = radius / eigs
    eng.adjacency_scale(factor)
    rng = np.random.default_rng(3)
    start = {}
    for r in regions:
        w = ws[r]
        start[r] = (0.3 * rng.standard_normal(w["n"]), rng.standard_normal(w["D"]), rng.standard_normal(w["S"]))
        eng.state_set(r, start[r][0])
        eng.feedback_set(r, start[r][1])
        eng.local_model_set(r, start[r][2])
    eng.predict()
    for i, r in enumerate(regions):
        w = dict(ws[r])
        w["vals"] = on.gen_res_scale(w["vals"], eigs[i], radius)
        rc = c_region(w)
        rc.x[:], rc.feedback[:], rc.local_model[:] = start[r]
        rc.predict()
        assert rel_inf(eng.state_get(r), rc.x) < 1e-12
        assert rel_inf(eng.outvec_get(r), rc.outvec) < 1e-12
    # the scaled matrix has the requested spectral radius
    eigs3, _, ok3 = eng.sparse_eigen(maxit=600, tol=1e-14)
    assert ok3 and np.allclose(eigs3, radius, rtol=1e-12, atol=0)
    eng.close()


def test_sparse_eigen_reports_non_convergence(E):
    w = _unscaled(region_weights(1152, 555, m=450, with_dense_win=False), 1)
    eng = E.Engine(number_of_regions=1152, irank=555, numprocs=1152)
    eng.region_upload(555, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                      win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    eigs, iters, ok = eng.sparse_eigen(maxit=2, tol=1e-15)
    assert not ok and iters == 2 and eigs[0] > 0
    with pytest.raises(E.EngineError):
        eng.gen_res(0.7, maxit=2, tol=1e-15)
    eng.close()


# ---- makesparse + W_in on the device (round 2): structure bit-exact against the restated counter streams ------------
@pytest.mark.parametrize("region,m", [(555, 600), (0, 2000), (1151, 6000)])
def test_device_makesparse_and_win_match_the_oracle_streams(E, region, m):
    from helpers import oc
    w = region_weights(1152, region, m=m, with_dense_win=False)
    n, k, D = w["n"], w["k"], w["D"]
    seed, sigma, radius = 20251018 + 7 * region, 0.5, 0.7
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    eng.region_generate(region, n, k, D, w["P"], w["S"], w["mean"], w["std"], seed, sigma, wout=w["wout"],
                        sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    rows, cols, vals = eng.region_coo_get(region)
    winc, wcol = eng.region_win_get(region)
    r0, c0, v0 = oc.makesparse(n, k, seed, region)
    wi0, wc0 = oc.gen_win(n, D, sigma, seed, region)
    assert np.array_equal(rows, r0) and np.array_equal(cols, c0)        # structure: bit-exact
    assert np.array_equal(vals, v0)                                      # unscaled values: the same 53-bit draws
    assert np.array_equal(winc, wi0) and np.array_equal(wcol, wc0)
    deg = np.bincount(rows - 1, minlength=n)
    assert deg.min() >= k // n and deg.max() <= k // n + 1
    # gen_res: spectral radius by power iteration, rescale on the device (ELL and the COO copy alike)
    factor, eigs = eng.gen_res(radius)
    rows2, cols2, vals2 = eng.region_coo_get(region)
    assert np.array_equal(rows2, rows) and np.array_equal(cols2, cols)
    assert rel_inf(vals2, v0 * factor[0]) < 1e-15
    # the constructed reservoir steps exactly like an uploaded one with the same arrays (oracle twin)
    w2 = dict(w, rows=rows2, cols=cols2, vals=vals2, winc=winc, wcol=wcol, win=None)
    rc = c_region(w2)
    rng = np.random.default_rng(3)
    fb, lm = rng.standard_normal(D), rng.standard_normal(w["S"])
    for _ in range(3):
        rc.feedback[:], rc.local_model[:] = fb, lm
        eng.feedback_set(region, fb)
        eng.local_model_set(region, lm)
        rc.predict()
        eng.predict()
    assert rel_inf(eng.state_get(region), rc.x) < 1e-12
    assert rel_inf(eng.outvec_get(region), rc.outvec) < 1e-12
    if n <= 600:
        import scipy.sparse as sp
        A = sp.coo_matrix((vals2, (rows2 - 1, cols2 - 1)), shape=(n, n)).toarray()
        assert abs(np.max(np.abs(np.linalg.eigvals(A))) - radius) < 1e-9
    eng.close()
