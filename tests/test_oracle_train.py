"""Training accumulation + ridge solve: C oracle (statement-by-statement, batch ring buffer) vs NumPy
(plain recurrence + batched Gram)."""
import numpy as np

from helpers import c_region, np_region, oc, on, region_weights, rel_inf, syn


def _series(w, T, seed):
    rng = np.random.default_rng(seed)
    td = syn.ar1_series(w["D"], T, rng)
    im = np.asfortranarray(rng.standard_normal((w["S"], T)))
    return td, im


def test_target_rows_are_region_interior():
    w = region_weights(1152, 24 * 3, m=450)   # south-pole region: 4x3 halo, tdata y 1..2
    rn = np_region(w)
    rc = c_region(w)
    rng = np.random.default_rng(1)
    sv = np.asfortranarray(rng.standard_normal((w["D"], 5)))
    tn = on.tile_full_input_to_target_data(rn, sv)
    tc = np.zeros((w["P"], 5), order="F")
    oc.lib().orc_tile_full_input_to_target_data2d(rc.g, rc.d, oc._d(sv), w["D"], 5, oc._d(tc))
    assert tn.shape == (136, 5)
    assert np.array_equal(tn, tc)


def test_hybrid_training_c_vs_numpy():
    w = region_weights(1152, 555, m=450)       # n = 576
    rc, rn = c_region(w), np_region(w)
    bs, discard = 7, 5
    phases = [_series(w, discard + 4 * bs, 40 + p) for p in range(2)]
    rc.train_init(bs)
    for td, im in phases:
        rc.train_phase(td, im, discard)
    sxs_c, sxt_c = rc.sxs.copy(), rc.sxt.copy()
    info = rc.fit(beta_res=1e-3, beta_model=1.0, using_prior=True, prior_val=0.0)
    assert info == 0
    wout_n, sxs_n, sxt_n, info_n = on.train_hybrid(rn, phases, bs, discard, 1e-3, 1.0)
    assert info_n == 0
    N = w["n"] + w["S"]
    d = np.arange(N)
    sxs_n_noreg = sxs_n.copy()
    sxs_n_noreg[d[:w["S"]], d[:w["S"]]] -= 1.0
    sxs_n_noreg[d[w["S"]:], d[w["S"]:]] -= 1e-6
    assert rel_inf(sxs_c, sxs_n_noreg) < 1e-12
    assert rel_inf(sxt_c, sxt_n) < 1e-12
    # W_out: ill-conditioned ridge -> judge by residual (SURVEY.md 8c tolerances)
    A = sxs_n.T
    for wout in (rc.wout, wout_n):
        X = wout.T
        res = np.linalg.norm(A @ X - sxt_n.T) / (np.linalg.norm(A) * np.linalg.norm(X) + np.linalg.norm(sxt_n))
        assert res < 1e-13


def test_ml_only_training_restart_quirk():
    # ML-only paths restart each batch from states(:,batch_size) AFTER it was squared
    # (src/mod_reservoir.f90:1027,1034) -- the C oracle reproduces it; check it differs from the plain
    # recurrence and that the first batch agrees with it.
    w = region_weights(1152, 555, m=450, ml_only=True)
    rc = c_region(w)
    bs, discard = 5, 3
    td, _ = _series(w, discard + 3 * bs, 9)
    rc.train_init(bs)
    rc.train_phase(td, None, discard)
    rn = np_region(w)
    x = np.zeros(rn.n)
    for i in range(discard):
        x = on.state_update(rn, x, td[:, i])
    states = [x]
    for s in range(1, bs):
        x = on.state_update(rn, x, td[:, discard + s - 1])
        states.append(x)
    S1 = np.array(states).T
    S1[1::2] **= 2
    first_batch = S1 @ S1.T
    x = S1[:, -1]    # squared copy is the restart state
    states = []
    for s in range(bs, 2 * bs):
        x = on.state_update(rn, x, td[:, discard + s - 1])
        states.append(x)
    S2 = np.array(states).T
    S2[1::2] **= 2
    x = S2[:, -1]
    states = []
    for s in range(2 * bs, 3 * bs):
        x = on.state_update(rn, x, td[:, discard + s - 1])
        states.append(x)
    S3 = np.array(states).T
    S3[1::2] **= 2
    expect = first_batch + S2 @ S2.T + S3 @ S3.T
    assert rel_inf(rc.sxs, expect) < 1e-12
