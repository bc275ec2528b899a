"""Properties of the step plan and of the failure detection (-m gpu):

* a region's arithmetic does not depend on how the model is sharded: its outvec and state are BIT-identical whether
  its rank owns 144 or 72 regions (the persistent kernel's partials follow fixed row blocks, the classic kernel's
  fixed 720-row chunks) -- the property bench.py's grid_checksum asserts across 1/2/4/8 GPUs;
* the two fused step kernels agree within rounding on a whole shard;
* a non-finite outvec is caught by the grid assembly (sml_step_exchange_begin returns 1, SML_GRID_NONFINITE sticks),
  SPEEDY's own input bounds (src/ppo_iogrid.f90:562-577) raise their bits, and run_speedy travels with the forecast.
"""
import importlib
import os

import numpy as np
import pytest

from helpers import initial_grids, region_weights, rel_inf

pytestmark = pytest.mark.gpu

R = 1152


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


@pytest.fixture(scope="module")
def shard_weights():
    # rank 3 of 8 owns regions 432..575; ranks 6 and 7 of 16 own its two halves
    return {r: region_weights(R, r, m=1200, with_dense_win=False) for r in range(432, 576)}


def build(E, ws, irank, numprocs, kernel):
    old = os.environ.get("SML_STEP_KERNEL")
    os.environ["SML_STEP_KERNEL"] = kernel
    try:
        eng = E.Engine(number_of_regions=R, irank=irank, numprocs=numprocs, sst_prescribed=True)
        for r in eng.region_indices:
            w = ws[r]
            eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                              win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
        eng.finalize()
    finally:
        if old is None:
            os.environ.pop("SML_STEP_KERNEL", None)
        else:
            os.environ["SML_STEP_KERNEL"] = old
    return eng


def drive(eng, ws, steps=3):
    rng = np.random.default_rng(99)
    draws = {r: (rng.standard_normal(576), rng.standard_normal(132), 0.2 * rng.standard_normal(2000)) for r in range(432, 576)}
    for r in eng.region_indices:
        fb, lm, x0 = draws[r]
        w = ws[r]
        eng.feedback_set(r, fb[:w["D"]])
        eng.local_model_set(r, lm[:w["S"]])
        eng.state_set(r, x0[:w["n"]])
    for _ in range(steps):
        eng.predict()
    return {r: (eng.outvec_get(r), eng.state_get(r)) for r in eng.region_indices}


@pytest.mark.parametrize("kernel", ["persist", "classic"])
def test_results_do_not_depend_on_the_sharding(E, shard_weights, kernel):
    ws = shard_weights
    big = build(E, ws, 3, 8, kernel)
    assert big.region_indices == list(range(432, 576))
    a = drive(big, ws)
    plan_big = big.step_plan()
    big.close()
    for irank in (6, 7):
        small = build(E, ws, irank, 16, kernel)
        b = drive(small, ws)
        plan_small = small.step_plan()
        small.close()
        assert plan_small["part_rows"] == plan_big["part_rows"]
        for r, (ov, x) in b.items():
            assert np.array_equal(ov, a[r][0]), f"outvec of region {r} differs between 144- and 72-region shards"
            assert np.array_equal(x, a[r][1])


def test_step_kernels_agree_on_a_shard(E, shard_weights):
    ws = shard_weights
    e1 = build(E, ws, 3, 8, "persist")
    e2 = build(E, ws, 3, 8, "classic")
    assert e1.step_plan()["kernel"] == "k_step_persist" and e2.step_plan()["kernel"] == "k_step"
    a, b = drive(e1, ws), drive(e2, ws)
    e1.close()
    e2.close()
    for r in a:
        assert rel_inf(a[r][0], b[r][0]) < 1e-12
        assert np.array_equal(a[r][1], b[r][1])          # the state update is the same arithmetic in both


def test_nonfinite_and_range_detection(E):
    ws = {r: region_weights(R, r, m=300, with_dense_win=False) for r in range(R)}
    eng = E.Engine(number_of_regions=R, sst_prescribed=True)
    for r, w in ws.items():
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    G = initial_grids()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    eng.predict()
    grids = eng.step_exchange_begin(1)
    assert not eng.grid_nonfinite and np.isfinite(grids[0]).all()
    assert eng.grid_status() & E.GRID_NONFINITE == 0
    eng.step_exchange_end(1, G["clim4d"], G["clim2d"], G["tisr"])
    # one region's readout goes bad: the assembled grid carries the NaN and the assembly says so
    bad = ws[700]["wout"].copy(order="F")
    bad[5, 17] = np.nan
    eng.wout_set(700, bad)
    eng.predict()
    grids = eng.step_exchange_begin(2)
    assert eng.grid_nonfinite
    assert not np.isfinite(grids[0]).all()
    assert eng.grid_status() & E.GRID_NONFINITE
    eng.step_exchange_end(2, G["clim4d"], G["clim2d"], G["tisr"])
    # sticky until reset; after repairing the weights, clearing what the NaN has reached through the exchange (the
    # neighbours' feedback and states) and resetting the status it stays clear
    eng.wout_set(700, ws[700]["wout"])
    for r, w in ws.items():
        eng.state_set(r, np.zeros(w["n"]))
        eng.feedback_set(r, np.zeros(w["D"]))
        eng.local_model_set(r, np.zeros(w["S"]))
    assert eng.grid_status() & E.GRID_NONFINITE       # still raised: sticky
    eng.grid_status_reset()
    eng.predict()
    eng.step_exchange_begin(3)
    assert not eng.grid_nonfinite and eng.grid_status() & E.GRID_NONFINITE == 0
    eng.step_exchange_end(3, G["clim4d"], G["clim2d"], G["tisr"])
    # SPEEDY's bounds (src/ppo_iogrid.f90:562-577): with a zero readout every outvec is the region's mean vector, all
    # inside the bounds; then one temperature (output element 0 = var 1 at the region's first cell) is pushed out
    for r, w in ws.items():
        eng.wout_set(r, np.zeros_like(w["wout"]))
    eng.grid_status_reset()
    eng.predict()
    eng.step_exchange_begin(4)
    assert eng.grid_status() == 0
    eng.step_exchange_end(4, G["clim4d"], G["clim2d"], G["tisr"])
    eng.predict()
    ov = eng.outvec_get(10)
    ov[0] = 500.0
    eng.outvec_set(10, ov)
    ov = eng.outvec_get(11)
    ov[1] = -200.0                      # var 2 = u at the first cell
    eng.outvec_set(11, ov)
    eng.step_exchange_begin(5)
    assert eng.grid_status() == (E.GRID_T_RANGE | E.GRID_U_RANGE)
    assert not eng.grid_nonfinite
    eng.step_exchange_end(5, G["clim4d"], G["clim2d"], G["tisr"])
    # run_speedy: on a single rank it is what the host set
    assert eng.run_speedy() is True
    eng.set_run_speedy(False)
    assert eng.run_speedy() is False
    eng.close()
