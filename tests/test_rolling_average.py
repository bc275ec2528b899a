"""rolling_average_over_a_period_2d (src/mod_utilities.f90:1773-1815): the time smoothing of the slab-ocean reservoirs'
atmosphere inputs (src/mod_slab_ocean_reservoir.f90:398, :452).  The reference holds no test for it and the worked example
in the subroutine's comment contradicts the code (it assumes a window of `period` values; the code sums period+1), so the
known answers below are derived by hand from the code as written.  CPU: NumPy oracle; -m gpu: engine vs oracle, exact."""
import importlib

import numpy as np
import pytest

from helpers import oc, on


def test_oracle_matches_the_code_as_written():
    # the comment's own input, period 6: t <= 6 -> mean of the first t values; then 7 values summed, divided by 6
    g = np.array([[1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0]], dtype=np.float64)
    out = on.rolling_average_over_a_period_2d(g, 6)
    want = np.array([1, 1 / 2, 1 / 3, 1 / 4, 1 / 5, 1 / 6, 1 / 6, 1 / 6, 2 / 6, 2 / 6, 2 / 6, 2 / 6, 2 / 6, 2 / 6, 1 / 6])
    assert np.array_equal(out[0], want)
    # a window whose sum is (numerically) zero keeps the input value -- only in the 2-D variant
    z = np.array([[3.0, -3.0, 0.0, 0.0, 5e-8, 0.0]])
    o2 = on.rolling_average_over_a_period_2d(z, 2)
    o3 = on.rolling_average_over_a_period_2d(z, 2, keep_small=False)
    assert np.array_equal(o2[0], [3.0, 0.0, 0.0, -1.5, 5e-8, 0.0])      # windows 3..5 and 4..6 sum to 5e-8: input kept
    assert np.array_equal(o3[0], [3.0, 0.0, 0.0, -1.5, 2.5e-8, 2.5e-8])
    # head of the series: plain mean of what has been seen
    r = np.random.default_rng(0).standard_normal((5, 40))
    o = on.rolling_average_over_a_period_2d(r, 28)
    for t in range(28):
        assert np.allclose(o[:, t], r[:, :t + 1].mean(axis=1), rtol=1e-14, atol=0)
    assert np.allclose(o[:, 33], r[:, 5:34].sum(axis=1) / 28, rtol=1e-14, atol=0)


def test_c_and_numpy_restatements_agree_bitwise():
    rng = np.random.default_rng(3)
    for rows, T, period in ((7, 50, 6), (80, 300, 28), (3, 10, 168), (5, 40, 1)):
        g = np.asfortranarray(rng.standard_normal((rows, T)))
        g[0, 5:5 + 2 * period + 2] = 0.0
        g[1 % rows, 3:9] = 1e-9
        for keep_small in (True, False):
            assert np.array_equal(oc.rolling_average_over_a_period_2d(g, period, keep_small),
                                  on.rolling_average_over_a_period_2d(g, period, keep_small))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,row0,nrows,T,period", [(128, 0, 80, 400, 28), (37, 5, 20, 60, 168), (16, 0, 16, 9, 1), (9, 8, 1, 300, 7)])
def test_engine_rolling_average_is_exact(rows, row0, nrows, T, period):
    E = importlib.import_module("speedy-ml_b200.engine")
    rng = np.random.default_rng(rows + T)
    g = np.asfortranarray(rng.standard_normal((rows, T)))
    g[row0, 10:10 + 3 * period] = 0.0        # an all-zero stretch: the |sum| <= 1e-7 branch
    if nrows > 1:
        g[row0 + 1, 20:25] = 1e-9
    eng = E.Engine(number_of_regions=1152, irank=0, numprocs=1152)
    for keep_small in (True, False):
        mine = g.copy(order="F")
        eng.rolling_average_over_a_period_2d(mine, period, row0=row0, nrows=nrows, keep_small=keep_small)
        want = g.copy(order="F")
        want[row0:row0 + nrows] = on.rolling_average_over_a_period_2d(g[row0:row0 + nrows], period, keep_small=keep_small)
        assert np.array_equal(mine, want)    # rows outside [row0, row0+nrows) untouched, the rest bit for bit
    eng.close()
