"""Reservoir construction restated twice (C and NumPy) with the counter-based generator the engine shares: makesparse
(src/mod_linalg.f90:180-218), the k-shuffle (src/mod_utilities.f90:1569-1596) and the W_in build
(src/mod_reservoir.f90:262-283).  Structure facts of the reference's recipe plus bitwise agreement of the two."""
import numpy as np

from helpers import oc, on


def test_shuffle_is_a_permutation_prefix():
    for n, size in ((50, 50), (50, 17), (288, 288), (1000, 1)):
        s = oc.shuffle(n, size, 99, 7, 3)
        assert len(set(s.tolist())) == size and s.min() >= 1 and s.max() <= n
        assert np.array_equal(s, on.shuffle(n, size, 99, 7, 3))


def test_makesparse_structure_and_two_restatements():
    n, k = 288, 1537                      # 5 full rounds + 97 left over
    for region in (0, 555):
        r, c, v = oc.makesparse(n, k, 20251018, region)
        r2, c2, v2 = on.makesparse(n, k, 20251018, region)
        assert np.array_equal(r, r2) and np.array_equal(c, c2) and np.array_equal(v, v2)
        for i in range(k // n):           # every round visits every row and every column exactly once
            assert sorted(r[i * n:(i + 1) * n].tolist()) == list(range(1, n + 1))
            assert sorted(c[i * n:(i + 1) * n].tolist()) == list(range(1, n + 1))
        assert len(set(r[5 * n:].tolist())) == k % n          # the partial round: distinct rows
        deg = np.bincount(r - 1, minlength=n)
        assert deg.min() == 5 and deg.max() == 6              # row degree 5 or 6
        assert v.min() >= 0.0 and v.max() < 1.0
    a, b = oc.makesparse(n, k, 1, 3), oc.makesparse(n, k, 1, 4)
    assert not np.array_equal(a[0], b[0])                      # regions draw from different streams
    r, c, v = oc.makesparse(n, 100, 5, 9)                      # k <= n: one partial shuffle each
    assert len(set(r.tolist())) == 100 and np.array_equal(r, on.makesparse(n, 100, 5, 9)[0])


def test_gen_win_one_per_row_blocks():
    n, D, sigma = 288, 72, 0.5
    w, col = oc.gen_win(n, D, sigma, 11, 42)
    w2, col2 = on.gen_win(n, D, sigma, 11, 42)
    assert np.array_equal(w, w2) and np.array_equal(col, col2)
    assert np.array_equal(col, np.arange(n) // (n // D))
    assert np.abs(w).max() <= sigma and np.abs(w).mean() > 0.1
