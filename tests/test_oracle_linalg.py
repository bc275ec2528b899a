"""mod_linalg restatements: known answers + LAPACK cross-checks."""
import numpy as np
from scipy.linalg import lapack

from helpers import oc, on, rel_inf


def test_reference_known_answer_pinv_diag():
    # tests/mod_unit_test.f90:16-47: pinv(diag(1..10)) == diag(1/i), threshold 1e-10 on the summed difference
    A = np.diag(np.arange(1.0, 11.0))
    Ainv = on.pinv_svd(A)
    realinv = np.diag(1.0 / np.arange(1.0, 11.0))
    assert abs(np.sum(Ainv - realinv)) < 1e-10
    assert np.max(np.abs(Ainv - realinv)) < 1e-14


def test_pinv_threshold_zeroes_small_singular_values():
    # thres = 1e-2 (src/mod_linalg.f90:44,77-81)
    A = np.diag([1.0, 0.5, 0.005])
    assert np.allclose(on.pinv_svd(A), np.diag([1.0, 2.0, 0.0]))


def test_coo_duplicates_sum_and_one_based():
    rows = np.array([1, 1, 3, 3, 2], dtype=np.int32)
    cols = np.array([2, 2, 1, 3, 2], dtype=np.int32)
    vals = np.array([0.5, 0.25, 2.0, 1.0, -1.0])
    x = np.array([1.0, 2.0, 3.0])
    y = oc.coo_mv(3, rows, cols, vals, x)
    assert np.array_equal(y, np.array([1.5, -2.0, 5.0]))


def test_mldivide_vs_lapack_dgesv():
    rng = np.random.default_rng(0)
    for n, k in ((1, 1), (7, 3), (150, 20), (333, 136)):
        A = rng.standard_normal((n, n)) + 0.1 * n * np.eye(n)
        B = rng.standard_normal((n, k))
        X, info = oc.mldivide(A, B)
        _, _, Xl, infol = lapack.dgesv(A, B)
        assert info == 0 and infol == 0
        assert rel_inf(X, Xl) < 1e-11
        assert np.linalg.norm(A @ X - B) / (np.linalg.norm(A) * np.linalg.norm(X) + np.linalg.norm(B)) < 1e-14


def test_mldivide_needs_pivoting():
    A = np.array([[0.0, 2.0], [3.0, 1.0]])
    B = np.array([[4.0], [5.0]])
    X, info = oc.mldivide(A, B)
    assert info == 0
    assert np.allclose(A @ X, B)


def test_mldivide_singular_reports_info_like_dgesv():
    # reference prints and continues (src/mod_linalg.f90:147-150)
    A = np.array([[1.0, 2.0], [2.0, 4.0]])
    B = np.ones((2, 1))
    _, info = oc.mldivide(A, B)
    assert info == 2


def test_mldivide_shape_mismatch_returns_unchanged():
    A = np.eye(3)
    B = np.ones((2, 1))
    X, info = oc.mldivide(A, B)
    assert info == -1 and np.array_equal(X, B)
