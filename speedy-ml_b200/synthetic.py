"""Synthetic SPEEDY-grid workloads (seeded; no external data) -- SURVEY.md section 8(d).

Host-side NumPy only.  The reservoir construction follows the reference's recipe so that the
structural facts the kernels exploit really hold:
  * adjacency: rounds of full permutations for rows and cols plus one partial round
    (makesparse, src/mod_linalg.f90:180-218; shuffle, src/mod_utilities.f90:1569-1596),
    values U[0,1) rescaled to the target spectral radius (gen_res, src/mod_reservoir.f90:182-212;
    power iteration stands in for ARPACK -- Perron root of a non-negative matrix);
  * W_in: exactly one non-zero per row, row j in block i=j//q, sigma*U(-1,1)
    (src/mod_reservoir.f90:262-283).
"""
from __future__ import annotations

import numpy as np

XGRID, YGRID, ZGRID = 96, 48, 8


def make_adjacency(n: int, k: int, rng: np.random.Generator, radius: float = 0.7, power_iters: int = 60):
    """COO (rows, cols 1-based int32; vals float64), duplicates allowed (they sum)."""
    vals = rng.random(k)
    rows = np.empty(k, dtype=np.int32)
    cols = np.empty(k, dtype=np.int32)
    if k > n:
        counter, leftover = divmod(k, n)
        for i in range(counter):
            rows[i * n:(i + 1) * n] = rng.permutation(n) + 1
            cols[i * n:(i + 1) * n] = rng.permutation(n) + 1
        if leftover:
            rows[counter * n:] = rng.permutation(n)[:leftover] + 1
            cols[counter * n:] = rng.permutation(n)[:leftover] + 1
    else:
        rows[:] = rng.permutation(n)[:k] + 1
        cols[:] = rng.permutation(n)[:k] + 1
    # spectral radius by power iteration (non-negative matrix -> Perron root)
    v = np.full(n, 1.0 / np.sqrt(n))
    lam = 1.0
    for _ in range(power_iters):
        w = np.bincount(rows - 1, weights=vals * v[cols - 1], minlength=n)
        lam = np.linalg.norm(w)
        if lam == 0.0:
            break
        v = w / lam
    if lam > 0.0:
        vals = (vals / lam) * radius
    return rows, cols, vals


def make_win_compact(n: int, D: int, rng: np.random.Generator, sigma: float = 0.5):
    """(values[n], column index 0-based int32[n]) of the one-per-row W_in"""
    q = n // D
    vals = sigma * (-1.0 + 2.0 * rng.random(n))
    col = (np.arange(n) // q).astype(np.int32)
    return vals, col


def win_dense_from_compact(vals, col, D):
    n = vals.size
    win = np.zeros((n, D), order="F")
    win[np.arange(n), col] = vals
    return win


def make_mean_std(L: int, rng: np.random.Generator):
    """plausible per-(variable,level) constants; slots after 32: logp, tisr, precip, sst"""
    mean = np.empty(L)
    std = np.empty(L)
    base_mean = [250.0, 5.0, 0.5, 3.0]   # T, U, V, q(g/kg)
    base_std = [12.0, 9.0, 6.0, 2.5]
    for v in range(4):
        for z in range(ZGRID):
            mean[v * ZGRID + z] = base_mean[v] * (1.0 + 0.01 * z) + 0.1 * rng.standard_normal()
            std[v * ZGRID + z] = base_std[v] * (1.0 + 0.02 * z) * (0.9 + 0.2 * rng.random())
    extra_mean = [0.0, 1.0e6, 0.5, 290.0]
    extra_std = [0.05, 4.0e5, 0.8, 6.0]
    for i in range(32, L):
        mean[i] = extra_mean[i - 32] * (1.0 + 0.01 * rng.standard_normal())
        std[i] = extra_std[i - 32] * (0.9 + 0.2 * rng.random())
    return mean, std


def smooth_field(shape, rng: np.random.Generator, nwaves: int = 8):
    """sum of low-wavenumber sinusoids on the 96x48 grid (leading dims broadcast)"""
    x = np.arange(XGRID)[:, None] * (2.0 * np.pi / XGRID)
    y = np.arange(YGRID)[None, :] * (np.pi / YGRID)
    lead = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    out = np.zeros((lead, XGRID, YGRID))
    for i in range(lead):
        for _ in range(nwaves):
            kx, ky = rng.integers(0, 5), rng.integers(0, 4)
            ph1, ph2 = rng.random(2) * 2 * np.pi
            out[i] += rng.standard_normal() * np.cos(kx * x + ph1) * np.cos(ky * y + ph2)
    out /= np.sqrt(nwaves)
    return out.reshape(shape)


def climatology(rng: np.random.Generator):
    """(clim4d (4,96,48,8) F-order physical units, clim2d logp, tisr grid, base sst, sea mask)"""
    base_mean = np.array([250.0, 5.0, 0.5, 3.0])
    base_std = np.array([12.0, 9.0, 6.0, 2.5])
    f = smooth_field((4, ZGRID, XGRID, YGRID), rng)
    clim4d = np.empty((4, XGRID, YGRID, ZGRID), order="F")
    for v in range(4):
        for z in range(ZGRID):
            clim4d[v, :, :, z] = base_mean[v] * (1 + 0.01 * z) + base_std[v] * f[v, z]
    clim4d[3] = np.abs(clim4d[3]) + 0.01
    clim2d = np.asfortranarray(0.05 * smooth_field((XGRID, YGRID), rng))
    tisr = np.asfortranarray(1.0e6 + 4.0e5 * smooth_field((XGRID, YGRID), rng))
    base_sst = np.asfortranarray(290.0 + 8.0 * smooth_field((XGRID, YGRID), rng))
    sea_mask = np.asfortranarray((smooth_field((XGRID, YGRID), rng) > 0.6).astype(np.float64))
    return clim4d, clim2d, tisr, base_sst, sea_mask


def ar1_series(D: int, T: int, rng: np.random.Generator, phi: float = 0.95):
    """standardised AR(1) inputs (D, T) F-order, unit variance"""
    out = np.empty((D, T), order="F")
    x = rng.standard_normal(D)
    s = np.sqrt(1.0 - phi * phi)
    for t in range(T):
        x = phi * x + s * rng.standard_normal(D)
        out[:, t] = x
    return out
