"""shared builders for the parity tests: one seeded synthetic model, fed identically to the C oracle,
the NumPy oracle and the CUDA engine."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle_c as oc  # noqa: E402
from oracle import oracle_np as on  # noqa: E402

syn = importlib.import_module("speedy-ml_b200.synthetic")

SEED0 = 20251018  # seed = 20251018 + region_id


def sst_input_mask(region: int) -> bool:
    """region_id mod 10 < 7 -> 'ocean' region with SST input"""
    return region % 10 < 7


def region_weights(num_regions, region, m, precip_bool=True, sst_bool=True, sst_bool_input=None,
                   ml_only=False, radius=0.7, sigma=0.5, wout_scale=None, deg=6.0, with_dense_win=True):
    """-> dict with dims + rows/cols/vals/win(dense)/winc/wcol/wout/mean/std for one region"""
    if sst_bool_input is None:
        sst_bool_input = sst_bool and sst_input_mask(region)
    rc = oc.Region(num_regions, region, m=m, deg=deg, precip_bool=precip_bool, sst_bool=sst_bool,
                   sst_bool_input=sst_bool_input, ml_only=ml_only)
    rng = np.random.default_rng(SEED0 + region)
    rows, cols, vals = syn.make_adjacency(rc.n, rc.k, rng, radius=radius)
    winc, wcol = syn.make_win_compact(rc.n, rc.D, rng, sigma=sigma)
    N = rc.n + rc.S
    scale = (1.0 / np.sqrt(N)) if wout_scale is None else wout_scale
    wout = np.asfortranarray(rng.standard_normal((rc.P, N)) * scale)
    mean, std = syn.make_mean_std(rc.L, rng)
    w = dict(num_regions=num_regions, region=region, m=m, deg=deg, precip_bool=precip_bool, sst_bool=sst_bool,
             sst_bool_input=sst_bool_input, ml_only=ml_only, n=rc.n, D=rc.D, P=rc.P, S=rc.S, k=rc.k, L=rc.L,
             rows=rows, cols=cols, vals=vals, winc=winc, wcol=wcol, wout=wout, mean=mean, std=std)
    if with_dense_win:
        w["win"] = syn.win_dense_from_compact(winc, wcol, rc.D)
    return w


def c_region(w) -> "oc.Region":
    r = oc.Region(w["num_regions"], w["region"], m=w["m"], deg=w["deg"], precip_bool=w["precip_bool"],
                  sst_bool=w["sst_bool"], sst_bool_input=w["sst_bool_input"], ml_only=w["ml_only"])
    r.set_weights(w["rows"], w["cols"], w["vals"], w.get("win"), w["wout"], w["mean"], w["std"])
    if w.get("win") is None:
        r.set_win_compact(w["winc"], w["wcol"])
    return r


def np_region(w) -> "on.Region":
    r = on.Region(w["num_regions"], w["region"], m=w["m"], deg=w["deg"], precip_bool=w["precip_bool"],
                  sst_bool=w["sst_bool"], sst_bool_input=w["sst_bool_input"], ml_only=w["ml_only"])
    assert (r.n, r.D, r.P, r.S, r.k, r.mean_std_length) == (w["n"], w["D"], w["P"], w["S"], w["k"], w["L"])
    r.rows, r.cols, r.vals = w["rows"], w["cols"], w["vals"]
    r.win, r.wout, r.mean, r.std = w["win"], w["wout"], w["mean"], w["std"]
    r.x = np.zeros(r.n)
    r.feedback = np.zeros(r.D)
    r.local_model = np.zeros(r.S)
    r.outvec = np.zeros(r.P)
    return r


def initial_grids(seed=7):
    rng = np.random.default_rng(seed)
    clim4d, clim2d, tisr, base_sst, sea_mask = syn.climatology(rng)
    return dict(clim4d=clim4d, clim2d=clim2d, tisr=tisr, base_sst=base_sst, sea_mask=sea_mask)


def rel_inf(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / den)


# ---- slab-ocean reservoirs (res%reservoir_special) --------------------------------------------
def ocean_weights(num_regions, region, m=4000, precip_bool=True, radius=0.9, sigma=0.6, wout_scale=None, deg=6.0,
                  with_dense_win=True, mean=None, std=None, hybrid=False):
    """weights of one ocean reservoir; mean/std default to a fresh atmosphere-style vector (grid_special is a
    copy of the atmosphere grid, so the SST slot is the atmosphere's sst_mean_std_idx).  hybrid: the predict_slab
    variant (ml_only_ocean = .False.) whose W_out has chunk_size_prediction extra "model" columns"""
    rc = oc.OceanRegion(num_regions, region, m=m, deg=deg, precip_bool=precip_bool, hybrid=hybrid)
    rng = np.random.default_rng(SEED0 + 100000 + region)
    rows, cols, vals = syn.make_adjacency(rc.n, rc.k, rng, radius=radius)
    winc, wcol = syn.make_win_compact(rc.n, rc.D, rng, sigma=sigma)
    scale = (1.0 / np.sqrt(rc.n)) if wout_scale is None else wout_scale
    wout = np.asfortranarray(rng.standard_normal((rc.P, rc.n + rc.S)) * scale)
    if mean is None:
        mean, std = syn.make_mean_std(rc.L, rng)
    w = dict(num_regions=num_regions, region=region, m=m, deg=deg, precip_bool=precip_bool, n=rc.n, D=rc.D, P=rc.P,
             S=rc.S, hybrid=hybrid, k=rc.k, L=rc.L, A=rc.A, rows=rows, cols=cols, vals=vals, winc=winc, wcol=wcol, wout=wout,
             mean=np.array(mean), std=np.array(std), sst_idx=rc.g.sst_mean_std_idx)
    if with_dense_win:
        w["win"] = syn.win_dense_from_compact(winc, wcol, rc.D)
    return w


def c_ocean(w) -> "oc.OceanRegion":
    r = oc.OceanRegion(w["num_regions"], w["region"], m=w["m"], deg=w["deg"], precip_bool=w["precip_bool"],
                       hybrid=w.get("hybrid", False))
    r.set_weights(w["rows"], w["cols"], w["vals"], w.get("win"), w["wout"], w["mean"], w["std"])
    if w.get("win") is None:
        r.set_win_compact(w["winc"], w["wcol"])
    return r


def np_ocean(w) -> "on.OceanRegion":
    r = on.OceanRegion(w["num_regions"], w["region"], m=w["m"], deg=w["deg"], precip_bool=w["precip_bool"])
    assert (r.n, r.D, r.P, r.k, r.mean_std_length, r.sst_idx) == (w["n"], w["D"], w["P"], w["k"], w["L"], w["sst_idx"])
    r.rows, r.cols, r.vals = w["rows"], w["cols"], w["vals"]
    r.win, r.wout, r.mean, r.std = w["win"], w["wout"], w["mean"], w["std"]
    r.feedback = np.zeros(r.D)
    r.outvec = np.zeros(r.P)
    return r
