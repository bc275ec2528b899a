"""The one-launch spin-up (k_sync_persist) against the step-per-launch update kernels and the CPU oracle (-m gpu).

synchronize (src/mod_reservoir.f90:1354-1381; slab twin src/mod_slab_ocean_reservoir.f90:1237-1266) advances every
region T times with no interaction between regions, so the engine runs a region's whole time loop inside one CTA
(state in shared memory, adjacency re-read from L2).  Per-row arithmetic and order are unchanged, hence

* states BIT-identical to T launches of the update kernel (SML_SYNC_KERNEL=steps), for every ring geometry
  (consumer groups, ring depth, tile rows; fewer tiles than slots; several regions per CTA; odd and even T; T = 1);
* within the usual tolerance of the oracle.
"""
import importlib
import os

import numpy as np
import pytest

from helpers import c_ocean, c_region, ocean_weights, region_weights, rel_inf, syn

pytestmark = pytest.mark.gpu

R = 1152
TOL_STEP = 1e-13


@pytest.fixture(scope="module")
def E():
    return importlib.import_module("speedy-ml_b200.engine")


class env:
    def __init__(self, **kv):
        self.kv = {k: (None if v is None else str(v)) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def one_region_engine(E, w):
    eng = E.Engine(number_of_regions=w["num_regions"], irank=w["region"], numprocs=w["num_regions"])
    eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                      win_compact=w["winc"], win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    return eng


@pytest.mark.parametrize("T", [1, 2, 7, 56])
def test_full_size_region_bitwise_and_oracle(E, T):
    region = 555
    w = region_weights(R, region, m=6000, with_dense_win=False)
    rng = np.random.default_rng(100 + T)
    series = syn.ar1_series(w["D"], T, rng)
    x0 = 0.2 * rng.standard_normal(w["n"])
    rc = c_region(w)
    rc.x[:] = x0
    rc.synchronize(series, T)
    eng = one_region_engine(E, w)
    got = {}
    for name in ("steps", None):
        with env(SML_SYNC_KERNEL=name):
            eng.state_set(region, x0)
            n0 = eng.kernel_launch_count()
            eng.synchronize(region, series)
            got[name] = (eng.state_get(region), eng.kernel_launch_count() - n0)
    eng.close()
    assert got[None][1] <= 2, "the persistent spin-up is ONE launch (+ one for the tile-major pack the first time)"
    assert got["steps"][1] >= T
    assert np.array_equal(got[None][0], got["steps"][0])
    assert rel_inf(got[None][0], rc.x) < TOL_STEP * max(10, T)


GEOMETRIES = [dict(SML_SYNC_GROUPS=1, SML_SYNC_STAGES=2), dict(SML_SYNC_GROUPS=1, SML_SYNC_STAGES=5, SML_SYNC_TILE_ROWS=96),
              dict(SML_SYNC_GROUPS=2, SML_SYNC_STAGES=4), dict(SML_SYNC_GROUPS=2, SML_SYNC_STAGES=8, SML_SYNC_TILE_ROWS=64),
              dict(SML_SYNC_GROUPS=3, SML_SYNC_STAGES=6, SML_SYNC_TILE_ROWS=160), dict(SML_SYNC_GROUPS=4, SML_SYNC_STAGES=4),
              dict(SML_SYNC_GROUPS=2, SML_SYNC_STAGES=12, SML_SYNC_TILE_ROWS=480),
              dict(SML_SYNC_GENERIC=1), dict(SML_SYNC_GENERIC=1, SML_SYNC_GROUPS=3, SML_SYNC_STAGES=3, SML_SYNC_TILE_ROWS=128)]


@pytest.fixture(scope="module")
def shard():
    # a polar row (clipped halo: smaller D) and interior rows, SST and non-SST regions: sizes differ inside the shard
    regions = list(range(0, 6)) + list(range(552, 560))
    return {r: region_weights(R, r, m=1200, with_dense_win=False) for r in regions}


def shard_engine(E, ws):
    """one rank that owns exactly the listed regions is not a real decomposition; use one engine per contiguous run"""
    engs = []
    for lo, hi in ((0, 6), (552, 560)):
        # numprocs such that rank irank owns [lo, hi): contiguous shards of equal size
        size = hi - lo
        assert R % size == 0 and lo % size == 0
        eng = E.Engine(number_of_regions=R, irank=lo // size, numprocs=R // size)
        assert eng.region_indices == list(range(lo, hi))
        for r in eng.region_indices:
            w = ws[r]
            eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                              win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
        eng.finalize()
        engs.append(eng)
    return engs


def test_shard_every_geometry_bitwise(E, shard):
    ws = shard
    T = 9
    rng = np.random.default_rng(5)
    series = {r: syn.ar1_series(ws[r]["D"], T, rng) for r in ws}
    x0 = {r: 0.2 * rng.standard_normal(ws[r]["n"]) for r in ws}
    want = {}
    for r in ws:
        rc = c_region(ws[r])
        rc.x[:] = x0[r]
        rc.synchronize(series[r], T)
        want[r] = rc.x.copy()
    for eng in shard_engine(E, ws):
        def run(**kv):
            with env(**kv):
                for r in eng.region_indices:
                    eng.state_set(r, x0[r])
                n0 = eng.kernel_launch_count()
                eng.synchronize_all([series[r] for r in eng.region_indices], T)
                return {r: eng.state_get(r) for r in eng.region_indices}, eng.kernel_launch_count() - n0
        base, nl = run(SML_SYNC_KERNEL="steps")
        assert nl >= T
        for r in eng.region_indices:
            assert rel_inf(base[r], want[r]) < TOL_STEP * 10
        for geo in GEOMETRIES:
            for ctas in (None, 1, 3):     # 1 and 3: several regions per CTA, the barriers' phases run on across regions
                got, nl = run(SML_SYNC_KERNEL=None, SML_SYNC_CTAS=ctas, **geo)
                assert nl <= 2, (geo, ctas)     # the spin-up itself + the re-pack when the tile geometry changes
                for r in eng.region_indices:
                    assert np.array_equal(got[r], base[r]), (geo, ctas, r)
        # one region of the shard: the others keep their state
        r1 = eng.region_indices[2]
        for r in eng.region_indices:
            eng.state_set(r, x0[r])
        eng.synchronize(r1, series[r1])
        assert np.array_equal(eng.state_get(r1), base[r1])
        for r in eng.region_indices:
            if r != r1:
                assert np.array_equal(eng.state_get(r), x0[r])
        eng.close()


@pytest.mark.parametrize("deg", [3.0, 12.0, 24.0])
def test_wide_and_narrow_rows_bitwise(E, deg):
    """adjacency degrees 3 / 12 / 24 (BASELINE config 5): odd widths, several column slabs, the any-width instantiation"""
    region = 555
    w = region_weights(R, region, m=2000, deg=deg, with_dense_win=False)
    rng = np.random.default_rng(int(deg))
    T = 5
    series = syn.ar1_series(w["D"], T, rng)
    x0 = 0.2 * rng.standard_normal(w["n"])
    rc = c_region(w)
    rc.x[:] = x0
    rc.synchronize(series, T)
    eng = one_region_engine(E, w)
    got = {}
    for name in ("steps", None):
        with env(SML_SYNC_KERNEL=name):
            eng.state_set(region, x0)
            eng.synchronize(region, series)
            got[name] = eng.state_get(region)
    eng.close()
    assert np.array_equal(got[None], got["steps"])
    assert rel_inf(got[None], rc.x) < TOL_STEP * 10


def test_large_reservoir_keeps_the_step_launches(E):
    """m = 12000: two state vectors take 190 KB of shared memory, the ring that is left would feed ~200 consumer threads --
    slower than one launch per step -- so the engine falls back (and still matches the oracle)"""
    region = 555
    w = region_weights(R, region, m=12000, with_dense_win=False)
    rng = np.random.default_rng(12)
    T = 4
    series = syn.ar1_series(w["D"], T, rng)
    rc = c_region(w)
    rc.synchronize(series, T)
    eng = one_region_engine(E, w)
    n0 = eng.kernel_launch_count()
    eng.synchronize(region, series)
    assert eng.kernel_launch_count() - n0 >= T
    assert rel_inf(eng.state_get(region), rc.x) < TOL_STEP * 10
    eng.close()


def test_ocean_spin_up_bitwise(E):
    """the slab-ocean reservoirs' synchronize goes through the same kernel (kind = OCEAN)"""
    region = 700
    wa = region_weights(R, region, m=300, sst_bool_input=True, with_dense_win=False)
    wo = ocean_weights(R, region, m=1000, with_dense_win=False, mean=wa["mean"], std=wa["std"])
    eng = E.Engine(number_of_regions=R, irank=region, numprocs=R)
    eng.region_upload(region, wa["rows"], wa["cols"], wa["vals"], wa["wout"], wa["mean"], wa["std"], win_compact=wa["winc"],
                      win_col=wa["wcol"], D=wa["D"], sst_bool_input=wa["sst_bool_input"])
    eng.region_upload(region, wo["rows"], wo["cols"], wo["vals"], wo["wout"], wo["mean"], wo["std"], win_compact=wo["winc"],
                      win_col=wo["wcol"], D=wo["D"], kind=E.OCEAN, sst_mean=wo["mean"][wo["sst_idx"] - 1],
                      sst_std=wo["std"][wo["sst_idx"] - 1])
    eng.finalize()
    rng = np.random.default_rng(8)
    T = 12
    series = syn.ar1_series(wo["D"], T, rng)
    co = c_ocean(wo)
    co.synchronize(series, T)
    got = {}
    for name in ("steps", None):
        with env(SML_SYNC_KERNEL=name):
            eng.state_set(region, np.zeros(wo["n"]), kind=E.OCEAN)
            eng.synchronize(region, series, kind=E.OCEAN)
            got[name] = eng.state_get(region, kind=E.OCEAN)
    eng.close()
    assert np.array_equal(got[None], got["steps"])
    assert rel_inf(got[None], co.x) < TOL_STEP * 100
