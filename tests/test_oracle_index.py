"""Index maps: C oracle vs NumPy oracle (bit-exact) + the reference's own known answers."""
from collections import Counter

import numpy as np
import pytest

from helpers import oc, on


def test_reference_known_answer_getxyresextent():
    # tests/mod_unit_test.f90:63-96: 288 regions, region 145 -> 4x4 chunks, x 49..52.
    # Its y expectation (9..12) is stale against the CURRENT getworkerlower_leftcorner
    # (src/res_domain.f90:282-292), which gives y 5..8; x and chunk are the pin.
    for mod in (oc, on):
        xs, xe, ys, ye, xc, yc = mod.getxyresextent(288, 145)
        assert (xc, yc) == (4, 4)
        assert (xs, xe) == (49, 52)
        assert (ys, ye) == (5, 8)


@pytest.mark.parametrize("R", [1152, 288, 576, 4608, 2304, 144, 72])
def test_all_regions_c_vs_numpy(R):
    assert oc.domaindecomposition(R) == on.domaindecomposition(R)
    for r in range(R):
        assert oc.getxyresextent(R, r) == on.getxyresextent(R, r)
        for ov in (0, 1, 2):
            if ov == 2 and R == 4608:
                continue
            assert oc.getoverlapindices(R, r, ov) == on.getoverlapindices(R, r, ov)
            assert oc.get_trainingdataindices(R, r, ov) == on.get_trainingdataindices(R, r, ov)


def test_unsupported_region_count_is_an_error_not_a_crash():
    # the Fortran loop would reach MOD(ygrid,0) (src/res_domain.f90:268-269)
    with pytest.raises(ValueError):
        oc.domaindecomposition(658)
    with pytest.raises(ValueError):
        on.domaindecomposition(658)


def test_region_classes_at_1152():
    cnt = Counter()
    for r in range(1152):
        *_, pole, per = oc.getoverlapindices(1152, r, 1)
        cnt[(pole, per)] += 1
    assert cnt == {(False, False): 1012, (True, False): 92, (False, True): 44, (True, True): 4}


def test_region_order_south_to_north_then_west_to_east():
    assert oc.getxyresextent(1152, 0)[:4] == (1, 2, 1, 2)
    assert oc.getxyresextent(1152, 1)[:4] == (1, 2, 3, 4)
    assert oc.getxyresextent(1152, 24)[:4] == (3, 4, 1, 2)
    assert oc.getxyresextent(1152, 1151)[:4] == (95, 96, 47, 48)


def test_halo_wrap_and_pole_clip():
    g = np.arange(96 * 48, dtype=np.float64).reshape((96, 48), order="F")
    west = oc.tileoverlapgrid2d(g, 1152, 5, 1)       # x-tile 0 -> global x [96,1,2,3]
    assert list((west[:, 0] % 96).astype(int) + 1) == [96, 1, 2, 3]
    east = oc.tileoverlapgrid2d(g, 1152, 1151 - 5, 1)  # x-tile 47 -> [94,95,96,1]
    assert list((east[:, 0] % 96).astype(int) + 1) == [94, 95, 96, 1]
    south = oc.tileoverlapgrid2d(g, 1152, 24 * 3, 1)
    assert south.shape == (4, 3)
    assert oc.get_trainingdataindices(1152, 24 * 3, 1) == (2, 3, 1, 2)
    north = oc.tileoverlapgrid2d(g, 1152, 24 * 3 + 23, 1)
    assert north.shape == (4, 3)
    assert oc.get_trainingdataindices(1152, 24 * 3 + 23, 1) == (2, 3, 2, 3)


@pytest.mark.parametrize("R", [1152, 288])
def test_tilers_c_vs_numpy(R):
    rng = np.random.default_rng(3)
    g4 = np.asfortranarray(rng.standard_normal((4, 96, 48, 8)))
    g2 = np.asfortranarray(rng.standard_normal((96, 48)))
    for r in list(range(0, R, 37)) + [R - 1, R // 24 * 24 - 1]:
        assert np.array_equal(oc.tileoverlapgrid4d(g4, R, r, 1), on.tileoverlapgrid4d(g4, R, r, 1))
        assert np.array_equal(oc.tileoverlapgrid2d(g2, R, r, 1), on.tileoverlapgrid2d(g2, R, r, 1))


@pytest.mark.parametrize("nvl", [1, 2, 4, 8])
def test_vertical_localisation_c_vs_numpy(nvl):
    for level in range(1, nvl + 1):
        for vov in (0, 1, 2):
            assert oc.getoverlapindices_vert(nvl, level, vov) == on.getoverlapindices_vert(nvl, level, vov)
            assert oc.get_trainingdataindices_vert(nvl, level, vov) == on.get_trainingdataindices_vert(nvl, level, vov)


@pytest.mark.parametrize("P", [1, 2, 4, 8, 5, 7, 1152])
def test_processor_decomposition(P):
    seen = []
    for p in range(P):
        a = oc.processor_decomposition(p, P, 1152)
        assert a == on.processor_decomposition(p, P, 1152)
        seen += a
    assert sorted(seen) == list(range(1152))  # every region owned exactly once
    if 1152 % P == 0:
        assert oc.processor_decomposition(1, P, 1152)[0] == 1152 // P if P > 1 else True


def test_sizes_at_config():
    # SURVEY.md 2.3 table
    cases = {  # region -> (D, P, S, n, k)
        555: (576, 136, 132, 5760, 33177),
    }
    for region, exp in cases.items():
        r = oc.Region(1152, region)
        assert (r.D, r.P, r.S, r.n, r.k) == exp
    land = oc.Region(1152, 555, sst_bool_input=False)
    assert (land.D, land.n, land.k) == (560, 6160, 37945)
    polar = oc.Region(1152, 24 * 7)
    assert (polar.D, polar.n, polar.k) == (432, 6048, 36578)
    polar_land = oc.Region(1152, 24 * 7, sst_bool_input=False)
    assert (polar_land.D, polar_land.n, polar_land.k) == (420, 5880, 34574)
    assert oc.find_closest_divisor((12000 - 240) // (20 * 6), (12000 - 240) // 6) == 98
