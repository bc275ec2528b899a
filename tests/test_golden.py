"""Committed vectors (tests/golden/hotpath_v1.npz, written by tools/make_golden.py -- see its header for the
provenance: oracle outputs, NOT reference outputs; the reference-held known answers are the ref_* keys)."""
import importlib
import os

import numpy as np
import pytest

import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import make_golden as mg  # noqa: E402
from helpers import oc, on, ocean_weights, region_weights, rel_inf  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.npz"))
E = importlib.import_module("speedy-ml_b200.engine")


def test_reference_known_answers():
    # tests/mod_unit_test.f90:63-96: 288 regions -> 4x4 tiles, region 145 spans x 49..52
    assert list(oc.domaindecomposition(288)) == GOLD["ref_unit_test_288_tile"].tolist()
    xs, xe, *_ = oc.getxyresextent(288, 145)
    assert [xs, xe] == GOLD["ref_unit_test_288_region145_x"].tolist()
    assert list(E.getxyresextent(288, 145)[:2]) == [49, 52]
    # tests/mod_unit_test.f90:16-47: pinv(diag(1..10)) = diag(1/i)
    assert np.allclose(on.pinv_svd(np.diag(np.arange(1.0, 11.0))), GOLD["ref_pinv_diag_1_10"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("R", [1152, 288, 4608])
def test_index_tables_oracle_and_library(R):
    ext, ov, td = GOLD[f"extent_{R}"], GOLD[f"overlap_{R}"], GOLD[f"tdata_{R}"]
    step = 1 if R <= 1152 else 7
    for r in range(0, R, step):
        assert list(oc.getxyresextent(R, r)) == ext[r].tolist()
        assert list(on.getxyresextent(R, r)) == ext[r].tolist()
        assert list(E.getxyresextent(R, r)) == ext[r].tolist()          # host integer code of the C-ABI library
        assert [int(v) for v in E.getoverlapindices(R, r, 1)] == ov[r].tolist()
        assert list(E.get_trainingdataindices(R, r, 1)) == td[r].tolist()


def test_processor_decomposition_tables():
    for world in (3, 5, 8):
        for rank in (0, 1, world - 1):
            want = GOLD[f"procdecomp_{world}_{rank}"].tolist()
            assert oc.processor_decomposition(rank, world, 1152) == want
            assert E.processor_decomposition(rank, world, 1152) == want


@pytest.mark.parametrize("region", mg.PREDICT_REGIONS)
def test_oracle_reproduces_golden_predict(region):
    # bit-identical on the machine that wrote the file; 1e-13 leaves room for a libm whose tanh differs by an ulp
    x, outs = mg.predict_case(region)
    assert rel_inf(x, GOLD[f"predict_x_{region}"]) < 1e-13
    assert rel_inf(outs, GOLD[f"predict_out_{region}"]) < 1e-13


def test_oracle_reproduces_golden_ocean():
    x, outs = mg.ocean_case(555)
    assert rel_inf(x, GOLD["ocean_x_555"]) < 1e-13 and rel_inf(outs, GOLD["ocean_out_555"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("region", mg.PREDICT_REGIONS)
def test_engine_matches_golden_predict(region):
    w = region_weights(1152, region, m=mg.M)
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    eng.region_upload(region, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win=w["win"],
                      sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    rng = np.random.default_rng(9000 + region)
    series = np.asfortranarray(rng.standard_normal((w["D"], 5)))
    model = np.asfortranarray(rng.standard_normal((w["S"], 5)))
    eng.synchronize(region, series[:, :2])
    for i, t in enumerate(range(2, 5)):
        eng.feedback_set(region, series[:, t])
        eng.local_model_set(region, model[:, t])
        eng.predict()
        assert rel_inf(eng.outvec_get(region), GOLD[f"predict_out_{region}"][i]) < 1e-12   # 3 steps, FP64
    assert rel_inf(eng.state_get(region), GOLD[f"predict_x_{region}"]) < 1e-12
    eng.close()


@pytest.mark.gpu
def test_engine_matches_golden_ocean():
    region = 555
    wa = region_weights(1152, region, m=mg.M, sst_bool_input=True, with_dense_win=False)
    wo = ocean_weights(1152, region, m=mg.M, mean=wa["mean"], std=wa["std"], with_dense_win=False)
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    eng.region_upload(region, wa["rows"], wa["cols"], wa["vals"], wa["wout"], wa["mean"], wa["std"],
                      win_compact=wa["winc"], win_col=wa["wcol"], D=wa["D"], sst_bool_input=True)
    eng.region_upload(region, wo["rows"], wo["cols"], wo["vals"], wo["wout"], wo["mean"], wo["std"],
                      win_compact=wo["winc"], win_col=wo["wcol"], D=wo["D"], kind=E.OCEAN)
    eng.finalize()
    rng = np.random.default_rng(9500 + region)
    for i in range(3):
        eng.feedback_set(region, rng.standard_normal(wo["D"]), kind=E.OCEAN)
        eng.predict(kind=E.OCEAN)
        assert rel_inf(eng.outvec_get(region, kind=E.OCEAN), GOLD["ocean_out_555"][i]) < 1e-12
    eng.close()
