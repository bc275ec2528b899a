"""The reference's per-region weight container (NetCDF classic, seven variables): scipy writer/reader round trip, the
engine library's own C++ header reader (CPU), and upload-from-file parity on the GPU."""
import importlib

import numpy as np
import pytest

from helpers import c_region, region_weights, rel_inf

E = importlib.import_module("speedy-ml_b200.engine")
W = importlib.import_module("speedy-ml_b200.weights_io")


def _case(tmp_path, region=555, m=600):
    w = region_weights(1152, region, m=m)
    path = str(tmp_path / W.trained_res_filename(region, "trial"))
    W.write_trained_res(path, w["win"], w["wout"], w["rows"], w["cols"], w["vals"], w["mean"], w["std"])
    return w, path


def test_container_round_trip_and_header(tmp_path):
    w, path = _case(tmp_path)
    assert path.endswith("worker_0555_level_1_trial.nc")
    assert open(path, "rb").read(4) == b"CDF\x01"                     # classic format, as nf90_create(NF90_CLOBBER) writes
    g = W.read_trained_res(path)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    assert np.array_equal(g["win"], f32(w["win"])) and np.array_equal(g["wout"], f32(w["wout"]))
    assert np.array_equal(g["rows"], w["rows"]) and np.array_equal(g["cols"], w["cols"])
    assert np.array_equal(g["vals"], f32(w["vals"])) and np.array_equal(g["mean"], f32(w["mean"]))
    # the engine library's own reader (no NetCDF library, no GPU needed for the header)
    d = E.trained_res_dims(path)
    assert d == dict(n=w["n"], k=w["k"], D=w["D"], P=w["P"], S=w["S"], L=w["L"])
    with pytest.raises(E.EngineError):
        E.trained_res_dims(str(tmp_path / "missing.nc"))
    bad = tmp_path / "hdf5.nc"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(E.EngineError, match="classic"):
        E.trained_res_dims(str(bad))


@pytest.mark.gpu
def test_upload_from_file_matches_oracle_with_float32_weights(tmp_path):
    region = 556
    w, path = _case(tmp_path, region=region, m=900)
    g = W.read_trained_res(path)                                     # what read_trained_res hands the Fortran host
    w32 = dict(w)
    w32.update(win=g["win"], wout=g["wout"], vals=g["vals"], mean=g["mean"], std=g["std"])
    rc = c_region(w32)
    eng = E.Engine(number_of_regions=1152, irank=region, numprocs=1152)
    eng.region_upload_file(path, region, sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    rng = np.random.default_rng(4)
    x0, fb, lm = 0.3 * rng.standard_normal(w["n"]), rng.standard_normal(w["D"]), rng.standard_normal(w["S"])
    rc.x[:], rc.feedback[:], rc.local_model[:] = x0, fb, lm
    eng.state_set(region, x0)
    eng.feedback_set(region, fb)
    eng.local_model_set(region, lm)
    for _ in range(3):
        rc.predict()
        eng.predict()
    assert rel_inf(eng.state_get(region), rc.x) < 1e-12
    assert rel_inf(eng.outvec_get(region), rc.outvec) < 1e-12
    assert np.array_equal(eng.wout_get(region), g["wout"])          # float32 values widened exactly
    eng.close()
