"""Case files for the compiled C++ host program (speedy-ml_b200/drivers/replay_main.cpp): every local region's
trained reservoir, its start state and synchronize inputs, and the static fields, in one little-endian binary."""
import numpy as np

N4, N2 = 4 * 96 * 48 * 8, 96 * 48


def write_case(path, ws, x0, fb0, lm0, sync_inputs, fields, nsteps, number_of_regions=1152, overlap=1,
               precip_bool=True, slab_ocean_model_bool=True, ml_only=False, sst_prescribed=True):
    sync_len = sync_inputs[0].shape[1] if sync_inputs else 0
    with open(path, "wb") as f:
        f.write(b"SMLCASE1")
        np.array([number_of_regions, overlap, int(precip_bool), int(slab_ocean_model_bool), int(ml_only),
                  int(sst_prescribed), len(ws), nsteps, sync_len], dtype=np.int32).tofile(f)
        for i, w in enumerate(ws):
            np.array([w["region"], w["n"], w["k"], w["D"], w["P"], w["S"], w["L"], int(w["sst_bool_input"])],
                     dtype=np.int32).tofile(f)
            np.array([1.0], dtype=np.float64).tofile(f)              # leakage
            np.ascontiguousarray(w["rows"], dtype=np.int32).tofile(f)
            np.ascontiguousarray(w["cols"], dtype=np.int32).tofile(f)
            np.ascontiguousarray(w["vals"], dtype=np.float64).tofile(f)
            np.ascontiguousarray(w["winc"], dtype=np.float64).tofile(f)
            np.ascontiguousarray(w["wcol"], dtype=np.int32).tofile(f)
            np.asfortranarray(w["wout"], dtype=np.float64).ravel(order="F").tofile(f)
            np.ascontiguousarray(w["mean"], dtype=np.float64).tofile(f)
            np.ascontiguousarray(w["std"], dtype=np.float64).tofile(f)
            np.ascontiguousarray(x0[i], dtype=np.float64).tofile(f)
            np.ascontiguousarray(fb0[i], dtype=np.float64).tofile(f)
            np.ascontiguousarray(lm0[i], dtype=np.float64).tofile(f)
            if sync_len:
                np.asfortranarray(sync_inputs[i], dtype=np.float64).ravel(order="F").tofile(f)
        for name in ("clim4d", "clim2d", "tisr", "base_sst", "sea_mask"):
            np.asfortranarray(fields[name], dtype=np.float64).ravel(order="F").tofile(f)


def read_output(path, ws, nsteps):
    raw = np.fromfile(path, dtype=np.float64)
    per = N4 + 3 * N2
    steps = []
    for t in range(nsteps):
        b = raw[t * per:(t + 1) * per]
        steps.append((b[:N4].reshape((4, 96, 48, 8), order="F"), b[N4:N4 + N2].reshape((96, 48), order="F"),
                      b[N4 + N2:N4 + 2 * N2].reshape((96, 48), order="F"), b[N4 + 2 * N2:].reshape((96, 48), order="F")))
    pos = nsteps * per
    outvec, feedback = [], []
    for w in ws:
        outvec.append(raw[pos:pos + w["P"]]); pos += w["P"]
        feedback.append(raw[pos:pos + w["D"]]); pos += w["D"]
    assert pos == raw.size
    return steps, outvec, feedback
