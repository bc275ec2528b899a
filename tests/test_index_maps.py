"""The engine's flattened int32 gather/scatter maps (speedy-ml_b200/csrc/resdomain.hpp, through the C ABI;
host-only functions, no GPU needed) must equal, bit for bit, what the oracle's slice-by-slice tilers do."""
import numpy as np
import pytest

from helpers import oc, on

XG, YG, ZG = 96, 48, 8


@pytest.fixture(scope="module")
def eng(pkg):
    import importlib
    return importlib.import_module("speedy-ml_b200.engine")


def _global_buffers(eng):
    lay = eng.global_layout()
    G = np.arange(lay["g_total"], dtype=np.float64)
    w4d = G[lay["w4d"]:lay["w2d"]].reshape((4, XG, YG, ZG), order="F")
    w2d = G[lay["w2d"]:lay["precip"]].reshape((XG, YG), order="F")
    wp = G[lay["precip"]:lay["sst"]].reshape((XG, YG), order="F")
    wsst = G[lay["sst"]:lay["tisr"]].reshape((XG, YG), order="F")
    tisr = G[lay["tisr"]:lay["g_total"]].reshape((XG, YG), order="F")
    return lay, w4d, w2d, wp, wsst, tisr


@pytest.mark.parametrize("R", [1152, 288, 576, 4608])
def test_index_functions_vs_oracle(eng, R):
    assert eng.domaindecomposition(R) == oc.domaindecomposition(R)
    for r in range(R):
        assert eng.getxyresextent(R, r) == oc.getxyresextent(R, r)
        for ov in (0, 1, 2):
            if ov == 2 and R == 4608:
                continue
            assert eng.getoverlapindices(R, r, ov) == oc.getoverlapindices(R, r, ov)
            assert eng.get_trainingdataindices(R, r, ov) == oc.get_trainingdataindices(R, r, ov)


def test_unsupported_count(eng):
    with pytest.raises(ValueError):
        eng.domaindecomposition(658)


@pytest.mark.parametrize("P", [1, 2, 4, 8, 5, 7])
def test_processor_decomposition_vs_oracle(eng, P):
    for p in range(P):
        assert eng.processor_decomposition(p, P, 1152) == oc.processor_decomposition(p, P, 1152)


@pytest.mark.parametrize("R,precip,sst", [(1152, True, True), (1152, True, False), (1152, False, True), (288, True, True)])
def test_maps_bit_exact_all_regions(eng, R, precip, sst):
    lay, w4d, w2d, wp, wsst, tisr = _global_buffers(eng)
    F4 = np.arange(4 * XG * YG * ZG, dtype=np.float64).reshape((4, XG, YG, ZG), order="F")
    F2 = (4 * XG * YG * ZG + np.arange(XG * YG, dtype=np.float64)).reshape((XG, YG), order="F")
    step = 1 if R <= 1152 else 7
    for r in range(0, R, step):
        m = eng.region_maps(R, r, 1, precip, sst)
        dims = eng.region_dims(R, r, 1, precip_bool=precip, sst_bool=True, sst_bool_input=sst)
        reg = on.Region(R, r, precip_bool=precip, sst_bool=True, sst_bool_input=sst)
        assert (reg.n, reg.k, reg.D, reg.P, reg.S, reg.mean_std_length) == tuple(dims[k] for k in "nkDPSL")
        # input vector: what each feedback element reads
        head = on.tile_4d_and_logp_to_local_state_input(R, r, 1, precip, w4d, w2d, wp)
        exp = np.zeros(reg.D)
        exp[:head.size] = head
        if sst:
            exp[reg.sst_start - 1:reg.sst_end] = on.tileoverlapgrid2d(wsst, R, r, 1).ravel(order="F")
        exp[reg.tisr_start - 1:reg.tisr_end] = on.tileoverlapgrid2d(tisr, R, r, 1).ravel(order="F")
        assert np.array_equal(m["input_map"], exp.astype(np.int64))
        # mean/std slots of the input vector (standardize_state_vec_input + precip/tisr/sst)
        ms = np.full(reg.D, -1)
        ix, iy = reg.inputxchunk, reg.inputychunk
        slots = np.zeros((4, ix, iy, ZG), dtype=int, order="F")
        for v in range(4):
            for z in range(ZG):
                slots[v, :, :, z] = v * ZG + z
        ms[:reg.atmo3d_end] = slots.ravel(order="F")
        ms[reg.logp_start - 1:reg.logp_end] = reg.logp_idx - 1
        if precip:
            ms[reg.precip_start - 1:reg.precip_end] = reg.precip_idx - 1
        if sst:
            ms[reg.sst_start - 1:reg.sst_end] = reg.mean_std_length  # extra slot L
        ms[reg.tisr_start - 1:reg.tisr_end] = reg.tisr_idx - 1
        assert np.array_equal(m["input_ms"], ms)
        # output scatter: write element index+1 through the oracle tiler and read where it landed
        o4 = np.zeros((4, XG, YG, ZG), order="F")
        o2 = np.zeros((XG, YG), order="F")
        op = np.zeros((XG, YG), order="F")
        on.tile_full_grid_with_local_state_vec_res1d(R, r, precip, np.arange(1, reg.P + 1, dtype=np.float64), o4, o2, op)
        Gout = np.zeros(lay["g_total"])
        Gout[lay["w4d"]:lay["w2d"]] = o4.ravel(order="F")
        Gout[lay["w2d"]:lay["precip"]] = o2.ravel(order="F")
        Gout[lay["precip"]:lay["sst"]] = op.ravel(order="F")
        assert np.array_equal(Gout[m["output_map"]], np.arange(1, reg.P + 1))
        assert np.count_nonzero(Gout) == reg.P
        # local_model gather out of F
        lm = on.tile_4d_and_logp_full_grid_to_local_res_vec(R, r, F4, F2)
        assert np.array_equal(m["model_map"], lm.astype(np.int64))
        # target rows of the input vector
        sv = np.arange(reg.D, dtype=np.float64).reshape((reg.D, 1), order="F")
        tgt = on.tile_full_input_to_target_data(reg, sv)[:, 0]
        assert np.array_equal(m["target_map"], tgt.astype(np.int64))
        # un-standardise slots of the outvec
        rx, ry = reg.resxchunk, reg.resychunk
        oms = np.zeros((4, rx, ry, ZG), dtype=int, order="F")
        for v in range(4):
            for z in range(ZG):
                oms[v, :, :, z] = v * ZG + z
        exp_oms = list(oms.ravel(order="F")) + [reg.logp_idx - 1] * (rx * ry) + ([reg.precip_idx - 1] * (rx * ry) if precip else [])
        assert np.array_equal(m["output_ms"], np.asarray(exp_oms))
        assert np.array_equal(m["model_ms"], np.asarray(exp_oms[:reg.S]))


# ---- vertical localisation: num_vert_levels in {1, 2, 4, 8} (src/res_domain.f90:143-153, 206-256, 576-600) ---------
import ctypes as _C

VERT_LAYOUTS = [(1, 0), (2, 0), (2, 1), (2, 2), (4, 0), (4, 1), (4, 2), (8, 0), (8, 1)]


@pytest.mark.parametrize("nvl,vov", VERT_LAYOUTS)
def test_vertical_index_functions_vs_oracle(eng, nvl, vov):
    for level in range(1, nvl + 1):
        zc = 8 // nvl
        assert eng.get_z_res_extent(nvl, level) == ((level - 1) * zc + 1, level * zc, zc)
        assert eng.getoverlapindices_vert(nvl, level, vov) == oc.getoverlapindices_vert(nvl, level, vov)
        assert eng.get_trainingdataindices_vert(nvl, level, vov) == tuple(oc.get_trainingdataindices_vert(nvl, level, vov))


@pytest.mark.parametrize("nvl,vov", [(2, 1), (4, 0), (4, 2), (8, 1)])
@pytest.mark.parametrize("precip,sst", [(True, True), (False, False)])
def test_vertical_slab_sizes_and_maps_bit_exact(eng, nvl, vov, precip, sst):
    """every (region, level) reservoir of a vertically localised model: sizes of allocate_res_new and the flattened
    gather / scatter maps against the oracle's slice-by-slice tilers"""
    R = 1152
    lay, w4d, w2d, wp, wsst, tisr = _global_buffers(eng)
    F4 = np.arange(4 * XG * YG * ZG, dtype=np.float64).reshape((4, XG, YG, ZG), order="F")
    F2 = (4 * XG * YG * ZG + np.arange(XG * YG, dtype=np.float64)).reshape((XG, YG), order="F")
    L = oc.lib()
    for r in (0, 23, 555, 24 * 47 + 23, 1151):
        for level in range(1, nvl + 1):
            reg = oc.Region(R, r, precip_bool=precip, sst_bool=True, sst_bool_input=sst, num_vert_levels=nvl,
                            vert_level=level, vert_overlap=vov)
            d = eng.region_dims_vert(R, r, nvl, level, vov, precip_bool=precip, sst_bool=True, sst_bool_input=sst)
            assert (reg.n, reg.k, reg.D, reg.P, reg.S, reg.L) == tuple(d[k] for k in "nkDPSL"), (r, level)
            m = eng.region_maps_vert(R, r, nvl, level, vov, precip_bool=precip, sst_bool_input=sst)
            g, dd = reg.g, reg.d
            ixy = g.inputxchunk * g.inputychunk
            # input vector
            exp = np.zeros(reg.D)
            a4 = oc.tileoverlapgrid4d(w4d, R, r, 1, nvl, level, vov).ravel(order="F")
            exp[:a4.size] = a4
            if g.bottom:
                exp[g.logp_start - 1:g.logp_end] = oc.tileoverlapgrid2d(w2d, R, r, 1).ravel(order="F")
                if precip:
                    exp[g.precip_start - 1:g.precip_end] = oc.tileoverlapgrid2d(wp, R, r, 1).ravel(order="F")
                if sst:
                    exp[g.sst_start - 1:g.sst_end] = oc.tileoverlapgrid2d(wsst, R, r, 1).ravel(order="F")
            exp[g.tisr_start - 1:g.tisr_end] = oc.tileoverlapgrid2d(tisr, R, r, 1).ravel(order="F")
            assert np.array_equal(m["input_map"], exp.astype(np.int64)), (r, level)
            assert g.tisr_end == reg.D and a4.size == 4 * ixy * g.inputzchunk
            # output scatter through the oracle tiler
            o4, o2, op = np.zeros((4, XG, YG, ZG), order="F"), np.zeros((XG, YG), order="F"), np.zeros((XG, YG), order="F")
            sv = np.arange(1, reg.P + 1, dtype=np.float64)
            L.orc_tile_full_grid_with_local_state_vec_res1d(R, r, nvl, level, int(precip), sv.ctypes.data_as(_C.POINTER(_C.c_double)),
                                                            reg.P, o4.ctypes.data_as(_C.POINTER(_C.c_double)),
                                                            o2.ctypes.data_as(_C.POINTER(_C.c_double)),
                                                            op.ctypes.data_as(_C.POINTER(_C.c_double)))
            Gout = np.zeros(lay["g_total"])
            Gout[lay["w4d"]:lay["w2d"]] = o4.ravel(order="F")
            Gout[lay["w2d"]:lay["precip"]] = o2.ravel(order="F")
            Gout[lay["precip"]:lay["sst"]] = op.ravel(order="F")
            assert np.array_equal(Gout[m["output_map"]], np.arange(1, reg.P + 1)), (r, level)
            assert np.count_nonzero(Gout) == reg.P
            # local_model gather
            lm = np.zeros(max(reg.S, 1))
            L.orc_tile_4d_and_logp_full_grid_to_local_res_vec(R, r, nvl, level, F4.ctypes.data_as(_C.POINTER(_C.c_double)),
                                                              F2.ctypes.data_as(_C.POINTER(_C.c_double)),
                                                              lm.ctypes.data_as(_C.POINTER(_C.c_double)))
            assert np.array_equal(m["model_map"], lm[:reg.S].astype(np.int64)), (r, level)
            # target rows
            svin = np.asfortranarray(np.arange(reg.D, dtype=np.float64).reshape(-1, 1))
            tgt = np.zeros((reg.P, 1), order="F")
            L.orc_tile_full_input_to_target_data2d(_C.byref(g), _C.byref(dd), svin.ctypes.data_as(_C.POINTER(_C.c_double)), reg.D, 1,
                                                   tgt.ctypes.data_as(_C.POINTER(_C.c_double)))
            assert np.array_equal(m["target_map"], tgt[:, 0].astype(np.int64)), (r, level)
            # un-standardise slots: push "mean = slot index, std = 1" through the oracle's unstandardize on a zero vector
            mean = np.arange(reg.L, dtype=np.float64)
            std = np.ones(reg.L)
            ov = np.zeros(reg.P)
            L.orc_unstandardize_state_vec_res(_C.byref(g), _C.byref(dd), mean.ctypes.data_as(_C.POINTER(_C.c_double)),
                                              std.ctypes.data_as(_C.POINTER(_C.c_double)), ov.ctypes.data_as(_C.POINTER(_C.c_double)))
            assert np.array_equal(m["output_ms"], ov.astype(np.int64)), (r, level)
            assert np.array_equal(m["model_ms"], ov[:reg.S].astype(np.int64))
