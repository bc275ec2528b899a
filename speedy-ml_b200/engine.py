"""Host-side mirror of the reference call surface over the C ABI (include/speedyml_engine.h).

Method names follow the reference procedures they stand for (mod_reservoir.f90 / mpires.f90 /
mod_linalg.f90 / res_domain.f90); arguments keep their meaning.  Everything that computes goes
through libspeedyml_b200.so -- there is no NumPy/torch fallback here: if the library cannot be
loaded or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

ATMO, OCEAN, ALL_REGIONS = 0, 1, -1
XGRID, YGRID, ZGRID = 96, 48, 8

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class SmlParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "number_of_regions", "overlap", "precip_bool", "slab_ocean_model_bool", "ml_only", "irank", "numprocs",
        "device", "timestep", "timestep_slab", "sst_prescribed")] + [("reserved", C.c_int32 * 5)]


class SmlRegionWeights(C.Structure):
    _fields_ = [("region", C.c_int32), ("kind", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("D", C.c_int32),
                ("P", C.c_int32), ("S", C.c_int32), ("L", C.c_int32), ("sst_bool_input", C.c_int32),
                ("reserved0", C.c_int32), ("leakage", C.c_double), ("sst_mean", C.c_double), ("sst_std", C.c_double),
                ("rows", _ip), ("cols", _ip), ("vals", _dp), ("win_dense", _dp), ("win_compact", _dp),
                ("win_col", _ip), ("wout", _dp), ("mean", _dp), ("std", _dp)]


class EngineError(RuntimeError):
    pass


_LIB = None

# the one primitive sml_comm_bootstrap asks of the host: int allgather(ctx, send, recv, bytes_per_rank)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)

GRID_NONFINITE, GRID_U_RANGE, GRID_V_RANGE, GRID_T_RANGE, GRID_Q_RANGE = 1, 2, 4, 8, 16

_SIGNATURES = {
    "sml_create": ([C.POINTER(C.c_void_p), C.POINTER(SmlParams)], C.c_int),
    "sml_destroy": ([C.c_void_p], C.c_int),
    "sml_last_error": ([C.c_void_p], C.c_char_p),
    "sml_set_stream": ([C.c_void_p, C.c_void_p], C.c_int),
    "sml_synchronize_stream": ([C.c_void_p], C.c_int),
    "sml_num_local_regions": ([C.c_void_p], C.c_int),
    "sml_local_region_ids": ([C.c_void_p, _ip], C.c_int),
    "sml_domaindecomposition": ([C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "sml_getxyresextent": ([C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 6, C.c_int),
    "sml_getoverlapindices": ([C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 8, C.c_int),
    "sml_get_trainingdataindices": ([C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4, C.c_int),
    "sml_processor_decomposition": ([C.c_int, C.c_int, C.c_int, _ip, C.POINTER(C.c_int)], C.c_int),
    "sml_region_dims": ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
                        + [C.POINTER(C.c_int)] * 6, C.c_int),
    "sml_region_maps": ([C.c_int] * 5 + [_ip] * 7, C.c_int),
    "sml_get_z_res_extent": ([C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3, C.c_int),
    "sml_getoverlapindices_vert": ([C.c_int] * 3 + [C.POINTER(C.c_int)] * 5, C.c_int),
    "sml_get_trainingdataindices_vert": ([C.c_int] * 3 + [C.POINTER(C.c_int)] * 2, C.c_int),
    "sml_region_dims_vert": ([C.c_int] * 7 + [C.c_double] + [C.c_int] * 4 + [C.POINTER(C.c_int)] * 6, C.c_int),
    "sml_region_maps_vert": ([C.c_int] * 8 + [_ip] * 7, C.c_int),
    "sml_ocean_region_dims": ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [C.POINTER(C.c_int)] * 5, C.c_int),
    "sml_ocean_region_maps": ([C.c_int, C.c_int, C.c_int, _ip, _ip, C.POINTER(C.c_int)], C.c_int),
    "sml_global_layout": ([_lp, _lp, _lp], C.c_int),
    "sml_region_upload": ([C.c_void_p, C.POINTER(SmlRegionWeights)], C.c_int),
    "sml_region_generate": ([C.c_void_p, C.POINTER(SmlRegionWeights), C.c_uint64, C.c_double], C.c_int),
    "sml_region_coo_get": ([C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp], C.c_int),
    "sml_region_win_get": ([C.c_void_p, C.c_int, C.c_int, _dp, _ip], C.c_int),
    "sml_trained_res_dims": ([C.c_char_p] + [C.POINTER(C.c_int)] * 6, C.c_int),
    "sml_region_upload_file": ([C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double], C.c_int),
    "sml_finalize": ([C.c_void_p], C.c_int),
    "sml_sparse_eigen": ([C.c_void_p, C.c_int, C.c_int, C.c_double, _dp, C.POINTER(C.c_int)], C.c_int),
    "sml_adjacency_scale": ([C.c_void_p, C.c_int, _dp], C.c_int),
    "sml_state_set": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_state_get": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_feedback_set": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_feedback_get": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_local_model_set": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_local_model_get": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_outvec_get": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_outvec_set": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_outvec_get_all": ([C.c_void_p, C.c_int, _dp], C.c_int),
    "sml_wout_get": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_wout_set": ([C.c_void_p, C.c_int, C.c_int, _dp], C.c_int),
    "sml_synchronize": ([C.c_void_p, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _lp], C.c_int),
    "sml_predict": ([C.c_void_p, C.c_int], C.c_int),
    "sml_set_contribs": ([C.c_void_p, C.c_int], C.c_int),
    "sml_contribs_get": ([C.c_void_p, C.c_int, _dp, _dp], C.c_int),
    "sml_step_exchange_begin": ([C.c_void_p, C.c_int, _dp, _dp, _dp, _dp], C.c_int),
    "sml_step_exchange_begin_view": ([C.c_void_p, C.c_int] + [C.POINTER(_dp)] * 4, C.c_int),
    "sml_forecast_staging": ([C.c_void_p] + [C.POINTER(_dp)] * 3, C.c_int),
    "sml_grids_get": ([C.c_void_p, _dp, _dp, _dp, _dp], C.c_int),
    "sml_step_exchange_end": ([C.c_void_p, C.c_int, _dp, _dp, _dp], C.c_int),
    "sml_set_overlap": ([C.c_void_p, C.c_int], C.c_int),
    "sml_set_tisr": ([C.c_void_p, _dp], C.c_int),
    "sml_step_predict_ahead": ([C.c_void_p, C.c_int], C.c_int),
    "sml_set_sst_static": ([C.c_void_p, _dp, _dp], C.c_int),
    "sml_set_sst_prescribed": ([C.c_void_p, _dp], C.c_int),
    "sml_exchange_buffers": ([C.c_void_p] + [C.POINTER(C.c_void_p), _lp] * 4, C.c_int),
    "sml_ocean_exchange_buffers": ([C.c_void_p] + [C.POINTER(C.c_void_p), _lp] * 2, C.c_int),
    "sml_ocean_ring_reset": ([C.c_void_p], C.c_int),
    "sml_peer_export": ([C.c_void_p, C.c_void_p], C.c_int),
    "sml_peer_attach": ([C.c_void_p, C.c_void_p, C.c_int], C.c_int),
    "sml_peer_attached": ([C.c_void_p], C.c_int),
    "sml_comm_bootstrap": ([C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "sml_grid_status": ([C.c_void_p, C.POINTER(C.c_int)], C.c_int),
    "sml_grid_status_reset": ([C.c_void_p], C.c_int),
    "sml_set_run_speedy": ([C.c_void_p, C.c_int], C.c_int),
    "sml_run_speedy": ([C.c_void_p, C.POINTER(C.c_int)], C.c_int),
    "sml_step_plan": ([C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4, C.c_int),
    "sml_setup_stats": ([C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int)], C.c_int),
    "sml_peer_check": ([C.c_void_p], C.c_int),
    "sml_step_pack_device": ([C.c_void_p, C.c_int], C.c_int),
    "sml_step_unpack_device": ([C.c_void_p, C.c_int], C.c_int),
    "sml_step_exchange_device": ([C.c_void_p, C.c_int], C.c_int),
    "sml_train_begin": ([C.c_void_p, C.c_int, _ip, C.c_int, C.c_int], C.c_int),
    "sml_train_feed": ([C.c_void_p, _dp, _lp, _dp, _lp, C.c_int, C.c_int], C.c_int),
    "sml_train_global_series": ([C.c_void_p, _dp, _dp, C.c_int], C.c_int),
    "sml_train_feed_global": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
    "sml_train_global_release": ([C.c_void_p], C.c_int),
    "sml_train_set_noise": ([C.c_void_p, C.c_double, C.c_uint64, C.c_double], C.c_int),
    "sml_train_noise_sample": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp], C.c_int),
    "sml_condition_series": ([C.c_void_p, C.c_int, C.c_double], C.c_int),
    "sml_conditioning_stats": ([C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip], C.c_int),
    "sml_train_solve": ([C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double, _ip], C.c_int),
    "sml_train_solver_stats": ([C.c_void_p, C.POINTER(C.c_int)], C.c_int),
    "sml_train_gram_get": ([C.c_void_p, C.c_int, _dp, _dp], C.c_int),
    "sml_train_end": ([C.c_void_p], C.c_int),
    "sml_train_trim": ([C.c_void_p], C.c_int),
    "sml_train_set_overlap": ([C.c_void_p, C.c_int], C.c_int),
    "sml_train_stats": ([C.c_void_p, _dp, _dp, _dp, _dp], C.c_int),
    "sml_train_stategen_route": ([C.c_void_p], C.c_int),
    "sml_step_l2_keep": ([C.c_void_p, C.c_int], C.c_int),
    "sml_dmma_probe": ([C.c_void_p, _dp], C.c_int),
    "sml_rolling_average_2d": ([C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
    "sml_mldivide": ([C.c_void_p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int], C.c_int),
    "sml_profile": ([C.c_void_p, C.c_int], C.c_int),
    "sml_kernel_times": ([C.c_void_p, _dp, _dp, C.POINTER(C.c_int)], C.c_int),
    "sml_phase_times": ([C.c_void_p, _dp, _dp, C.POINTER(C.c_int)], C.c_int),
    "sml_sync_times": ([C.c_void_p, _dp, _lp], C.c_int),
    "sml_step_chunk_rows": ([C.c_void_p, C.c_int], C.c_int),
    "sml_kernel_launch_count": ([C.c_void_p], C.c_int64),
    "sml_predict_algorithmic_bytes": ([C.c_void_p, C.c_int], C.c_int64),
    "sml_update_algorithmic_bytes": ([C.c_void_p, C.c_int], C.c_int64),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library(path: str | None = None):
    """dlopen the C-ABI library (building it in-tree first if the sources are newer)."""
    global _LIB
    if _LIB is None:
        path = path or _build.build()
        if not os.path.exists(path):
            raise EngineError(f"{path} is missing: build it with python speedy-ml_b200/build.py")
        lib = C.CDLL(path)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library ever diverge
            fn.argtypes, fn.restype = argtypes, restype
        _LIB = lib
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _farr(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class DeviceArray:
    """zero-copy view of an engine-owned device buffer (__cuda_array_interface__), so the host can run
    torch.distributed collectives on it: torch.as_tensor(DeviceArray, device='cuda')."""

    def __init__(self, ptr: int, count: int, owner):
        self.ptr, self.count, self._owner = ptr, count, owner
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2,
                                         "strides": None}


# ------------------------------------------------------------------------------------------ index arithmetic
def domaindecomposition(numregions):
    """res_domain.f90:258-280"""
    fx, fy = C.c_int(), C.c_int()
    if load_library().sml_domaindecomposition(numregions, C.byref(fx), C.byref(fy)):
        raise ValueError(f"{numregions} regions do not tile the grid")
    return fx.value, fy.value


def getxyresextent(num_regions, region):
    """res_domain.f90:123-141 -> (xstart, xend, ystart, yend, xchunk, ychunk), 1-based"""
    v = [C.c_int() for _ in range(6)]
    if load_library().sml_getxyresextent(num_regions, region, *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    return tuple(a.value for a in v)


def getoverlapindices(num_regions, region, overlap):
    """res_domain.f90:155-204"""
    v = [C.c_int() for _ in range(8)]
    if load_library().sml_getoverlapindices(num_regions, region, overlap, *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    t = tuple(a.value for a in v)
    return t[:6] + (bool(t[6]), bool(t[7]))


def get_trainingdataindices(num_regions, region, overlap):
    """res_domain.f90:547-574"""
    v = [C.c_int() for _ in range(4)]
    if load_library().sml_get_trainingdataindices(num_regions, region, overlap, *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    return tuple(a.value for a in v)


def processor_decomposition(irank, numprocs, number_of_regions):
    """res_domain.f90:31-62"""
    buf = np.zeros(number_of_regions // numprocs + 2, dtype=np.int32)
    cnt = C.c_int()
    if load_library().sml_processor_decomposition(irank, numprocs, number_of_regions, _i(buf), C.byref(cnt)):
        raise ValueError("bad rank")
    return buf[:cnt.value].tolist()


def region_dims(num_regions, region, overlap=1, m=6000, deg=6.0, precip_bool=True, sst_bool=True,
                sst_bool_input=True, ml_only=False):
    """allocate_res_new sizes -> dict(n, k, D, P, S, L)"""
    v = [C.c_int() for _ in range(6)]
    if load_library().sml_region_dims(num_regions, region, overlap, m, float(deg), int(precip_bool), int(sst_bool),
                                      int(sst_bool_input), int(ml_only), *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    return dict(zip(("n", "k", "D", "P", "S", "L"), (a.value for a in v)))


def region_maps(num_regions, region, overlap=1, precip_bool=True, sst_bool_input=True):
    """flattened 0-based gather/scatter maps (see sml_region_maps)"""
    d = region_dims(num_regions, region, overlap, precip_bool=precip_bool, sst_bool_input=sst_bool_input)
    D, P, S = d["D"], d["P"], d["S"]
    arrs = [np.zeros(D, np.int32), np.zeros(D, np.int32), np.zeros(P, np.int32), np.zeros(P, np.int32),
            np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros(P, np.int32)]
    if load_library().sml_region_maps(num_regions, region, overlap, int(precip_bool), int(sst_bool_input),
                                      *[_i(a) for a in arrs]):
        raise ValueError("unsupported region count")
    return dict(zip(("input_map", "input_ms", "output_map", "output_ms", "model_map", "model_ms", "target_map"), arrs))


def get_z_res_extent(num_vert_levels, vert_level):
    """res_domain.f90:143-153 -> (zstart, zend, zchunk)"""
    v = [C.c_int() for _ in range(3)]
    if load_library().sml_get_z_res_extent(num_vert_levels, vert_level, *[C.byref(a) for a in v]):
        raise ValueError("unsupported vertical layout")
    return tuple(a.value for a in v)


def getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap):
    """res_domain.f90:206-256 -> (input_zstart, input_zend, inputzchunk, top, bottom)"""
    v = [C.c_int() for _ in range(5)]
    if load_library().sml_getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap, *[C.byref(a) for a in v]):
        raise ValueError("unsupported vertical layout")
    t = tuple(a.value for a in v)
    return t[:3] + (bool(t[3]), bool(t[4]))


def get_trainingdataindices_vert(num_vert_levels, vert_level, vert_overlap):
    """res_domain.f90:576-600 -> (zstart, zend)"""
    v = [C.c_int() for _ in range(2)]
    if load_library().sml_get_trainingdataindices_vert(num_vert_levels, vert_level, vert_overlap, *[C.byref(a) for a in v]):
        raise ValueError("unsupported vertical layout")
    return tuple(a.value for a in v)


def region_dims_vert(num_regions, region, num_vert_levels, vert_level, vert_overlap, overlap=1, m=6000, deg=6.0,
                     precip_bool=True, sst_bool=True, sst_bool_input=True, ml_only=False):
    """allocate_res_new sizes of one vertical slab -> dict(n, k, D, P, S, L)"""
    v = [C.c_int() for _ in range(6)]
    if load_library().sml_region_dims_vert(num_regions, region, overlap, num_vert_levels, vert_level, vert_overlap, m,
                                           float(deg), int(precip_bool), int(sst_bool), int(sst_bool_input), int(ml_only),
                                           *[C.byref(a) for a in v]):
        raise ValueError("unsupported layout")
    return dict(zip(("n", "k", "D", "P", "S", "L"), (a.value for a in v)))


def region_maps_vert(num_regions, region, num_vert_levels, vert_level, vert_overlap, overlap=1, precip_bool=True,
                     sst_bool_input=True):
    d = region_dims_vert(num_regions, region, num_vert_levels, vert_level, vert_overlap, overlap, precip_bool=precip_bool,
                         sst_bool_input=sst_bool_input)
    D, P, S = d["D"], d["P"], d["S"]
    arrs = [np.zeros(D, np.int32), np.zeros(D, np.int32), np.zeros(P, np.int32), np.zeros(P, np.int32),
            np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros(P, np.int32)]
    if load_library().sml_region_maps_vert(num_regions, region, overlap, num_vert_levels, vert_level, vert_overlap,
                                           int(precip_bool), int(sst_bool_input), *[_i(a) for a in arrs]):
        raise ValueError("unsupported layout")
    return dict(zip(("input_map", "input_ms", "output_map", "output_ms", "model_map", "model_ms", "target_map"), arrs))


def ocean_region_dims(num_regions, region, overlap=1, m=4000, deg=6.0):
    """initialize_slab_ocean_model sizes -> dict(n, k, D, P, S=0, A)"""
    v = [C.c_int() for _ in range(5)]
    if load_library().sml_ocean_region_dims(num_regions, region, overlap, m, float(deg), *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    d = dict(zip(("n", "k", "D", "P", "A"), (a.value for a in v)))
    d["S"] = 0
    return d


def ocean_region_maps(num_regions, region, overlap=1):
    d = ocean_region_dims(num_regions, region, overlap)
    sst = np.zeros(d["D"] // 8, np.int32)
    tgt = np.zeros(d["P"], np.int32)
    a0 = C.c_int()
    if load_library().sml_ocean_region_maps(num_regions, region, overlap, _i(sst), _i(tgt), C.byref(a0)):
        raise ValueError("unsupported region count")
    return dict(sst_map=sst, target_map=tgt, atmo_slice0=a0.value)


def trained_res_dims(path):
    """header of a write_trained_res file -> dict(n, k, D, P, S, L)"""
    v = [C.c_int() for _ in range(6)]
    lib = load_library()
    if lib.sml_trained_res_dims(os.fsencode(path), *[C.byref(a) for a in v]):
        raise EngineError(lib.sml_last_error(None).decode())
    return dict(zip(("n", "k", "D", "P", "S", "L"), (a.value for a in v)))


def global_layout():
    off = (C.c_int64 * 5)()
    g, f = C.c_int64(), C.c_int64()
    load_library().sml_global_layout(off, C.byref(g), C.byref(f))
    return dict(w4d=off[0], w2d=off[1], precip=off[2], sst=off[3], tisr=off[4], g_total=g.value, f_total=f.value)


# ------------------------------------------------------------------------------------------ engine
class Engine:
    """one rank's shard of the reservoir model on one B200"""

    def __init__(self, number_of_regions=1152, overlap=1, precip_bool=True, slab_ocean_model_bool=True,
                 ml_only=False, irank=0, numprocs=1, device=0, timestep=6, timestep_slab=168,
                 sst_prescribed=False, stream=None):
        self.lib = load_library()
        self.p = SmlParams(number_of_regions, overlap, int(precip_bool), int(slab_ocean_model_bool), int(ml_only),
                           irank, numprocs, device, timestep, timestep_slab, int(sst_prescribed))
        self.h = C.c_void_p()
        if self.lib.sml_create(C.byref(self.h), C.byref(self.p)):
            raise EngineError(self.lib.sml_last_error(None).decode())
        if stream is not None:
            self.set_stream(stream)
        n = self.lib.sml_num_local_regions(self.h)
        ids = np.zeros(n, dtype=np.int32)
        self.lib.sml_local_region_ids(self.h, _i(ids))
        self.region_indices = ids.tolist()        # model_parameters%region_indices
        self.num_of_regions_on_proc = n
        self.dims = {}                             # (kind, region) -> dict(n, D, P, S, L)
        self._view_ptrs = [_dp() for _ in range(4)]
        self._view_refs = [C.byref(a) for a in self._view_ptrs]
        self._views, self._view_addr = None, None

    # -- plumbing
    def _ck(self, rc, allow_positive=False):
        if rc < 0 or (rc > 0 and not allow_positive):
            raise EngineError(self.lib.sml_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.sml_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: int cudaStream_t or an object with .cuda_stream (torch.cuda.Stream)"""
        ptr = getattr(stream, "cuda_stream", stream)
        self._ck(self.lib.sml_set_stream(self.h, C.c_void_p(ptr)))

    def synchronize_stream(self):
        self._ck(self.lib.sml_synchronize_stream(self.h))

    # -- mklsparse + trained_reservoir_prediction
    def region_upload(self, region, rows, cols, vals, wout, mean, std, win=None, win_compact=None, win_col=None,
                      kind=ATMO, leakage=1.0, sst_bool_input=True, sst_mean=None, sst_std=None, S=None, P=None, D=None):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        std = np.ascontiguousarray(std, dtype=np.float64)
        keep = [rows, cols, vals, mean, std]
        w = SmlRegionWeights()
        if win is not None:
            win = _farr(win)
            n, Dw = win.shape
            w.win_dense = _d(win)
            keep.append(win)
        else:
            win_compact = np.ascontiguousarray(win_compact, dtype=np.float64)
            win_col = np.ascontiguousarray(win_col, dtype=np.int32)
            n, Dw = win_compact.size, D
            w.win_compact, w.win_col = _d(win_compact), _i(win_col)
            keep += [win_compact, win_col]
        if wout is not None:
            wout = _farr(wout)
            Pw, N = wout.shape
            w.wout = _d(wout)
            keep.append(wout)
            Sw = N - n
        else:
            Pw, Sw = P, S
        if D is not None and Dw is None:
            Dw = D
        if Dw is None:
            raise ValueError("D is needed with win_compact")
        w.region, w.kind, w.n, w.k, w.D, w.P, w.S, w.L = region, kind, n, rows.size, Dw, Pw, Sw, mean.size
        w.sst_bool_input = int(sst_bool_input)
        w.leakage = float(leakage)
        w.sst_mean = float(mean[-1] if sst_mean is None else sst_mean)
        w.sst_std = float(std[-1] if sst_std is None else sst_std)
        w.rows, w.cols, w.vals, w.mean, w.std = _i(rows), _i(cols), _d(vals), _d(mean), _d(std)
        self._ck(self.lib.sml_region_upload(self.h, C.byref(w)))
        self.dims[(kind, region)] = dict(n=n, D=Dw, P=Pw, S=Sw, L=mean.size)

    def region_generate(self, region, n, k, D, P, S, mean, std, seed, sigma, wout=None, kind=ATMO, leakage=1.0,
                        sst_bool_input=True, sst_mean=None, sst_std=None):
        """gen_res's makesparse + the W_in build ON THE DEVICE (src/mod_linalg.f90:180-218, src/mod_reservoir.f90:262-283):
        the region is installed with the sizes of allocate_res_new; its structure is drawn at finalize()"""
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        std = np.ascontiguousarray(std, dtype=np.float64)
        w = SmlRegionWeights()
        keep = [mean, std]
        if wout is not None:
            wout = _farr(wout, (P, n + S))
            w.wout = _d(wout)
            keep.append(wout)
        w.region, w.kind, w.n, w.k, w.D, w.P, w.S, w.L = region, kind, n, k, D, P, S, mean.size
        w.sst_bool_input = int(sst_bool_input)
        w.leakage = float(leakage)
        w.sst_mean = float(mean[-1] if sst_mean is None else sst_mean)
        w.sst_std = float(std[-1] if sst_std is None else sst_std)
        w.mean, w.std = _d(mean), _d(std)
        self._ck(self.lib.sml_region_generate(self.h, C.byref(w), int(seed), float(sigma)))
        self.dims[(kind, region)] = dict(n=n, D=D, P=P, S=S, L=mean.size, k=k)

    def region_coo_get(self, region, kind=ATMO):
        """reservoir%rows / cols / vals of a generated region (1-based, makesparse's entry order)"""
        k = self.dims[(kind, region)]["k"]
        rows, cols, vals = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32), np.zeros(k)
        self._ck(self.lib.sml_region_coo_get(self.h, kind, region, _i(rows), _i(cols), _d(vals)))
        return rows, cols, vals

    def region_win_get(self, region, kind=ATMO):
        """(value, 0-based column) of the single non-zero of every W_in row"""
        n = self.dims[(kind, region)]["n"]
        winc, wcol = np.zeros(n), np.zeros(n, dtype=np.int32)
        self._ck(self.lib.sml_region_win_get(self.h, kind, region, _d(winc), _i(wcol)))
        return winc, wcol

    def region_upload_file(self, path, region, kind=ATMO, sst_bool_input=True, leakage=1.0):
        """read_trained_res + mklsparse from the region's NetCDF-classic weight file (float32 -> FP64 on the way)"""
        d = trained_res_dims(path)
        self._ck(self.lib.sml_region_upload_file(self.h, os.fsencode(path), region, kind, int(sst_bool_input), float(leakage)))
        self.dims[(kind, region)] = dict(n=d["n"], D=d["D"], P=d["P"], S=d["S"], L=d["L"])

    def finalize(self):
        self._ck(self.lib.sml_finalize(self.h))

    # -- gen_res: sparse_eigen + rescale to the target spectral radius   mod_reservoir.f90:182-212
    def sparse_eigen(self, kind=ATMO, maxit=500, tol=1e-13):
        """-> (eigs[nloc] in local region order, iterations, converged)"""
        eigs = np.zeros(self.num_of_regions_on_proc)
        it = C.c_int()
        rc = self._ck(self.lib.sml_sparse_eigen(self.h, kind, maxit, tol, _d(eigs), C.byref(it)), allow_positive=True)
        return eigs, it.value, rc == 0

    def adjacency_scale(self, factor, kind=ATMO):
        factor = np.ascontiguousarray(factor, dtype=np.float64)
        assert factor.size == self.num_of_regions_on_proc
        self._ck(self.lib.sml_adjacency_scale(self.h, kind, _d(factor)))

    def gen_res(self, radius, kind=ATMO, maxit=500, tol=1e-13):
        """spectral radius of every local adjacency, then vals <- vals / eig * radius on the device.
        Returns the factors radius/eig so the host can scale reservoir%vals the same way."""
        eigs, it, ok = self.sparse_eigen(kind, maxit, tol)
        if not ok:
            raise EngineError(f"sparse_eigen did not converge in {it} iterations")
        factor = np.where(eigs > 0, radius / np.where(eigs > 0, eigs, 1.0), 1.0)
        self.adjacency_scale(factor, kind)
        return factor, eigs

    # -- reservoir%current_state / feedback / local_model / outvec
    def _get(self, fn, kind, region, size):
        out = np.zeros(size)
        self._ck(fn(self.h, kind, region, _d(out)))
        return out

    def state_get(self, region, kind=ATMO):
        return self._get(self.lib.sml_state_get, kind, region, self.dims[(kind, region)]["n"])

    def state_set(self, region, x, kind=ATMO):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.dims[(kind, region)]["n"]
        self._ck(self.lib.sml_state_set(self.h, kind, region, _d(x)))

    def feedback_get(self, region, kind=ATMO):
        return self._get(self.lib.sml_feedback_get, kind, region, self.dims[(kind, region)]["D"])

    def feedback_set(self, region, v, kind=ATMO):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.dims[(kind, region)]["D"]
        self._ck(self.lib.sml_feedback_set(self.h, kind, region, _d(v)))

    def local_model_get(self, region, kind=ATMO):
        return self._get(self.lib.sml_local_model_get, kind, region, self.dims[(kind, region)]["S"])

    def local_model_set(self, region, v, kind=ATMO):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.dims[(kind, region)]["S"]
        self._ck(self.lib.sml_local_model_set(self.h, kind, region, _d(v)))

    def outvec_get(self, region, kind=ATMO):
        return self._get(self.lib.sml_outvec_get, kind, region, self.dims[(kind, region)]["P"])

    def outvec_get_all(self, P, kind=ATMO):
        """-> (nloc, P) array: every local region's outvec in one copy (local region order)"""
        out = np.zeros((self.num_of_regions_on_proc, P))
        self._ck(self.lib.sml_outvec_get_all(self.h, kind, _d(out)))
        return out

    def outvec_set(self, region, v, kind=ATMO):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.dims[(kind, region)]["P"]
        self._ck(self.lib.sml_outvec_set(self.h, kind, region, _d(v)))

    def wout_get(self, region, kind=ATMO):
        d = self.dims[(kind, region)]
        out = np.zeros((d["P"], d["n"] + d["S"]), order="F")
        self._ck(self.lib.sml_wout_get(self.h, kind, region, _d(out)))
        return out

    def wout_set(self, region, wout, kind=ATMO):
        d = self.dims[(kind, region)]
        wout = _farr(wout, (d["P"], d["n"] + d["S"]))
        self._ck(self.lib.sml_wout_set(self.h, kind, region, _d(wout)))

    # -- synchronize(reservoir, input, x, length)   mod_reservoir.f90:1354
    def synchronize(self, region, inputs, length=None, kind=ATMO):
        inputs = _farr(inputs)
        length = inputs.shape[1] if length is None else length
        self._ck(self.lib.sml_synchronize(self.h, kind, region, _d(inputs), inputs.shape[0], length, None))

    def synchronize_all(self, inputs_by_region, length, kind=ATMO):
        """inputs_by_region: list (local order) of (D_i, >=length) arrays"""
        blocks, offs, pos = [], [], 0
        for a in inputs_by_region:
            a = _farr(a)[:, :length]
            blocks.append(np.asfortranarray(a).ravel(order="F"))
            offs.append(pos)
            pos += blocks[-1].size
        flat = np.concatenate(blocks)
        offs = np.asarray(offs, dtype=np.int64)
        self._ck(self.lib.sml_synchronize(self.h, kind, ALL_REGIONS, _d(flat), 0, length, offs.ctypes.data_as(_lp)))

    # -- predict / predict_ml for every local region   mod_reservoir.f90:1418,1491
    def predict(self, kind=ATMO):
        self._ck(self.lib.sml_predict(self.h, kind))

    # -- outvec_component_contribs: reservoir%v_p, reservoir%v_ml   mod_reservoir.f90:1458-1461
    def set_contribs(self, on=True):
        self._ck(self.lib.sml_set_contribs(self.h, int(on)))

    def contribs_get(self, region):
        P = self.dims[(ATMO, region)]["P"]
        vp, vml = np.zeros(P), np.zeros(P)
        self._ck(self.lib.sml_contribs_get(self.h, region, _d(vp), _d(vml)))
        return vp, vml

    # -- sendrecievegrid   mpires.f90:218
    def set_sst_static(self, base_sst_grid, sea_mask):
        b, m = _farr(base_sst_grid, (XGRID, YGRID)), _farr(sea_mask, (XGRID, YGRID))
        self._ck(self.lib.sml_set_sst_static(self.h, _d(b), _d(m)))

    def set_sst_prescribed(self, sst_grid):
        s = _farr(sst_grid, (XGRID, YGRID))
        self._ck(self.lib.sml_set_sst_prescribed(self.h, _d(s)))

    def step_exchange_begin(self, timestep, copy_out=True, reuse=False):
        """-> (wholegrid4d, wholegrid2d, wholegrid_precip, wholegrid_sst).  reuse=True returns the same four
        arrays on every call (overwritten by the next call) instead of allocating 1.3 MB per step."""
        if not copy_out:
            self._ck(self.lib.sml_step_exchange_begin(self.h, timestep, None, None, None, None))
            return None
        self.grid_nonfinite = False
        if reuse and getattr(self, "_grids", None) is not None:
            w4d, w2d, wp, wsst = self._grids
        else:
            w4d = np.empty((4, XGRID, YGRID, ZGRID), order="F")
            w2d = np.empty((XGRID, YGRID), order="F")
            wp = np.empty((XGRID, YGRID), order="F")
            wsst = np.empty((XGRID, YGRID), order="F")
            if reuse:
                self._grids = (w4d, w2d, wp, wsst)
        rc = self._ck(self.lib.sml_step_exchange_begin(self.h, timestep, _d(w4d), _d(w2d), _d(wp), _d(wsst)),
                      allow_positive=True)
        self.grid_nonfinite = rc > 0        # the assembled grid holds a NaN / Inf (the grids are returned all the same)
        return w4d, w2d, wp, wsst

    def step_exchange_begin_view(self, timestep):
        """zero-copy: the four grids as read-only views of the engine's pinned staging (valid until the next begin)"""
        p = self._view_ptrs
        rc = self._ck(self.lib.sml_step_exchange_begin_view(self.h, timestep, *self._view_refs), allow_positive=True)
        self.grid_nonfinite = rc > 0
        addr = tuple(C.cast(a, C.c_void_p).value for a in p)
        if addr != self._view_addr:      # the staging is fixed for the engine's life: the NumPy views are built once
            shapes = ((4, XGRID, YGRID, ZGRID), (XGRID, YGRID), (XGRID, YGRID), (XGRID, YGRID))
            out = []
            for ptr, shp in zip(p, shapes):
                a = np.ctypeslib.as_array(ptr, shape=(int(np.prod(shp)),)).reshape(shp, order="F")
                a.flags.writeable = False
                out.append(a)
            self._views, self._view_addr = tuple(out), addr
        return self._views

    def grids_get(self):
        """(wholegrid4d, wholegrid2d, wholegrid_precip, wholegrid_sst) of the last assembly, from the device, on any rank"""
        w4d = np.empty((4, XGRID, YGRID, ZGRID), order="F")
        w2d, wp, wsst = (np.empty((XGRID, YGRID), order="F") for _ in range(3))
        self._ck(self.lib.sml_grids_get(self.h, _d(w4d), _d(w2d), _d(wp), _d(wsst)))
        return w4d, w2d, wp, wsst

    def forecast_staging(self):
        """(forecast_4d, forecast_2d, tisr) views of the pinned staging step_exchange_end uploads from: write the host
        model's output into them and pass them to step_exchange_end to skip the intermediate copy"""
        p = [_dp() for _ in range(3)]
        self._ck(self.lib.sml_forecast_staging(self.h, *[C.byref(a) for a in p]))
        shapes = ((4, XGRID, YGRID, ZGRID), (XGRID, YGRID), (XGRID, YGRID))
        return tuple(np.ctypeslib.as_array(ptr, shape=(int(np.prod(shp)),)).reshape(shp, order="F")
                     for ptr, shp in zip(p, shapes))

    def step_exchange_end(self, timestep, forecast_4d, forecast_2d, tisr_grid=None):
        f4 = _farr(forecast_4d, (4, XGRID, YGRID, ZGRID)) if forecast_4d is not None else None
        f2 = _farr(forecast_2d, (XGRID, YGRID)) if forecast_2d is not None else None
        ti = _farr(tisr_grid, (XGRID, YGRID)) if tisr_grid is not None else None
        self._ck(self.lib.sml_step_exchange_end(self.h, timestep, _d(f4), _d(f2), _d(ti)))

    # -- overlapped step: the next predict runs while the host model works
    def set_overlap(self, on=True):
        self._ck(self.lib.sml_set_overlap(self.h, int(on)))

    def set_tisr(self, tisr_grid):
        ti = _farr(tisr_grid, (XGRID, YGRID))
        self._ck(self.lib.sml_set_tisr(self.h, _d(ti)))

    def step_predict_ahead(self, timestep):
        self._ck(self.lib.sml_step_predict_ahead(self.h, timestep))

    def step_pack_device(self, timestep=0):
        self._ck(self.lib.sml_step_pack_device(self.h, timestep))

    def step_unpack_device(self, timestep=0):
        self._ck(self.lib.sml_step_unpack_device(self.h, timestep))

    def step_exchange_device(self, timestep=0):
        """pack + unpack of the device-resident step as one cooperative launch"""
        self._ck(self.lib.sml_step_exchange_device(self.h, timestep))

    def exchange_buffers(self):
        """-> dict of DeviceArray: outvec_slab, gathered, G, F"""
        ptrs = [C.c_void_p() for _ in range(4)]
        cnts = [C.c_int64() for _ in range(4)]
        args = []
        for p, c in zip(ptrs, cnts):
            args += [C.byref(p), C.byref(c)]
        self._ck(self.lib.sml_exchange_buffers(self.h, *args))
        names = ("outvec_slab", "gathered", "G", "F")
        return {n: DeviceArray(p.value, c.value, self) for n, p, c in zip(names, ptrs, cnts)}

    # -- fused all-gather over NVLink (CUDA IPC peer stores from the readout-finish kernel)
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.sml_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, handles):
        """handles: list of the 64-byte exports of every rank, in rank order"""
        blob = b"".join(handles)
        self._ck(self.lib.sml_peer_attach(self.h, blob, len(handles)))

    def comm_bootstrap(self, allgather_bytes):
        """the whole multi-rank set-up behind one call.  allgather_bytes(blob: bytes) -> list of every rank's blob in
        rank order is the only thing the host supplies (torch.distributed.all_gather_object, mpi4py allgather, ...)."""
        world = self.p.numprocs

        def cb(_ctx, send, recv, nbytes):
            try:
                blobs = allgather_bytes(C.string_at(send, nbytes))
                if len(blobs) != world or any(len(b) != nbytes for b in blobs):
                    return 1
                C.memmove(recv, b"".join(blobs), nbytes * world)
                return 0
            except Exception:          # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        fn = ALLGATHER_FN(cb)
        self._ck(self.lib.sml_comm_bootstrap(self.h, C.cast(fn, C.c_void_p), None))

    def grid_status(self) -> int:
        """sticky GRID_* bits of the assembled grids since the last reset (synchronises the engine's stream)"""
        bits = C.c_int()
        self._ck(self.lib.sml_grid_status(self.h, C.byref(bits)))
        return bits.value

    def grid_status_reset(self):
        self._ck(self.lib.sml_grid_status_reset(self.h))

    def set_run_speedy(self, flag: bool):
        """model_parameters%run_speedy, set by the root after run_model; travels with the next forecast"""
        self._ck(self.lib.sml_set_run_speedy(self.h, int(bool(flag))))

    def run_speedy(self) -> bool:
        v = C.c_int()
        self._ck(self.lib.sml_run_speedy(self.h, C.byref(v)))
        return bool(v.value)

    def step_plan(self, kind=ATMO):
        v = [C.c_int() for _ in range(4)]
        self._ck(self.lib.sml_step_plan(self.h, kind, *[C.byref(a) for a in v]))
        return dict(kernel="k_step_persist" if v[0].value else "k_step", slots=v[1].value, part_rows=v[2].value,
                    parts=v[3].value, l2_keep=bool(self.lib.sml_step_l2_keep(self.h, kind)))

    def setup_stats(self):
        a, b, c = C.c_double(), C.c_int64(), C.c_int()
        self._ck(self.lib.sml_setup_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(upload_s=a.value, arena_bytes=b.value, arena_chunks=c.value)

    def peer_attached(self) -> bool:
        return bool(self.lib.sml_peer_attached(self.h))

    def peer_check(self):
        self._ck(self.lib.sml_peer_check(self.h))

    def ocean_exchange_buffers(self):
        """-> dict of DeviceArray: ocean_slab [nloc*P_ocean], ocean_gathered [R*P_ocean]"""
        ptrs = [C.c_void_p() for _ in range(2)]
        cnts = [C.c_int64() for _ in range(2)]
        self._ck(self.lib.sml_ocean_exchange_buffers(self.h, C.byref(ptrs[0]), C.byref(cnts[0]), C.byref(ptrs[1]),
                                                     C.byref(cnts[1])))
        return {n: DeviceArray(p.value, c.value, self) for n, p, c in zip(("ocean_slab", "ocean_gathered"), ptrs, cnts)}

    def ocean_ring_reset(self):
        self._ck(self.lib.sml_ocean_ring_reset(self.h))

    # -- training: chunking_matmul / fit_chunk_hybrid   mod_reservoir.f90:1645,1235
    def train_begin(self, regions, batch_size, kind=ATMO):
        regions = np.ascontiguousarray(regions, dtype=np.int32)
        self._train_regions = regions.tolist()
        self._train_kind = kind
        self._ck(self.lib.sml_train_begin(self.h, kind, _i(regions), regions.size, batch_size))

    def train_feed(self, trainingdata_by_region, imperfect_by_region, discard_cols):
        """one phase for the wave: per-region (D_i, ncols) series and (S_i, ncols) imperfect-model series.  The arrays
        are handed over in place (offsets relative to the first one), not concatenated."""
        def in_place(arrays, key):
            keep, offs, base = [], [], None
            for r, a in zip(self._train_regions, arrays):
                a = _farr(a)
                assert a.shape == (self.dims[(self._train_kind, r)][key], ncols), (a.shape, key)
                keep.append(a)
                addr = a.ctypes.data
                base = addr if base is None else base
                assert (addr - base) % 8 == 0
                offs.append((addr - base) // 8)
            return keep, np.asarray(offs, dtype=np.int64)

        ncols = np.shape(trainingdata_by_region[0])[1]
        tds, tdo = in_place(trainingdata_by_region, "D")
        if imperfect_by_region is not None:
            ims, imo = in_place(imperfect_by_region, "S")
            im_ptr, imo_ptr = _d(ims[0]), imo.ctypes.data_as(_lp)
        else:
            ims, im_ptr, imo_ptr = None, None, None
        self._ck(self.lib.sml_train_feed(self.h, _d(tds[0]), tdo.ctypes.data_as(_lp), im_ptr, imo_ptr, ncols, discard_cols))

    def train_global_series(self, G_series, F_series=None):
        """G_series (g_total, T), F_series (f_total, T) column-major: the conditioned global series, kept on the device"""
        lay = global_layout()
        G = _farr(G_series)
        assert G.shape[0] == lay["g_total"]
        F = _farr(F_series) if F_series is not None else None
        assert F is None or F.shape == (lay["f_total"], G.shape[1])
        self._ck(self.lib.sml_train_global_series(self.h, _d(G), _d(F), G.shape[1]))

    def train_feed_global(self, first_col, stride, ncols, discard_cols):
        self._ck(self.lib.sml_train_feed_global(self.h, first_col, stride, ncols, discard_cols))

    def train_set_noise(self, noisemag, seed=0, precip_epsilon=0.001):
        """input noise of sml_train_feed_global: u*(1 + noisemag*N(0,1)), precip in linear space; 0 = off"""
        self._ck(self.lib.sml_train_set_noise(self.h, float(noisemag), int(seed), float(precip_epsilon)))

    def train_noise_sample(self, region, first_col, stride, col):
        """-> (clean, gauss, noisy) input vectors [D] of one region of the current wave at phase column col"""
        D = self.dims[(self._train_kind, region)]["D"]
        a, b, c = np.zeros(D), np.zeros(D), np.zeros(D)
        self._ck(self.lib.sml_train_noise_sample(self.h, region, first_col, stride, col, _d(a), _d(b), _d(c)))
        return a, b, c

    def condition_series(self, period=6, precip_epsilon=0.001):
        """get_training_data's unit conversion, floors and precip accumulation + log transform, in place on the device"""
        self._ck(self.lib.sml_condition_series(self.h, period, precip_epsilon))

    def conditioning_stats(self, first_col, stride, ncols):
        """-> (mean (nloc, L), std (nloc, L), sst_bool_input (nloc,)) of every local region from the resident series"""
        Lmax = 36
        mean = np.zeros(self.num_of_regions_on_proc * Lmax)
        std = np.zeros(self.num_of_regions_on_proc * Lmax)
        flag = np.zeros(self.num_of_regions_on_proc, dtype=np.int32)
        L = self._ck(self.lib.sml_conditioning_stats(self.h, first_col, stride, ncols, _d(mean), _d(std), _i(flag)),
                     allow_positive=True)
        n = self.num_of_regions_on_proc
        return mean[:n * L].reshape(n, L), std[:n * L].reshape(n, L), flag.astype(bool)

    def train_global_release(self):
        self._ck(self.lib.sml_train_global_release(self.h))

    def train_solve(self, beta_res, beta_model=1.0, using_prior=True, prior_val=0.0):
        info = np.zeros(len(self._train_regions), dtype=np.int32)
        self._ck(self.lib.sml_train_solve(self.h, beta_res, beta_model, int(using_prior), prior_val, _i(info)))
        return info

    def train_solver_stats(self):
        """-> number of regions of the current wave solved by the Cholesky path (the rest fell back to LU)"""
        n = C.c_int()
        self._ck(self.lib.sml_train_solver_stats(self.h, C.byref(n)))
        return n.value

    def train_gram_get(self, region):
        d = self.dims[(self._train_kind, region)]
        N = d["n"] + d["S"]
        sxs = np.zeros((N, N), order="F")
        sxt = np.zeros((d["P"], N), order="F")
        self._ck(self.lib.sml_train_gram_get(self.h, region, _d(sxs), _d(sxt)))
        return sxs, sxt

    def train_stats(self):
        """-> dict(gram_flops_useful, gram_ms, stategen_ms, solve_ms) of the current wave"""
        v = [C.c_double() for _ in range(4)]
        self._ck(self.lib.sml_train_stats(self.h, *[C.byref(a) for a in v]))
        return dict(zip(("gram_flops_useful", "gram_ms", "stategen_ms", "solve_ms"), (a.value for a in v)))

    def train_stategen_route(self) -> str:
        """state-generation route of the last training phase"""
        return {0: "steps", 1: "kernel", 2: "ring"}.get(int(self.lib.sml_train_stategen_route(self.h)), "none")

    def dmma_probe(self) -> float:
        """FP64 tensor-core issue peak of this GPU, TFLOP/s, measured now"""
        v = C.c_double()
        self._ck(self.lib.sml_dmma_probe(self.h, C.byref(v)))
        return v.value

    def train_end(self):
        self._ck(self.lib.sml_train_end(self.h))

    def train_set_overlap(self, on: bool):
        """state generation overlapping the previous slab's Gram (default off); applies from the next train_begin"""
        self._ck(self.lib.sml_train_set_overlap(self.h, 1 if on else 0))

    def train_trim(self):
        """return the device blocks kept from finished waves to the allocator"""
        self._ck(self.lib.sml_train_trim(self.h))

    # -- rolling_average_over_a_period_2d(grid, period)   mod_utilities.f90:1773
    def rolling_average_over_a_period_2d(self, grid, period, row0=0, nrows=None, keep_small=True):
        """in place on rows [row0, row0+nrows) of the Fortran-ordered (rows, time) array grid"""
        assert grid.dtype == np.float64 and grid.flags["F_CONTIGUOUS"] and grid.ndim == 2
        nrows = grid.shape[0] - row0 if nrows is None else nrows
        assert 0 <= row0 and row0 + nrows <= grid.shape[0]
        ptr = C.cast(grid.ctypes.data + 8 * row0, _dp)
        self._ck(self.lib.sml_rolling_average_2d(self.h, ptr, grid.shape[0], nrows, grid.shape[1], period, int(keep_small)))
        return grid

    # -- mldivide(A, B)   mod_linalg.f90:109
    def mldivide(self, A, B):
        """returns (X, info) like dgesv; A and B are not modified"""
        A = np.array(A, dtype=np.float64, order="F", copy=True)
        B = np.array(B, dtype=np.float64, order="F", copy=True)
        if A.shape[0] != B.shape[0]:
            return B, -1  # 'Cant compute solution returning A and B unchanged' (mod_linalg.f90:134-137)
        info = self.lib.sml_mldivide(self.h, _d(A), A.shape[0], _d(B), B.shape[0], A.shape[0], B.shape[1])
        if info < 0:
            raise EngineError(self.lib.sml_last_error(self.h).decode())
        return B, info

    # -- measurement
    def profile(self, on=True):
        self._ck(self.lib.sml_profile(self.h, int(on)))

    def kernel_times(self):
        """-> (sum of step-kernel ms, sum of finish-kernel ms, launches) since the last call"""
        a, b, c = C.c_double(), C.c_double(), C.c_int()
        self._ck(self.lib.sml_kernel_times(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def phase_times(self):
        """-> (sum of pack-kernel ms, sum of unpack-kernel ms, count) since the last call (profiling on)"""
        a, b, c = C.c_double(), C.c_double(), C.c_int()
        self._ck(self.lib.sml_phase_times(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def sync_times(self):
        """-> (ms of the update-only launches issued by synchronize while profiling, number of steps)"""
        a, n = C.c_double(), C.c_int64()
        self._ck(self.lib.sml_sync_times(self.h, C.byref(a), C.byref(n)))
        return a.value, n.value

    def step_chunk_rows(self, kind=ATMO):
        return int(self.lib.sml_step_chunk_rows(self.h, kind))

    def kernel_launch_count(self):
        return int(self.lib.sml_kernel_launch_count(self.h))

    def predict_algorithmic_bytes(self, kind=ATMO):
        return int(self.lib.sml_predict_algorithmic_bytes(self.h, kind))

    def update_algorithmic_bytes(self, kind=ATMO):
        return int(self.lib.sml_update_algorithmic_bytes(self.h, kind))
