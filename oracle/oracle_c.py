"""ctypes binding of the C oracle (oracle/speedyml_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by speedy-ml_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libspeedyml_oracle.so")

_GRID_FIELDS = """number_of_regions region overlap num_vert_levels level_index vert_overlap
res_xstart res_xend res_ystart res_yend resxchunk resychunk res_zstart res_zend reszchunk
input_xstart input_xend input_ystart input_yend inputxchunk inputychunk input_zstart input_zend inputzchunk
tdata_xstart tdata_xend tdata_ystart tdata_yend tdata_zstart tdata_zend pole periodicboundary top bottom
atmo3d_start atmo3d_end logp_start logp_end precip_start precip_end sst_start sst_end tisr_start tisr_end
predict_start predict_end logp_mean_std_idx tisr_mean_std_idx precip_mean_std_idx sst_mean_std_idx
mean_std_length ohtc_start ohtc_end ohtc_mean_std_idx is_ocean""".split()

_DIMS_INT_FIELDS = """local_predictvars local_heightlevels_input local_heightlevels_res
logp_bool precip_bool precip_input_bool sst_bool sst_bool_input tisr_input_bool
logp_size_input sst_size_input precip_size_input tisr_size_input
logp_size_res precip_size_res sst_size_res tisr_size_res
chunk_size chunk_size_prediction chunk_size_speedy locality
m n k reservoir_numinputs nodes_per_input ml_only""".split()


class Grid(C.Structure):
    _fields_ = [(f, C.c_int) for f in _GRID_FIELDS]


class Dims(C.Structure):
    _fields_ = [(f, C.c_int) for f in _DIMS_INT_FIELDS] + [("deg", C.c_double), ("density", C.c_double),
                                                           ("leakage", C.c_double)]


def build(force: bool = False) -> str:
    """compile the oracle with the recipe in oracle/Makefile"""
    src = os.path.join(_HERE, "speedyml_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ})
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_region_new.restype = C.c_void_p
        L.orc_region_new.argtypes = [C.POINTER(Grid), C.POINTER(Dims)]
        L.orc_region_free.argtypes = [C.c_void_p]
        L.orc_region_ptr.restype = _dp
        L.orc_region_ptr.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_region_set_weights.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_region_set_leakage.argtypes = [C.c_void_p, C.c_double]
        L.orc_region_set_win_compact.argtypes = [C.c_void_p, _dp, _ip]
        L.orc_region_densify_win.argtypes = [C.c_void_p]
        L.orc_counter_bits.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_longlong]
        L.orc_counter_bits.restype = C.c_uint64
        L.orc_shuffle.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, _ip]
        L.orc_shuffle.restype = None
        L.orc_makesparse.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, _ip, _ip, _dp]
        L.orc_makesparse.restype = None
        L.orc_gen_win.argtypes = [C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int, _dp, _ip]
        L.orc_gen_win.restype = None
        L.orc_synchronize.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
        L.orc_predict.argtypes = [C.c_void_p, _dp]
        L.orc_predict_ml.argtypes = [C.c_void_p, _dp]
        L.orc_predict_all.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        L.orc_step_gather.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _ip,
                                      _dp, _dp, _dp, _dp]
        L.orc_step_scatter.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                       _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_run_model_clamp.argtypes = [_dp]
        L.orc_host_stub.argtypes = [_dp] * 6
        L.orc_coo_mv.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_dgesv.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _ip, _dp, C.c_int]
        L.orc_mldivide.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, C.c_int]
        L.orc_rolling_average_2d.argtypes = [_dp] + [C.c_int] * 5
        L.orc_rolling_average_2d.restype = None
        L.orc_train_init.argtypes = [C.c_void_p, C.c_int]
        L.orc_train_phase_hybrid.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int]
        L.orc_train_phase_ml.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int]
        L.orc_fit_chunk_hybrid.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_fit_chunk_ml.argtypes = [C.c_void_p, C.c_double]
        L.orc_train_free.argtypes = [C.c_void_p]
        L.orc_setup_region.argtypes = [C.c_int] * 7 + [C.c_double] + [C.c_int] * 4 + [C.POINTER(Grid), C.POINTER(Dims)]
        L.orc_setup_ocean_region.argtypes = [C.c_int] * 4 + [C.c_double, C.c_int, C.POINTER(Grid), C.POINTER(Dims)]
        L.orc_predict_slab_ml.argtypes = [C.c_void_p, _dp]
        L.orc_predict_slab.argtypes = [C.c_void_p, _dp]
        L.orc_ocean_feedback.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_int, C.c_int, _dp]
        L.orc_tile_full_input_to_target_data2d.argtypes = [C.POINTER(Grid), C.POINTER(Dims), _dp, C.c_int, C.c_int, _dp]
        L.orc_tileoverlapgrid4d.argtypes = [_dp] + [C.c_int] * 7 + [_dp]
        L.orc_tileoverlapgrid2d.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64) if not (isinstance(a, np.ndarray) and a.flags.f_contiguous and a.dtype == np.float64) else a


# ---- index functions -------------------------------------------------------------------------
def _ints(k):
    return [C.c_int() for _ in range(k)]


def domaindecomposition(numregions):
    fx, fy = C.c_int(), C.c_int()
    rc = lib().orc_domaindecomposition(numregions, C.byref(fx), C.byref(fy))
    if rc:
        raise ValueError("unsupported region count")
    return fx.value, fy.value


def getxyresextent(num_regions, region):
    v = _ints(6)
    if lib().orc_getxyresextent(num_regions, region, *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    return tuple(a.value for a in v)


def getoverlapindices(num_regions, region, overlap):
    v = _ints(8)
    if lib().orc_getoverlapindices(num_regions, region, overlap, *[C.byref(a) for a in v]):
        raise ValueError("unsupported region count")
    t = tuple(a.value for a in v)
    return t[:6] + (bool(t[6]), bool(t[7]))


def getoverlapindices_vert(nvl, level, vov):
    v = _ints(5)
    lib().orc_getoverlapindices_vert(nvl, level, vov, *[C.byref(a) for a in v])
    t = tuple(a.value for a in v)
    return t[:3] + (bool(t[3]), bool(t[4]))


def get_trainingdataindices(num_regions, region, overlap):
    v = _ints(4)
    lib().orc_get_trainingdataindices(num_regions, region, overlap, *[C.byref(a) for a in v])
    return tuple(a.value for a in v)


def get_trainingdataindices_vert(nvl, level, vov):
    v = _ints(2)
    lib().orc_get_trainingdataindices_vert(nvl, level, vov, *[C.byref(a) for a in v])
    return tuple(a.value for a in v)


def processor_decomposition(irank, numprocs, nregions):
    buf = (C.c_int * (nregions // numprocs + 2))()
    cnt = lib().orc_processor_decomposition(irank, numprocs, nregions, buf)
    return [buf[i] for i in range(cnt)]


def find_closest_divisor(target, number):
    return lib().orc_find_closest_divisor(target, number)


def tileoverlapgrid4d(grid4d, num_regions, region, overlap, nvl=1, level=1, vov=0):
    g = np.asfortranarray(grid4d, dtype=np.float64)
    nv = g.shape[0]
    *_, ixc, iyc, _, _ = getoverlapindices(num_regions, region, overlap)
    izc = getoverlapindices_vert(nvl, level, vov)[2]
    out = np.zeros((nv, ixc, iyc, izc), order="F")
    lib().orc_tileoverlapgrid4d(_d(g), nv, num_regions, region, overlap, nvl, level, vov, _d(out))
    return out


def tileoverlapgrid2d(grid2d, num_regions, region, overlap):
    g = np.asfortranarray(grid2d, dtype=np.float64)
    *_, ixc, iyc, _, _ = getoverlapindices(num_regions, region, overlap)
    out = np.zeros((ixc, iyc), order="F")
    lib().orc_tileoverlapgrid2d(_d(g), num_regions, region, overlap, _d(out))
    return out


# ---- region object ----------------------------------------------------------------------------
class Region:
    """one reservoir (reservoir_type + grid_type) held by the C oracle"""

    def __init__(self, num_regions, region, overlap=1, m=6000, deg=6.0, precip_bool=True, sst_bool=True,
                 sst_bool_input=True, ml_only=False, num_vert_levels=1, vert_level=1, vert_overlap=0):
        self.g, self.d = Grid(), Dims()
        rc = lib().orc_setup_region(num_regions, region, overlap, num_vert_levels, vert_level, vert_overlap, m,
                                    float(deg), int(precip_bool), int(sst_bool), int(sst_bool_input), int(ml_only),
                                    C.byref(self.g), C.byref(self.d))
        if rc:
            raise ValueError("orc_setup_region failed")
        self.h = C.c_void_p(lib().orc_region_new(C.byref(self.g), C.byref(self.d)))
        self.n, self.D = self.d.n, self.d.reservoir_numinputs
        self.P, self.S, self.k = self.d.chunk_size_prediction, self.d.chunk_size_speedy, self.d.k
        self.L = self.g.mean_std_length
        self.region, self.num_regions = region, num_regions

    def __del__(self):
        try:
            if self.h:
                lib().orc_region_free(self.h)
                self.h = None
        except Exception:
            pass

    def set_weights(self, rows, cols, vals, win, wout, mean, std):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        win = np.asfortranarray(win, dtype=np.float64) if win is not None else None
        wout = np.asfortranarray(wout, dtype=np.float64) if wout is not None else None
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        std = np.ascontiguousarray(std, dtype=np.float64)
        assert rows.size == self.k and cols.size == self.k and vals.size == self.k
        if win is not None:
            assert win.shape == (self.n, self.D)
        if wout is not None:
            assert wout.shape == (self.P, self.n + self.S)
        rc = lib().orc_region_set_weights(self.h, _i(rows), _i(cols), _d(vals), _d(win) if win is not None else None,
                                          _d(wout) if wout is not None else None, _d(mean), _d(std), mean.size)
        if rc:
            raise ValueError(f"orc_region_set_weights rc={rc}")

    def set_win_compact(self, winc, wcol):
        winc = np.ascontiguousarray(winc, dtype=np.float64)
        wcol = np.ascontiguousarray(wcol, dtype=np.int32)
        assert winc.size == self.n and wcol.size == self.n
        if lib().orc_region_set_win_compact(self.h, _d(winc), _i(wcol)):
            raise ValueError("win_col out of range")

    def densify_win(self):
        """compact one-per-row W_in -> the dense (n, D) matrix the reference stores; predict uses the dense GEMV again"""
        if lib().orc_region_densify_win(self.h):
            raise MemoryError("orc_region_densify_win")

    def set_leakage(self, leak):
        lib().orc_region_set_leakage(self.h, float(leak))

    def view(self, name, shape):
        p = lib().orc_region_ptr(self.h, name.encode())
        if not p:
            raise KeyError(name)
        size = int(np.prod(shape))
        return np.ctypeslib.as_array(p, shape=(size,)).reshape(shape, order="F")

    @property
    def x(self):
        return self.view("x", (self.n,))

    @property
    def feedback(self):
        return self.view("feedback", (self.D,))

    @property
    def local_model(self):
        return self.view("local_model", (max(self.S, 1),))[:self.S]

    @property
    def outvec(self):
        return self.view("outvec", (self.P,))

    @property
    def wout(self):
        return self.view("wout", (self.P, self.n + self.S))

    def synchronize(self, inputs, length=None):
        inputs = np.asfortranarray(inputs, dtype=np.float64)
        length = inputs.shape[1] if length is None else length
        lib().orc_synchronize(self.h, _d(inputs), inputs.shape[0], _d(self.x), length)

    def predict(self):
        if self.d.ml_only:
            lib().orc_predict_ml(self.h, _d(self.x))
        else:
            lib().orc_predict(self.h, _d(self.x))

    # training
    def train_init(self, batch_size):
        if lib().orc_train_init(self.h, batch_size):
            raise MemoryError
        self._N = self.n + self.S

    def train_phase(self, trainingdata, imperfect, discard_cols):
        td = np.asfortranarray(trainingdata, dtype=np.float64)
        if self.d.ml_only:
            lib().orc_train_phase_ml(self.h, _d(td), td.shape[0], td.shape[1], discard_cols)
        else:
            im = np.asfortranarray(imperfect, dtype=np.float64)
            lib().orc_train_phase_hybrid(self.h, _d(td), td.shape[0], _d(im), im.shape[0], td.shape[1], discard_cols)

    def fit(self, beta_res, beta_model=1.0, using_prior=True, prior_val=0.0):
        if self.d.ml_only:
            return lib().orc_fit_chunk_ml(self.h, beta_res)
        return lib().orc_fit_chunk_hybrid(self.h, beta_res, beta_model, int(using_prior), prior_val)

    @property
    def sxs(self):
        N = self.n + self.S
        return self.view("states_x_states_aug", (N, N))

    @property
    def sxt(self):
        return self.view("states_x_trainingdata_aug", (self.P, self.n + self.S))


class OceanRegion(Region):
    """res%reservoir_special / res%grid_special of one region (src/mod_slab_ocean_reservoir.f90)"""

    def __init__(self, num_regions, region, overlap=1, m=4000, deg=6.0, precip_bool=True, nslots=27, hybrid=False):
        self.g, self.d = Grid(), Dims()
        rc = lib().orc_setup_ocean_region(num_regions, region, overlap, m, float(deg), int(precip_bool),
                                          C.byref(self.g), C.byref(self.d))
        if rc:
            raise ValueError("orc_setup_ocean_region failed")
        self.hybrid = bool(hybrid)
        if self.hybrid:   # ml_only_ocean = .False.: the feature vector carries chunk_size_prediction model entries
            self.d.chunk_size_speedy = self.d.chunk_size_prediction
        self.h = C.c_void_p(lib().orc_region_new(C.byref(self.g), C.byref(self.d)))
        self.n, self.D = self.d.n, self.d.reservoir_numinputs
        self.P, self.S, self.k = self.d.chunk_size_prediction, self.d.chunk_size_speedy, self.d.k
        self.L = self.g.mean_std_length
        self.region, self.num_regions = region, num_regions
        self.A = self.g.logp_end
        self.nslots = nslots
        self.ring = np.zeros((self.A, nslots), order="F")  # averaged_atmo_input_vec, zeroed (:810-811)

    def predict(self):
        if self.hybrid:
            lib().orc_predict_slab(self.h, _d(self.x))      # predict_slab: local_model <- standardised outvec
        else:
            lib().orc_predict_slab_ml(self.h, _d(self.x))

    def build_feedback(self, atmo: Region, timestep: int, wholegrid_sst):
        sst = np.asfortranarray(wholegrid_sst, dtype=np.float64)
        lib().orc_ocean_feedback(self.h, atmo.h, _d(self.ring), self.nslots, timestep, _d(sst))

    def target(self, statevec):
        sv = np.asfortranarray(statevec, dtype=np.float64)
        out = np.zeros((self.P, sv.shape[1]), order="F")
        lib().orc_tile_full_input_to_target_data2d(C.byref(self.g), C.byref(self.d), _d(sv), sv.shape[0], sv.shape[1],
                                                   _d(out))
        return out


def _handles(regs):
    arr = (C.c_void_p * len(regs))()
    for i, r in enumerate(regs):
        arr[i] = r.h
    return arr


def predict_all(regs, ml_only=False, nthreads=1):
    lib().orc_predict_all(_handles(regs), len(regs), int(ml_only), nthreads)


def step_gather(regs, precip_bool, ocean_model, base_sst, sea_mask, ocean_out=None, has_ocean=None):
    w4d = np.zeros((4, 96, 48, 8), order="F")
    w2d = np.zeros((96, 48), order="F")
    wp = np.zeros((96, 48), order="F")
    wsst = np.zeros((96, 48), order="F")
    base_sst = np.asfortranarray(base_sst, dtype=np.float64) if base_sst is not None else wsst
    sea_mask = np.asfortranarray(sea_mask, dtype=np.float64) if sea_mask is not None else wsst
    oo = np.ascontiguousarray(ocean_out, dtype=np.float64) if ocean_out is not None else None
    ho = np.ascontiguousarray(has_ocean, dtype=np.int32) if has_ocean is not None else None
    lib().orc_step_gather(_handles(regs), len(regs), int(precip_bool), int(ocean_model), _d(base_sst), _d(sea_mask),
                          _d(oo) if oo is not None else None, _i(ho) if ho is not None else None,
                          _d(w4d), _d(w2d), _d(wp), _d(wsst))
    return w4d, w2d, wp, wsst


def step_scatter(regs, precip_bool, ocean_model, ml_only, w4d, w2d, wp, wsst, f4d, f2d, tisr_grid,
                 sst_mean, sst_std, nthreads=1):
    a = [np.asfortranarray(v, dtype=np.float64) for v in (w4d, w2d, wp, wsst, f4d, f2d, tisr_grid)]
    sm = np.ascontiguousarray(sst_mean, dtype=np.float64)
    ss = np.ascontiguousarray(sst_std, dtype=np.float64)
    lib().orc_step_scatter(_handles(regs), len(regs), int(precip_bool), int(ocean_model), int(ml_only),
                           *[_d(v) for v in a], _d(sm), _d(ss), nthreads)


def host_stub(w4d, w2d, clim4d, clim2d):
    f4d = np.zeros((4, 96, 48, 8), order="F")
    f2d = np.zeros((96, 48), order="F")
    a = [np.asfortranarray(v, dtype=np.float64) for v in (w4d, w2d, clim4d, clim2d)]
    lib().orc_host_stub(*[_d(v) for v in a], _d(f4d), _d(f2d))
    return f4d, f2d


def run_model_clamp(grid4d):
    assert grid4d.flags.f_contiguous
    lib().orc_run_model_clamp(_d(grid4d))


def makesparse(n, k, seed, region):
    """src/mod_linalg.f90:180-218 with the counter-based generator -> (rows, cols 1-based int32, vals U[0,1))"""
    rows, cols, vals = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32), np.zeros(k)
    lib().orc_makesparse(n, k, seed, region, _i(rows), _i(cols), _d(vals))
    return rows, cols, vals


def shuffle(n, returnsize, seed, region, stream):
    out = np.zeros(returnsize, dtype=np.int32)
    lib().orc_shuffle(n, returnsize, seed, region, stream, _i(out))
    return out


def gen_win(n, D, sigma, seed, region):
    """src/mod_reservoir.f90:262-283 -> (winc[n], wcol[n] 0-based)"""
    winc, wcol = np.zeros(n), np.zeros(n, dtype=np.int32)
    lib().orc_gen_win(n, D, float(sigma), seed, region, _d(winc), _i(wcol))
    return winc, wcol


def counter_bits(seed, region, stream, index):
    return int(lib().orc_counter_bits(seed, region, stream, index))


def coo_mv(n, rows, cols, vals, x):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(n)
    lib().orc_coo_mv(n, rows.size, _i(rows), _i(cols), _d(vals), _d(x), _d(y))
    return y


def rolling_average_over_a_period_2d(grid, period, keep_small=True):
    """src/mod_utilities.f90:1773-1815 on a (rows, T) array; returns a copy"""
    g = np.array(grid, dtype=np.float64, order="F", copy=True)
    lib().orc_rolling_average_2d(_d(g), g.shape[0], g.shape[0], g.shape[1], int(period), int(bool(keep_small)))
    return g


def mldivide(A, B):
    """returns (X, info); A (n,n), B (n,k)"""
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    B = np.array(B, dtype=np.float64, order="F", copy=True)
    info = lib().orc_mldivide(_d(A), A.shape[0], A.shape[1], _d(B), B.shape[0], B.shape[1])
    return B, info
