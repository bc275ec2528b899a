"""NumPy/SciPy restatement of the SPEEDY-ML reservoir hot path (secondary oracle).

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and by nothing under speedy-ml_b200/).
PARITY UNPINNED by the reference (no runnable golden vectors; see speedyml_oracle.h).

Written independently of oracle/speedyml_oracle.c, in "Fortran slice" style: global arrays are
held as order='F' ndarrays and sliced the way the reference slices them, so the two restatements
check each other (tests/test_oracle_*.py).  Citations are relative to /root/reference.
All public index values are 1-based like the Fortran source.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

XGRID, YGRID, ZGRID = 96, 48, 8  # src/mod_utilities.f90:17-20


# --------------------------------------------------------------------------- index arithmetic
def domaindecomposition(numregions: int):
    """src/res_domain.f90:258-280"""
    n = (XGRID * YGRID) // numregions
    factor_max = int(math.floor(math.sqrt(float(n))))
    for i in range(factor_max, 0, -1):
        if YGRID % i == 0:
            factory = i
            if n % factory == 0:
                factorx = n // factory
                if XGRID % factorx == 0:
                    return factorx, factory
    raise ValueError(f"domaindecomposition: {numregions} regions would reach MOD(ygrid,0)")


def getxyresextent(num_regions: int, region: int):
    """src/res_domain.f90:123-141 with getworkerlower_leftcorner :282-292"""
    fx, fy = domaindecomposition(num_regions)
    col = region % (YGRID // fy)
    row = int(math.floor(region / (YGRID / fy)))
    return (row * fx + 1, (row + 1) * fx, col * fy + 1, (col + 1) * fy, fx, fy)


def get_z_res_extent(num_vert_levels: int, vert_level: int):
    """src/res_domain.f90:143-153"""
    zc = ZGRID // num_vert_levels
    return ((vert_level - 1) * zc + 1, vert_level * zc, zc)


def getoverlapindices(numregions: int, region: int, overlap: int):
    """src/res_domain.f90:155-204 -> (ixs, ixe, iys, iye, ixc, iyc, pole, periodic)"""
    xs, xe, ys, ye, xc, yc = getxyresextent(numregions, region)
    ixc, iyc = xc + 2 * overlap, yc + 2 * overlap
    periodic = pole = False
    if xs - overlap < 1:
        ixs, periodic = XGRID - overlap + 1, True
    else:
        ixs = xs - overlap
    if xe + overlap > XGRID:
        ixe, periodic = overlap, True
    else:
        ixe = overlap + xe
    if ys - overlap < 1:
        iys, iyc, pole = 1, yc + overlap + (ys - 1), True
    else:
        iys = ys - overlap
    if ye + overlap > YGRID:
        iye, iyc, pole = YGRID, yc + overlap + (YGRID - ye), True
    else:
        iye = overlap + ye
    return ixs, ixe, iys, iye, ixc, iyc, pole, periodic


def getoverlapindices_vert(num_vert_levels: int, vert_level: int, vert_overlap: int):
    """src/res_domain.f90:206-256 -> (izs, ize, izc, top, bottom)"""
    zs, ze, zc = get_z_res_extent(num_vert_levels, vert_level)
    top, bottom = zs == 1, ze == ZGRID
    if zs - vert_overlap >= 1 and ze + vert_overlap <= ZGRID:
        return zs - vert_overlap, ze + vert_overlap, zc + 2 * vert_overlap, top, bottom
    if zs - vert_overlap < 1:
        return 1, ze + vert_overlap, zc + vert_overlap + (zs - 1), top, bottom
    return zs - vert_overlap, ZGRID, zc + vert_overlap + (ZGRID - ze), top, bottom


def get_trainingdataindices(num_regions: int, region: int, overlap: int):
    """src/res_domain.f90:547-574"""
    _, _, ys, ye, _, _ = getxyresextent(num_regions, region)
    _, _, _, _, ixc, iyc, _, _ = getoverlapindices(num_regions, region, overlap)
    xstart, xend = 1 + overlap, ixc - overlap
    if ys - overlap < 1:
        ystart, yend = 1 + (ys - 1), iyc - overlap
    elif ye + overlap > YGRID:
        ystart, yend = 1 + overlap, iyc - (YGRID - ye)
    else:
        ystart, yend = 1 + overlap, iyc - overlap
    return xstart, xend, ystart, yend


def get_trainingdataindices_vert(num_vert_levels: int, vert_level: int, vert_overlap: int):
    """src/res_domain.f90:576-600"""
    zs, ze, _ = get_z_res_extent(num_vert_levels, vert_level)
    _, _, izc, _, _ = getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap)
    if zs - vert_overlap < 1:
        return 1 + (zs - 1), izc - vert_overlap
    if ze + vert_overlap > ZGRID:
        return 1 + vert_overlap, izc - (ZGRID - ze)
    return 1 + vert_overlap, izc - vert_overlap


def processor_decomposition(irank: int, numprocs: int, number_of_regions: int):
    """src/res_domain.f90:31-62"""
    per, left_over = divmod(number_of_regions, numprocs)
    if irank >= left_over + 1 and irank > 0:
        return [per * irank + i - 1 for i in range(1, per + 1)]
    if irank == 0:
        return [i - 1 for i in range(1, per + 1)]
    return [per * irank + i - 1 for i in range(1, per + 1)] + [number_of_regions - left_over + irank - 1]


# --------------------------------------------------------------------------- tilers
def _xslice(numregions, region, overlap):
    """global x indices (1-based) in local order, src/res_domain.f90:380-418"""
    xs, xe, *_ = getxyresextent(numregions, region)
    ixs, ixe, _, _, ixc, _, _, periodic = getoverlapindices(numregions, region, overlap)
    if periodic and (xe > ixe or ixs > xs):
        gx = list(range(ixs, XGRID + 1)) + list(range(1, ixe + 1))
    else:
        gx = list(range(ixs, ixe + 1))
    assert len(gx) == ixc
    return np.asarray(gx) - 1


def tileoverlapgrid4d(grid4d, numregions, region, overlap, nvl=1, level=1, vov=0):
    """src/res_domain.f90:348-420; grid4d is (nv,96,48,8) order='F'"""
    _, _, iys, iye, *_ = getoverlapindices(numregions, region, overlap)
    izs, ize, *_ = getoverlapindices_vert(nvl, level, vov)
    gx = _xslice(numregions, region, overlap)
    return np.asfortranarray(grid4d[:, gx][:, :, iys - 1:iye][:, :, :, izs - 1:ize])


def tileoverlapgrid2d(grid2d, numregions, region, overlap):
    """src/res_domain.f90:484-545"""
    _, _, iys, iye, *_ = getoverlapindices(numregions, region, overlap)
    gx = _xslice(numregions, region, overlap)
    return np.asfortranarray(grid2d[gx][:, iys - 1:iye])


def tile_4d_and_logp_to_local_state_input(numregions, region, overlap, precip_bool, grid4d, grid2d, precip_grid):
    """src/res_domain.f90:1081-1125 (bottom level: logp and precip appended)"""
    parts = [tileoverlapgrid4d(grid4d, numregions, region, overlap).ravel(order="F"),
             tileoverlapgrid2d(grid2d, numregions, region, overlap).ravel(order="F")]
    if precip_bool:
        parts.append(tileoverlapgrid2d(precip_grid, numregions, region, overlap).ravel(order="F"))
    return np.concatenate(parts)


def tile_full_grid_with_local_state_vec_res1d(numregions, region, precip_bool, statevec, w4d, w2d, wprecip):
    """src/res_domain.f90:791-826 (single vertical level == bottom)"""
    xs, xe, ys, ye, xc, yc = getxyresextent(numregions, region)
    n4 = 4 * xc * yc * ZGRID
    w4d[:, xs - 1:xe, ys - 1:ye, :] = statevec[:n4].reshape((4, xc, yc, ZGRID), order="F")
    w2d[xs - 1:xe, ys - 1:ye] = statevec[n4:n4 + xc * yc].reshape((xc, yc), order="F")
    if precip_bool:
        wprecip[xs - 1:xe, ys - 1:ye] = statevec[n4 + xc * yc:n4 + 2 * xc * yc].reshape((xc, yc), order="F")


def tile_4d_and_logp_full_grid_to_local_res_vec(numregions, region, grid4d, grid2d):
    """src/res_domain.f90:1022-1053"""
    xs, xe, ys, ye, _, _ = getxyresextent(numregions, region)
    return np.concatenate([grid4d[:, xs - 1:xe, ys - 1:ye, :].ravel(order="F"),
                           grid2d[xs - 1:xe, ys - 1:ye].ravel(order="F")])


# --------------------------------------------------------------------------- region descriptor
@dataclass
class Region:
    """initializedomain + trained_reservoir_prediction + allocate_res_new, 1 vertical level.
    src/res_domain.f90:96-121, src/mod_reservoir.f90:80-180,1783-1886"""
    num_regions: int
    region: int
    overlap: int = 1
    m: int = 6000
    deg: float = 6.0
    precip_bool: bool = True
    sst_bool: bool = True          # model_parameters%slab_ocean_model_bool
    sst_bool_input: bool = True    # std(sst) > 0.2
    ml_only: bool = False
    leakage: float = 1.0
    # weights
    rows: np.ndarray = None
    cols: np.ndarray = None
    vals: np.ndarray = None
    win: np.ndarray = None
    wout: np.ndarray = None
    mean: np.ndarray = None
    std: np.ndarray = None
    # state
    x: np.ndarray = None
    feedback: np.ndarray = None
    local_model: np.ndarray = None
    outvec: np.ndarray = None
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        R, r, ov = self.num_regions, self.region, self.overlap
        (self.res_xstart, self.res_xend, self.res_ystart, self.res_yend,
         self.resxchunk, self.resychunk) = getxyresextent(R, r)
        (self.input_xstart, self.input_xend, self.input_ystart, self.input_yend,
         self.inputxchunk, self.inputychunk, self.pole, self.periodic) = getoverlapindices(R, r, ov)
        self.inputzchunk = self.reszchunk = ZGRID
        (self.tdata_xstart, self.tdata_xend, self.tdata_ystart, self.tdata_yend) = get_trainingdataindices(R, r, ov)
        if not self.sst_bool:
            self.sst_bool_input = False
        ixy, rxy = self.inputxchunk * self.inputychunk, self.resxchunk * self.resychunk
        msl = 4 * ZGRID
        msl += 1; self.logp_idx = msl
        msl += 1; self.tisr_idx = msl
        self.precip_idx = self.sst_idx = 0
        if self.precip_bool:
            msl += 1; self.precip_idx = msl
        if self.sst_bool:
            msl += 1; self.sst_idx = msl
        self.mean_std_length = msl
        precip_res = rxy if self.precip_bool else 0
        self.P = rxy * 4 * ZGRID + rxy + precip_res
        self.S = 0 if self.ml_only else rxy * 4 * ZGRID + rxy
        chunk = self.P
        locality = (ixy * ZGRID * 4 + ixy + (ixy if self.precip_bool else 0) + ixy
                    + (ixy if self.sst_bool_input else 0) - chunk)
        self.D = chunk + locality
        q = self.m / float(self.D)
        self.q = int(math.floor(q + 0.5))  # NINT
        self.n = self.q * self.D
        self.k = int((self.deg / float(self.m)) * self.n * self.n)
        self.atmo3d_end = 4 * ixy * ZGRID
        self.logp_start, self.logp_end = self.atmo3d_end + 1, self.atmo3d_end + ixy
        nxt = self.logp_end
        self.precip_start = self.precip_end = 0
        if self.precip_bool:
            self.precip_start, self.precip_end = nxt + 1, nxt + ixy
            nxt = self.precip_end
        self.sst_start = self.sst_end = 0
        if self.sst_bool_input:
            self.sst_start, self.sst_end = nxt + 1, nxt + ixy
            nxt = self.sst_end
        self.tisr_start, self.tisr_end = nxt + 1, nxt + ixy
        assert self.tisr_end == self.D


def coo_mv(reg: Region, x):
    """MKL_SPARSE_D_MV over the COO handle (src/mod_linalg.f90:10-25): duplicates sum"""
    y = np.zeros(reg.n)
    np.add.at(y, reg.rows - 1, reg.vals * x[reg.cols - 1])
    return y


def state_update(reg: Region, x, u):
    """src/mod_reservoir.f90:1444-1448"""
    y = coo_mv(reg, x)
    temp = reg.win @ u
    x_ = np.tanh(y + temp)
    return (1.0 - reg.leakage) * x + reg.leakage * x_


def synchronize(reg: Region, inputs, x, length):
    """src/mod_reservoir.f90:1354-1381"""
    for i in range(length):
        x = state_update(reg, x, inputs[:, i])
    return x


def unstandardize_state_vec_res(reg: Region, v):
    """src/res_domain.f90:1424-1475"""
    rx, ry = reg.resxchunk, reg.resychunk
    n4 = 4 * rx * ry * ZGRID
    t4 = v[:n4].reshape((4, rx, ry, ZGRID), order="F").copy()
    l = 0
    for i in range(4):
        for j in range(ZGRID):
            t4[i, :, :, j] = t4[i, :, :, j] * reg.std[l]
            t4[i, :, :, j] = t4[i, :, :, j] + reg.mean[l]
            l += 1
    out = v.copy()
    out[:n4] = t4.ravel(order="F")
    lp = out[n4:n4 + rx * ry] * reg.std[reg.logp_idx - 1]
    out[n4:n4 + rx * ry] = lp + reg.mean[reg.logp_idx - 1]
    if reg.precip_bool:
        pp = out[n4 + rx * ry:n4 + 2 * rx * ry] * reg.std[reg.precip_idx - 1]
        out[n4 + rx * ry:n4 + 2 * rx * ry] = pp + reg.mean[reg.precip_idx - 1]
    return out


def standardize_state_vec_res(reg: Region, v):
    """src/res_domain.f90:1270-1315"""
    rx, ry = reg.resxchunk, reg.resychunk
    n4 = 4 * rx * ry * ZGRID
    t4 = v[:n4].reshape((4, rx, ry, ZGRID), order="F").copy()
    l = 0
    for i in range(4):
        for j in range(ZGRID):
            t4[i, :, :, j] = (t4[i, :, :, j] - reg.mean[l]) / reg.std[l]
            l += 1
    out = v.copy()
    out[:n4] = t4.ravel(order="F")
    out[n4:n4 + rx * ry] = (v[n4:n4 + rx * ry] - reg.mean[l]) / reg.std[l]
    return out


def standardize_state_vec_input(reg: Region, fb):
    """src/res_domain.f90:1211-1268 on feedback(1:logp_end)"""
    ix, iy = reg.inputxchunk, reg.inputychunk
    t4 = fb[:reg.atmo3d_end].reshape((4, ix, iy, ZGRID), order="F").copy()
    l = 0
    for i in range(4):
        for j in range(ZGRID):
            t4[i, :, :, j] = (t4[i, :, :, j] - reg.mean[l]) / reg.std[l]
            l += 1
    out = fb.copy()
    out[:reg.atmo3d_end] = t4.ravel(order="F")
    out[reg.logp_start - 1:reg.logp_end] = (fb[reg.logp_start - 1:reg.logp_end] - reg.mean[l]) / reg.std[l]
    return out


def predict(reg: Region, x):
    """src/mod_reservoir.f90:1418-1489 (hybrid) / :1491-1535 (ml_only); returns (x, outvec)"""
    x = state_update(reg, x, reg.feedback)
    x_temp = x.copy()
    x_temp[1::2] = x_temp[1::2] ** 2  # Fortran 2:n:2
    if reg.ml_only:
        x_aug = x_temp
    else:
        x_aug = np.concatenate([reg.local_model, x_temp])
    outvec = reg.wout @ x_aug
    return x, unstandardize_state_vec_res(reg, outvec)


def step_gather(regs, precip_bool, ocean_model, base_sst, sea_mask, ocean_out=None, has_ocean=None):
    """src/mpires.f90:281-331,456-490"""
    w4d = np.zeros((4, XGRID, YGRID, ZGRID), order="F")
    w2d = np.zeros((XGRID, YGRID), order="F")
    wp = np.zeros((XGRID, YGRID), order="F")
    wsst = np.array(base_sst, order="F", copy=True) if ocean_model else np.zeros((XGRID, YGRID), order="F")
    for i, reg in enumerate(regs):
        tile_full_grid_with_local_state_vec_res1d(reg.num_regions, reg.region, precip_bool, reg.outvec, w4d, w2d, wp)
        if ocean_model:
            xs, xe, ys, ye, xc, yc = getxyresextent(reg.num_regions, reg.region)
            if has_ocean is not None and has_ocean[i]:
                wsst[xs - 1:xe, ys - 1:ye] = ocean_out[i][:xc * yc].reshape((xc, yc), order="F")
            else:
                wsst[xs - 1:xe, ys - 1:ye] = 272.0
    q = w4d[3]
    q[q < 0.000001] = 0.000001
    if ocean_model:
        wsst[sea_mask > 0.0] = base_sst[sea_mask > 0.0]
        wsst[wsst < 272.0] = 272.0
    if precip_bool:
        wp[wp < 0.00001] = 0.0
    return w4d, w2d, wp, wsst


def step_scatter(regs, precip_bool, ocean_model, w4d, w2d, wp, wsst, f4d, f2d, tisr_grid, sst_mean, sst_std):
    """src/mpires.f90:581-604,749-775"""
    for i, reg in enumerate(regs):
        R, r, ov = reg.num_regions, reg.region, reg.overlap
        fb = np.zeros(reg.D)
        head = tile_4d_and_logp_to_local_state_input(R, r, ov, precip_bool, w4d, w2d, wp)
        fb[:head.size] = head
        if not reg.ml_only:
            lm = tile_4d_and_logp_full_grid_to_local_res_vec(R, r, f4d, f2d)
            reg.local_model = standardize_state_vec_res(reg, lm)
        t = tileoverlapgrid2d(tisr_grid, R, r, ov).ravel(order="F")
        fb[reg.tisr_start - 1:reg.tisr_end] = (t - reg.mean[reg.tisr_idx - 1]) / reg.std[reg.tisr_idx - 1]
        if ocean_model and reg.sst_bool_input:
            s = tileoverlapgrid2d(wsst, R, r, ov).ravel(order="F")
            fb[reg.sst_start - 1:reg.sst_end] = (s - sst_mean[i]) / sst_std[i]
        fb = standardize_state_vec_input(reg, fb)
        if precip_bool:
            p = fb[reg.precip_start - 1:reg.precip_end]
            fb[reg.precip_start - 1:reg.precip_end] = (p - reg.mean[reg.precip_idx - 1]) / reg.std[reg.precip_idx - 1]
        reg.feedback = fb


def host_stub(w4d, w2d, clim4d, clim2d):
    """deterministic stand-in for agcm_main (not reference code): 0.98*grid + 0.02*climatology"""
    return 0.98 * w4d + 0.02 * clim4d, 0.98 * w2d + 0.02 * clim2d


# --------------------------------------------------------------------------- training
def tile_full_input_to_target_data(reg: Region, statevec):
    """src/res_domain.f90:602-651; statevec (D, T) -> (P, T)"""
    ix, iy = reg.inputxchunk, reg.inputychunk
    T = statevec.shape[1]
    t5 = statevec[:reg.atmo3d_end].reshape((4, ix, iy, ZGRID, T), order="F")
    xs, xe, ys, ye = reg.tdata_xstart, reg.tdata_xend, reg.tdata_ystart, reg.tdata_yend
    parts = [t5[:, xs - 1:xe, ys - 1:ye, :, :].reshape((-1, T), order="F")]
    t3 = statevec[reg.logp_start - 1:reg.logp_end].reshape((ix, iy, T), order="F")
    parts.append(t3[xs - 1:xe, ys - 1:ye, :].reshape((-1, T), order="F"))
    if reg.precip_bool:
        t3 = statevec[reg.precip_start - 1:reg.precip_end].reshape((ix, iy, T), order="F")
        parts.append(t3[xs - 1:xe, ys - 1:ye, :].reshape((-1, T), order="F"))
    return np.concatenate(parts, axis=0)


def train_hybrid(reg: Region, phases, batch_size, discard_cols, beta_res, beta_model, using_prior=True, prior_val=0.0):
    """train_reservoir's accumulation + fit (src/mod_reservoir.f90:289-316,1067-1175,1645-1701,1235-1334).
    phases = list of (trainingdata(D,T), imperfect(S,T)) already strided per phase (pre-noised).
    The state sequence is the plain recurrence (the hybrid path restarts each batch from the unsquared
    saved_state, :1142); batches only set which states are used and the summation order."""
    from scipy.linalg import lapack
    N = reg.n + reg.S
    sxs = np.zeros((N, N), order="F")
    sxt = np.zeros((reg.P, N), order="F")
    for td, im in phases:
        x = np.zeros(reg.n)
        for i in range(discard_cols):
            x = state_update(reg, x, td[:, i])
        TL = td.shape[1] - discard_cols
        nb = TL // batch_size
        states = np.zeros((reg.n, nb * batch_size))
        states[:, 0] = x
        for s in range(1, nb * batch_size):
            x = state_update(reg, x, td[:, discard_cols + s - 1])
            states[:, s] = x
        states[1::2, :] = states[1::2, :] ** 2
        for b in range(nb):
            c0 = discard_cols + b * batch_size
            aug = np.concatenate([im[:, c0:c0 + batch_size], states[:, b * batch_size:(b + 1) * batch_size]], axis=0)
            target = tile_full_input_to_target_data(reg, td[:, c0:c0 + batch_size])
            sxt += target @ aug.T
            sxs += aug @ aug.T
    d = np.arange(N)
    if using_prior:
        sxs[d[:reg.S], d[:reg.S]] += beta_model ** 2.0
        sxs[d[reg.S:], d[reg.S:]] += beta_res ** 2.0
        sxt[d[:reg.S], d[:reg.S]] += prior_val * beta_model ** 2.0
    else:
        sxs[d[:reg.S], d[:reg.S]] += beta_model
        sxs[d[reg.S:], d[reg.S:]] += beta_res
    _, _, xsol, info = lapack.dgesv(sxs.T.copy(order="F"), sxt.T.copy(order="F"))
    return np.asfortranarray(xsol.T), sxs, sxt, info


# --------------------------------------------------------------------------- slab ocean reservoir
@dataclass
class OceanRegion:
    """res%reservoir_special + res%grid_special: initialize_slab_ocean_model
    (src/mod_slab_ocean_reservoir.f90:9-133) and the offsets of trained_ocean_reservoir_prediction (:1598-1640)"""
    num_regions: int
    region: int
    overlap: int = 1
    m: int = 4000
    deg: float = 6.0
    precip_bool: bool = True   # only decides where the SST slot sits in the copied atmosphere mean/std
    nslots: int = 27           # timestep_slab/timestep - 1 (:810)
    rows: np.ndarray = None
    cols: np.ndarray = None
    vals: np.ndarray = None
    win: np.ndarray = None
    wout: np.ndarray = None
    mean: np.ndarray = None
    std: np.ndarray = None
    feedback: np.ndarray = None
    outvec: np.ndarray = None
    leakage: float = 1.0

    def __post_init__(self):
        R, r, ov = self.num_regions, self.region, self.overlap
        *_, self.resxchunk, self.resychunk = getxyresextent(R, r)
        (_, _, _, _, self.inputxchunk, self.inputychunk, _, _) = getoverlapindices(R, r, ov)
        self.tdata = get_trainingdataindices(R, r, ov)
        ixy, rxy = self.inputxchunk * self.inputychunk, self.resxchunk * self.resychunk
        self.P = 2 * rxy                       # sst_size_res + ohtc_res_size (:118)
        self.D = 4 * ixy + ixy + 3 * ixy       # atmo (lowest level) + logp + sst + tisr + ohtc (:122-127)
        self.n = int(math.floor(self.m / float(self.D) + 0.5)) * self.D
        self.k = int((self.deg / float(self.m)) * self.n * self.n)
        self.A = 5 * ixy                       # logp_end
        self.sst_start = self.A + 1
        self.sst_end = self.A + ixy
        self.tisr_start, self.tisr_end = self.sst_end + 1, self.sst_end + ixy
        self.ohtc_start, self.ohtc_end = self.tisr_end + 1, self.tisr_end + ixy
        self.sst_idx = 4 * ZGRID + 2 + (1 if self.precip_bool else 0) + 1  # after logp, tisr, (precip)
        self.mean_std_length = self.sst_idx
        self.ring = np.zeros((self.A, self.nslots), order="F")
        self.ml_only = True
        self.S = 0


def predict_slab_ml(reg: OceanRegion, x):
    """src/mod_slab_ocean_reservoir.f90:1318-1363; returns (x, outvec)"""
    x = state_update(reg, x, reg.feedback)
    x_temp = x.copy()
    x_temp[1::2] = x_temp[1::2] ** 2
    outvec = reg.wout @ x_temp
    return x, outvec * reg.std[reg.sst_idx - 1] + reg.mean[reg.sst_idx - 1]


def ocean_feedback(ocean: OceanRegion, atmo: Region, timestep: int, wsst):
    """SURVEY.md Appendix C intended semantics of src/mpires.f90:594-600,776-781"""
    ixy = ocean.inputxchunk * ocean.inputychunk
    a0 = atmo.atmo3d_end - 4 * ixy
    ocean.ring[:, (timestep - 1) % ocean.nslots] = atmo.feedback[a0:a0 + ocean.A]
    acc = np.zeros(ocean.A)
    for k in range(ocean.nslots):      # sum(ring, dim=2), slot order
        acc = acc + ocean.ring[:, k]
    ocean.feedback[:ocean.A] = acc / float(ocean.nslots)
    s = tileoverlapgrid2d(wsst, ocean.num_regions, ocean.region, ocean.overlap).ravel(order="F")
    ocean.feedback[ocean.sst_start - 1:ocean.sst_end] = (s - ocean.mean[ocean.sst_idx - 1]) / ocean.std[ocean.sst_idx - 1]


def tile_full_input_to_target_data_ocean(reg: OceanRegion, statevec):
    """src/res_domain.f90:691-728 (ohtc_prediction branch); statevec (D, T) -> (P, T)"""
    ix, iy = reg.inputxchunk, reg.inputychunk
    T = statevec.shape[1]
    xs, xe, ys, ye = reg.tdata
    parts = []
    for a, b in ((reg.sst_start, reg.sst_end), (reg.ohtc_start, reg.ohtc_end)):
        t3 = statevec[a - 1:b].reshape((ix, iy, T), order="F")
        parts.append(t3[xs - 1:xe, ys - 1:ye, :].reshape((-1, T), order="F"))
    return np.concatenate(parts, axis=0)


def train_ml(reg, phases, batch_size, discard_cols, beta_res, target_fn):
    """ML-only accumulation + fit: reservoir_layer_chunking_ml / chunking_matmul_ml / fit_chunk_ml
    (src/mod_reservoir.f90:963-1065,1594-1642,1177-1233; slab twins src/mod_slab_ocean_reservoir.f90:869-954,
    1420-1464,1061-1116).  Unlike the hybrid path, every batch after the first restarts its SpMV operand
    from the SQUARED copy states(:,batch_size) (:1034 / slab :933)."""
    from scipy.linalg import lapack
    n = reg.n
    sxs = np.zeros((n, n), order="F")
    sxt = np.zeros((reg.P, n), order="F")
    for td in phases:
        x = np.zeros(n)
        for i in range(discard_cols):
            x = state_update(reg, x, td[:, i])
        TL = td.shape[1] - discard_cols
        nb = TL // batch_size
        for b in range(nb):
            states = np.zeros((n, batch_size))
            for c in range(batch_size):
                s = b * batch_size + c
                if s == 0:
                    states[:, 0] = x
                    continue
                u = td[:, discard_cols + s - 1]
                if c == 0:
                    # operand = previous batch's last column AFTER the in-place squaring; the leak term uses x
                    y = coo_mv(reg, prev_last_sq)
                    xn = np.tanh(y + reg.win @ u)
                    x = (1.0 - reg.leakage) * x + reg.leakage * xn
                else:
                    x = state_update(reg, x, u)
                states[:, c] = x
            states[1::2, :] = states[1::2, :] ** 2
            prev_last_sq = states[:, -1].copy()
            c0 = discard_cols + b * batch_size
            target = target_fn(reg, td[:, c0:c0 + batch_size])
            sxt += target @ states.T
            sxs += states @ states.T
    d = np.arange(n)
    sxs[d, d] += beta_res
    _, _, xsol, info = lapack.dgesv(sxs.T.copy(order="F"), sxt.T.copy(order="F"))
    return np.asfortranarray(xsol.T), sxs, sxt, info


def sparse_eigen(n, rows, cols, vals, nev=6):
    """src/mod_linalg.f90:220-514: ARPACK dnaupd/dneupd (which='LM', nev=6) on the COO matrix, then
    eigs = maxval(d) over the whole (ncv,3) work array -- real parts, imaginary parts and residuals (:246,511).
    Restated with a dense eigen-decomposition (test sizes only): take the nev largest-magnitude eigenvalues and
    the max over their real and imaginary parts (residuals of converged pairs are ~0)."""
    A = np.zeros((n, n))
    np.add.at(A, (rows - 1, cols - 1), vals)      # duplicates sum, like the COO handle
    lam = np.linalg.eigvals(A)
    top = lam[np.argsort(-np.abs(lam))[:nev]]
    return float(max(top.real.max(), top.imag.max(), 0.0))


def gen_res_scale(vals, eigs, radius):
    """src/mod_reservoir.f90:193-195: newvals = (vals/eigs)*radius"""
    return (vals / eigs) * radius


def gaussian_noise_1d_function_precip(inputdata, gaussnoise, noisemag, precip_start, precip_end, mean_p, std_p, eps):
    """src/mod_utilities.f90:1410-1464 with the N(0,1) draws passed in (1-based inclusive precip range; 0 if none):
    noisy = x + g*noisemag*x everywhere except precip, which is un-standardised, taken back to linear space
    (eps*(e**t - 1)), noised, made non-negative, log-transformed and standardised again."""
    x = np.asarray(inputdata, dtype=np.float64)
    g = np.asarray(gaussnoise, dtype=np.float64)
    out = x + g * noisemag * x
    if precip_start > 0:
        sl = slice(precip_start - 1, precip_end)
        t = x[sl] * std_p + mean_p
        t = eps * (math.e ** t - 1)
        t = t + g[sl] * noisemag * t
        t = np.abs(t)
        t = np.log(1 + t / eps)
        out[sl] = (t - mean_p) / std_p
    return out


def total_precip_over_a_period(precip, period):
    """src/mod_utilities.f90:1688-1729; precip (..., T) hourly: out[t] = sum(copy[t-period : t]) (period+1 values, 1-based
    inclusive), sum(copy[1 : t]) while t - period < 1"""
    out = np.empty_like(precip)
    T = precip.shape[-1]
    for t in range(1, T + 1):
        lo = 1 if t - period < 1 else t - period
        acc = np.zeros(precip.shape[:-1])
        for k in range(lo, t + 1):          # Fortran sum over the slice, first to last
            acc = acc + precip[..., k - 1]
        out[..., t - 1] = acc
    return out


def rolling_average_over_a_period_2d(grid, period, keep_small=True):
    """src/mod_utilities.f90:1773-1815 (keep_small=False: the 3-D variant :1731-1771 with the leading axes merged).
    grid (rows, T); returns a copy.  As written: sum(copy(i,1:t))/t while t-period < 1, else sum(copy(i,t-period:t))/period
    -- period+1 values divided by period -- kept only when |sum| > 1e-7.  (The worked example in the subroutine's own
    comment assumes a window of `period` values and does not match the code; the code is what runs.)"""
    out = np.array(grid, dtype=np.float64, copy=True)
    T = grid.shape[-1]
    for t in range(1, T + 1):
        lo = 1 if t - period < 1 else t - period
        acc = np.zeros(grid.shape[:-1])
        for k in range(lo, t + 1):          # Fortran sum over the slice, first to last
            acc = acc + grid[..., k - 1]
        if t - period < 1:
            out[..., t - 1] = acc / t
        else:
            avg = acc / period
            out[..., t - 1] = np.where(np.abs(acc) > 0.0000001, avg, grid[..., t - 1]) if keep_small else avg
    return out


def condition_raw_series(w4d_t, tisr_t, precip_t, sst_t, period, eps):
    """get_training_data's conditioning (src/mod_reservoir.f90:362-395) on raw series (time last); returns copies"""
    w4d = w4d_t.copy()
    w4d[3] = w4d[3] * 1000.0
    w4d[3][w4d[3] < 0.000001] = 0.000001
    tisr = np.where(tisr_t < 0.0, 0.0, tisr_t)
    p = np.where(precip_t < 0.0, 0.0, precip_t)
    p = np.log(1 + total_precip_over_a_period(p, period) / eps)
    sst = np.where(sst_t < 272.0, 272.0, sst_t)
    return w4d, tisr, p, sst


def conditioning_stats(w4d_t, logp_t, tisr_t, precip_t, sst_t, numregions, region, overlap):
    """grid%mean / grid%std of one region from its training window (get_training_data, src/mod_reservoir.f90:413-470).
    Inputs are the conditioned global series, time last: w4d_t (4,96,48,8,T), 2-D fields (96,48,T); precip_t / sst_t may
    be None.  Slot order: (var, level) var-major, logp, tisr, [precip], [sst].  Formulas:
      standardize_data_5d_logp_tisr (src/mod_utilities.f90:1144-1193): mean = sum/size, std = sqrt(sum((x-mean)**2)/size)
      standardize_data_3d (:894-912, precip):  std = sqrt((sum(x**2) - sum(x)**2/size)/size)
      standardize_sst_data_3d (:853-892): two-pass std, kept only if sum(x**2)-sum(x)**2/size > 0 and std > 0.2,
                                          else mean = std = 0 and any_change = .False.
    Returns (mean, std, sst_any_change)."""
    T = w4d_t.shape[-1]

    def tile2(f):   # (96,48,T) -> halo block (ixc, iyc, T)
        return np.stack([tileoverlapgrid2d(f[:, :, t], numregions, region, overlap) for t in range(T)], axis=-1)

    def two_pass(x):
        m = np.sum(x) / x.size
        return m, math.sqrt(np.sum((x - m) ** 2) / x.size)

    loc4 = np.stack([tileoverlapgrid4d(w4d_t[..., t], numregions, region, overlap) for t in range(T)], axis=-1)
    mean, std = [], []
    for v in range(4):
        for z in range(ZGRID):
            m, sd = two_pass(loc4[v, :, :, z, :])
            mean.append(m); std.append(sd)
    for f in (logp_t, tisr_t):
        m, sd = two_pass(tile2(f))
        mean.append(m); std.append(sd)
    if precip_t is not None:
        x = tile2(precip_t)
        mean.append(np.sum(x) / x.size)
        std.append(math.sqrt((np.sum(x ** 2) - np.sum(x) ** 2 / x.size) / x.size))
    any_change = True
    if sst_t is not None:
        x = tile2(sst_t)
        m, sd = 0.0, 0.0
        any_change = False
        if (np.sum(x ** 2) - np.sum(x) ** 2 / x.size) > 0:
            m_, sd_ = two_pass(x)
            if sd_ > 0.2:
                m, sd, any_change = m_, sd_, True
        mean.append(m); std.append(sd)
    return np.array(mean), np.array(std), any_change


def pinv_svd(A, thres=1e-2):
    """src/mod_linalg.f90:27-107: Moore-Penrose via SVD, singular values <= thres zeroed"""
    U, s, VT = np.linalg.svd(A, full_matrices=False)
    sinv = np.where(s > thres, 1.0 / np.where(s > thres, s, 1.0), 0.0)
    return (VT.T * sinv) @ U.T


# ---- reservoir construction with the counter-based generator (second, independent restatement) --------------------
_M64 = (1 << 64) - 1
STREAM_WIN = 4096


def _mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def counter_bits(seed, region, stream, index):
    """the draw for (seed, region, stream, index): a pure function of the four numbers (splitmix64 finaliser)"""
    h1 = _mix64((seed ^ ((region & 0xFFFFFFFF) << 32) ^ (stream & 0xFFFFFFFF)) & _M64)
    return _mix64((h1 + index) & _M64)


def shuffle(n, returnsize, seed, region, stream):
    """src/mod_utilities.f90:1569-1596: the k-shuffle; `a` is a default real, so this = a*(n - n_chosen) + 1 is
    evaluated in single precision and truncated; the pick is clamped to the live range (guard, see the C oracle)"""
    choices = list(range(1, n + 1))
    out = np.zeros(returnsize, dtype=np.int32)
    for i in range(returnsize):
        a = np.float32(counter_bits(seed, region, stream, i) >> 40) * np.float32(1.0 / 16777216.0)
        live = n - i
        this = int(np.float32(np.float32(a * np.float32(live)) + np.float32(1.0)))
        this = min(this, live)
        tmp = choices[this - 1]
        out[i] = tmp
        choices[this - 1] = choices[live - 1]
        choices[live - 1] = tmp
    return out


def makesparse(n, k, seed, region):
    """src/mod_linalg.f90:180-218 -> rows, cols (1-based int32), vals"""
    vals = np.array([(counter_bits(seed, region, 0, e) >> 11) * (1.0 / 9007199254740992.0) for e in range(k)])
    rows, cols = np.zeros(k, dtype=np.int32), np.zeros(k, dtype=np.int32)
    if k > n:
        counter, leftover = k // n, k % n
        for i in range(counter):
            rows[i * n:(i + 1) * n] = shuffle(n, n, seed, region, 1 + 2 * i)
            cols[i * n:(i + 1) * n] = shuffle(n, n, seed, region, 2 + 2 * i)
        if leftover:
            rows[counter * n:] = shuffle(n, leftover, seed, region, 1 + 2 * counter)
            cols[counter * n:] = shuffle(n, leftover, seed, region, 2 + 2 * counter)
    else:
        rows[:] = shuffle(n, k, seed, region, 1)
        cols[:] = shuffle(n, k, seed, region, 2)
    return rows, cols, vals


def gen_win(n, D, sigma, seed, region):
    """src/mod_reservoir.f90:262-283 in the one-per-row form -> (winc[n], wcol[n] 0-based)"""
    q = n // D
    winc, wcol = np.zeros(n), np.zeros(n, dtype=np.int32)
    for i in range(D):
        for t in range(q):
            j = i * q + t
            rnd = (counter_bits(seed, region, STREAM_WIN, j) >> 11) * (1.0 / 9007199254740992.0)
            winc[j] = sigma * (-1.0 + 2.0 * rnd)
            wcol[j] = i
    return winc, wcol
