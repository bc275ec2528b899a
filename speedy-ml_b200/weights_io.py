"""The per-region weight container of the reference, host side.

write_trained_res (src/mod_reservoir.f90:1703-1738) writes one NetCDF *classic* file per region and vertical level,
`worker_<region:04d>_level_<level>_<trial>.nc`, with seven variables -- win, wout (NF90_REAL, two-dimensional), rows,
cols (NF90_INT), vals, mean, std (NF90_REAL) -- each on its own dimensions (`win_x, win_y`, ...; src/mod_io.f90:1275-1320),
and read_trained_res (src/mod_io.f90:2938-2983) reads them back into double precision.  This module writes and reads
the same container with scipy's pure-Python classic NetCDF codec (no HDF5/NetCDF library): files produced by the
reference load here and in the engine (sml_region_upload_file has its own C++ reader), and files written here are
readable by the reference.  NetCDF-4/HDF5 copies must be converted first (`nccopy -k classic`).
"""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file


def trained_res_filename(region: int, trial_name: str, level: int = 1) -> str:
    return f"worker_{region:04d}_level_{level:d}_{trial_name}.nc"


def write_trained_res(path, win, wout, rows, cols, vals, mean, std):
    """win (n, D), wout (P, n+S): stored as float32 in Fortran element order, like nf90_put_var of the Fortran arrays"""
    win = np.asarray(win, dtype=np.float64)
    wout = np.asarray(wout, dtype=np.float64)
    with netcdf_file(path, "w", version=1) as f:
        def real2d(name, a):
            f.createDimension(name + "_x", a.shape[0])
            f.createDimension(name + "_y", a.shape[1])
            v = f.createVariable(name, "f4", (name + "_y", name + "_x"))   # file order: slowest first
            v.units = "unitless"
            v[:] = np.ascontiguousarray(a.T, dtype=np.float32)

        def vec(name, a, code):
            f.createDimension(name + "_x", len(a))
            v = f.createVariable(name, code, (name + "_x",))
            v.units = "unitless"
            v[:] = np.asarray(a).astype(code)

        real2d("win", win)
        real2d("wout", wout)
        vec("rows", rows, "i4")
        vec("cols", cols, "i4")
        vec("vals", vals, "f4")
        vec("mean", mean, "f4")
        vec("std", std, "f4")


def read_trained_res(path):
    """-> dict(win (n, D), wout (P, n+S), rows, cols int32, vals, mean, std) widened to float64"""
    with netcdf_file(path, "r", mmap=False) as f:
        g = {k: f.variables[k][:].copy() for k in ("win", "wout", "rows", "cols", "vals", "mean", "std")}
    return dict(win=np.asfortranarray(g["win"].T.astype(np.float64)), wout=np.asfortranarray(g["wout"].T.astype(np.float64)),
                rows=g["rows"].astype(np.int32), cols=g["cols"].astype(np.int32), vals=g["vals"].astype(np.float64),
                mean=g["mean"].astype(np.float64), std=g["std"].astype(np.float64))
