#!/usr/bin/env python
"""bench.py -- headline benchmark of the SPEEDY-ML reservoir hot path on B200.

Workload (BASELINE.json configs[1]): the full 1152-region hybrid atmosphere forecast, T30 / 8 sigma
levels, reservoir size m=6000 (n = 5760..6160 per region class), degree 6, overlap 1, precip + SST +
TISR inputs, regions sharded over N GPUs exactly as processor_decomposition does.
A "step" is one hybrid 6-hour step: predict (state update + readout) for every region, the exchange of
the outvec slabs (peer stores over NVLink when N > 1), scatter into the global grids with the clamps, and the
rebuild of every region's feedback / local_model.  1 step = 0.25 sim-day.

  value  : sim-days per wall-second with everything resident in HBM (the host model's forecast grid F
           stays the one the last e2e step left on the device)
  e2e    : the same step through the reference-facing API with HOST buffers: sendrecievegrid's
           wholegrid copy-out (D2H), the host model stub, forecast + TISR copy-in (H2D), every step.  At N > 1
           every exchange -- outvec all-gather, forecast distribution -- runs inside the engine
           (sml_comm_bootstrap); this file issues no collective inside a step
  oracle_check  : the CPU oracle (checker only, outside the timed regions) reproduces the engine's state and outvec of three
                  regions per rank at every one of the 6 correctness steps (<= 1e-12)
  grid_checksum : FP64 sum and XOR of the bit patterns of the global grids after 6 sequential hybrid steps from a
           fixed start, on every rank: sharding does not change any region's arithmetic, so the value is the same
           for N = 1, 2, 4, 8 (and all ranks of a run must agree)
  train  : BASELINE's second metric at every N (regions sharded, no collective): aggregate Gram FP64 TFLOP/s
  --impl reference : the CPU oracle (port of the reference's algorithmic form) on the host cores, ALL 1152 regions.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_TOTAL = 1152
M_RES = 6000
SIM_DAYS_PER_STEP = 0.25
METRIC = "hybrid forecast sim-days/wall-sec (1152 regions)"
UNIT = "sim-days/s"
WORKLOAD = "full 1152-region hybrid atmosphere prediction, T30 8-sigma, reservoir 6000"
CHECK_STEPS = 6


def sst_input_mask(region: int) -> bool:
    return region % 10 < 7  # 70 % of the regions carry an SST input slot


def _synthetic():
    return importlib.import_module("speedy-ml_b200.synthetic")   # NumPy only; loads no native library


def engine_dims(region, sst_in):
    E = importlib.import_module("speedy-ml_b200.engine")
    return E.region_dims(R_TOTAL, region, 1, M_RES, 6.0, True, True, sst_in, False)


def gen_region(region: int, dims=None):
    """seeded synthetic weights of one region (seed = 20251018 + region), reference construction recipe.
    dims(region, sst_bool_input) -> dict(n, k, D, P, S, L): the engine's sizes by default (GPU arm, tests, tools); the CPU
    arm passes the oracle's so that it never loads the engine library"""
    syn = _synthetic()
    sst_in = sst_input_mask(region)
    d = (dims or engine_dims)(region, sst_in)
    rng = np.random.default_rng(20251018 + region)
    rows, cols, vals = syn.make_adjacency(d["n"], d["k"], rng, radius=0.7, power_iters=30)
    winc, wcol = syn.make_win_compact(d["n"], d["D"], rng, sigma=0.5)
    N = d["n"] + d["S"]
    wout = np.empty((d["P"], N), order="F")
    flat = wout.reshape(-1, order="F")
    flat[:] = (rng.random(flat.size) - 0.5) * (np.sqrt(12.0) / np.sqrt(N))  # unit-variance/sqrt(N)
    mean, std = syn.make_mean_std(d["L"], rng)
    return dict(region=region, sst_bool_input=sst_in, rows=rows, cols=cols, vals=vals, winc=winc, wcol=wcol,
                wout=wout, mean=mean, std=std, **d)


def initial_fields(seed=7):
    rng = np.random.default_rng(seed)
    clim4d, clim2d, tisr, base_sst, sea_mask = _synthetic().climatology(rng)
    return dict(clim4d=clim4d, clim2d=clim2d, tisr=tisr, base_sst=base_sst, sea_mask=sea_mask)


class HostStub:
    """deterministic stand-in for run_model/agcm_main (SPEEDY stays on the host and is out of scope):
    forecast = 0.98*grid + 0.02*climatology with run_model's q floor (src/mpires.f90:1648-1650).  Two roundings and an
    add per element -- bit for bit the arithmetic of the CPU arm's stub -- written into preallocated (pinned) arrays.
    The element-wise passes run through torch's CPU kernels (threaded over the host cores; NumPy's are single-threaded
    and cost 0.2 ms per step on the GPU boxes' hosts, more than a whole device step at 8 GPUs)."""

    def __init__(self, clim4d, clim2d, out4=None, out2=None):
        import torch
        self.torch = torch
        self.c4 = np.asfortranarray(0.02 * clim4d)
        self.c2 = np.asfortranarray(0.02 * clim2d)
        self.f4 = out4 if out4 is not None else np.empty((4, 96, 48, 8), order="F")
        self.f2 = out2 if out2 is not None else np.empty((96, 48), order="F")
        self.t_c4, self.t_c2 = torch.from_numpy(self.c4), torch.from_numpy(self.c2)
        self.t_f4, self.t_f2 = torch.from_numpy(self.f4), torch.from_numpy(self.f2)
        self.t_q = torch.from_numpy(self.f4.reshape(-1, order="F"))[3::4]     # view: var 4 (q) of every cell
        self._in = {}

    def _tensor(self, a):
        key = (a.ctypes.data, a.shape)
        t = self._in.get(key)
        if t is None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")          # read-only views of the engine's pinned staging
                t = self.torch.from_numpy(a)
            if len(self._in) > 8:
                self._in.clear()
            self._in[key] = t
        return t

    def __call__(self, w4d, w2d, wsst=None):
        torch = self.torch
        torch.mul(self._tensor(w4d), 0.98, out=self.t_f4)
        self.t_f4.add_(self.t_c4)
        torch.mul(self._tensor(w2d), 0.98, out=self.t_f2)
        self.t_f2.add_(self.t_c2)
        self.t_q.clamp_(min=0.000001)
        return self.f4, self.f2


def host_stub(w4d, w2d, clim4d, clim2d, out=None):
    """one-shot form of HostStub (tests and tools)"""
    return HostStub(clim4d, clim2d, *(out or (None, None)))(w4d, w2d)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = device_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the step kernel from the COMMITTED ncu --set full summary (static evidence, not
    something this run measured) -> (bytes or None, source label)"""
    for name in ("step_kernel_traffic_r02.json", "step_kernel_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                d = json.load(open(path))
                return d.get("traffic_bytes_per_launch"), f"profiles/{name} (static: committed ncu capture of kernel {d.get('kernel', '?')})"
            except Exception:
                continue
    return None, "none"


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ CPU arm
class CpuModel:
    """the oracle (reference algorithmic form: COO SpMV, DENSE n x D W_in GEMV as the reference stores it, dense W_out
    GEMV, un-standardise, gather with clamps, host stub, scatter + standardise) over a set of regions.  Only oracle/
    and NumPy are touched: the engine library is never loaded by this class."""

    def __init__(self, regions, nthreads):
        from oracle import oracle_c as oc
        self.oc, self.nthreads, self.regions = oc, nthreads, list(regions)
        self.full = len(self.regions) == R_TOTAL
        self.F = initial_fields()

        def dims(region, sst_in):
            rc = oc.Region(R_TOTAL, region, m=M_RES, precip_bool=True, sst_bool=True, sst_bool_input=sst_in)
            return dict(n=rc.n, k=rc.k, D=rc.D, P=rc.P, S=rc.S, L=rc.L)

        def build(r):
            w = gen_region(r, dims)
            rc = oc.Region(R_TOTAL, r, m=M_RES, precip_bool=True, sst_bool=True, sst_bool_input=w["sst_bool_input"])
            rc.set_weights(w["rows"], w["cols"], w["vals"], None, w["wout"], w["mean"], w["std"])
            rc.set_win_compact(w["winc"], w["wcol"])
            rc.densify_win()            # the dense win(n, D) of the reference, built inside the oracle
            return rc, w["mean"][-1], w["std"][-1]

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=max(1, min(32, nthreads))) as ex:
            built = list(ex.map(build, self.regions))
        self.regs = [b[0] for b in built]
        self.sst_mean = np.array([b[1] for b in built])
        self.sst_std = np.array([b[2] for b in built])
        self.setup_s = time.perf_counter() - t0
        F = self.F
        self.w4d, self.w2d = F["clim4d"].copy(order="F"), F["clim2d"].copy(order="F")
        self.wp = np.zeros((96, 48), order="F")
        self.wsst = np.maximum(F["base_sst"], 272.0)
        self.has = np.ones(len(self.regs), dtype=np.int32)
        self.oo = np.zeros((len(self.regs), 4))
        for i, r in enumerate(self.regions):
            xs, xe, ys, ye, *_ = oc.getxyresextent(R_TOTAL, r)
            self.oo[i] = F["base_sst"][xs - 1:xe, ys - 1:ye].ravel(order="F")
        # bytes the reference's form streams per region-step: dense W_in + W_out + COO adjacency
        self.bytes_per_step = sum(8 * rc.n * rc.D + 8 * rc.P * (rc.n + rc.S) + 16 * rc.k for rc in self.regs)

    def step(self):
        oc, F = self.oc, self.F
        oc.predict_all(self.regs, nthreads=self.nthreads)
        if self.full:   # closed loop: the gather covers the whole grid only when every region is there
            self.w4d, self.w2d, self.wp, self.wsst = oc.step_gather(self.regs, True, True, F["base_sst"], F["sea_mask"],
                                                                    ocean_out=self.oo, has_ocean=self.has)
        f4, f2 = oc.host_stub(self.w4d, self.w2d, F["clim4d"], F["clim2d"])
        oc.step_scatter(self.regs, True, True, False, self.w4d, self.w2d, self.wp, self.wsst, f4, f2, F["tisr"],
                        self.sst_mean, self.sst_std, nthreads=self.nthreads)

    def time_steps(self, warmup, steps):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return time.perf_counter() - t0


def cpu_sample_baseline(seconds_target: float, nthreads: int, nsample: int = 128):
    """cpu_baseline of the GPU arm: a bounded sample (every 9th region id keeps the class mix), scaled to the model"""
    regions = list(range(0, R_TOTAL, R_TOTAL // nsample))[:nsample]
    m = CpuModel(regions, nthreads)
    dt1 = m.time_steps(1, 1)
    steps = max(2, int(seconds_target / max(dt1, 1e-6)))
    dt = m.time_steps(0, steps)
    value = steps * len(regions) / dt / R_TOTAL * SIM_DAYS_PER_STEP
    return {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
            "blas": "none: plain C loops (axpy-order dense GEMV, COO SpMV), gcc -O2",
            "achieved_GBps": m.bytes_per_step * steps / dt / 1e9,
            "sample": f"{len(regions)} of 1152 regions (every {R_TOTAL // nsample}th id, m=6000, dense W_in as the "
                      f"reference stores it) x {steps} hybrid steps in {dt:.1f} s; scaled by 1152/{len(regions)}; predict + "
                      f"host stub + feedback/local_model rebuild (the gather needs every region: only in --impl reference), "
                      f"NetCDF excluded; the full-model figure is the --impl reference arm"}


def run_reference(args):
    """the reference's algorithmic form on the host cores, ALL 1152 regions, closed loop, honouring --steps/--warmup"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = os.cpu_count() or 1
    avail = mem_available_gb()
    regions = list(range(R_TOTAL))
    note = ""
    if avail is not None and avail < 48.0:   # 38 GB of dense W_in + W_out do not fit: say so instead of swapping
        regions = list(range(0, R_TOTAL, 9))
        note = f"; host has only {avail:.0f} GB available: {len(regions)} regions timed and scaled"
    m = CpuModel(regions, nthreads)
    dt = m.time_steps(args.warmup, args.steps)
    value = args.steps * len(regions) / dt / R_TOTAL * SIM_DAYS_PER_STEP
    info = {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
            "blas": "none: plain C loops (axpy-order dense GEMV, COO SpMV), gcc -O2; the path is memory-bound on the "
                    "26.5 MB dense W_in per region, so a vendor BLAS would not change the order of magnitude",
            "achieved_GBps": m.bytes_per_step * args.steps / dt / 1e9,
            "setup_s": round(m.setup_s, 1),
            "sample": f"{len(regions)} of 1152 regions x {args.steps} timed hybrid steps ({args.warmup} warm-up) in {dt:.1f} s: "
                      f"predict (COO SpMV + dense W_in GEMV + tanh + dense W_out GEMV) + gather with clamps + host stub + "
                      f"scatter/standardise, closed loop, NetCDF excluded{note}"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (R_TOTAL / len(regions)), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "regions": len(regions), "reservoir_m": M_RES, "degree": 6, "overlap": 1,
                       "sim_days_per_step": SIM_DAYS_PER_STEP,
                       "note": "CPU oracle port of the reference's algorithmic form on the host cores; the "
                               "Fortran/MPI/MKL reference cannot be built in this image (no f951, MPI, MKL, ARPACK, NetCDF)"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ update-only leg
def update_leg(eng, E, regions, peak, peak_src, steps=40, spin_steps=359):
    """The state update alone (synchronize / spin-up, src/mod_reservoir.f90:1354-1381): all regions driven by a
    synthetic input series, CUDA events around the launches inside the engine (sml_sync_times); algorithmic bytes = the
    update part of DESIGN.md section 4.1.  Two measurements:
      * one launch per step (SML_SYNC_KERNEL=steps, k_update_ring): every step streams the adjacency from HBM -- the
        HBM roofline applies;
      * `spin_up`: the engine's default, ONE launch for the whole time loop (k_sync_persist; spin_steps = 359 is the
        reference's synchronization_length/timestep at the headline config): a region's T steps run back to back on one
        SM, so its adjacency is re-read from L2, not HBM; the algorithmic rate may exceed the HBM peak and says so.
    Runs after the timed region; the model state it leaves is not used again."""
    rng = np.random.default_rng(5)
    by_D = {}
    for r in regions:
        D = eng.dims[(E.ATMO, r)]["D"]
        if D not in by_D:
            by_D[D] = np.asfortranarray(rng.standard_normal((D, max(steps, spin_steps))))
    inputs = [by_D[eng.dims[(E.ATMO, r)]["D"]] for r in regions]
    nbytes = eng.update_algorithmic_bytes()
    old = os.environ.get("SML_SYNC_KERNEL")
    os.environ["SML_SYNC_KERNEL"] = "steps"
    try:
        eng.synchronize_all(inputs, 3)
        eng.profile(True)
        eng.synchronize_all(inputs, steps)
        ms, n = eng.sync_times()
        eng.profile(False)
    finally:
        if old is None:
            os.environ.pop("SML_SYNC_KERNEL", None)
        else:
            os.environ["SML_SYNC_KERNEL"] = old
    ms /= max(1, n)
    achieved = nbytes / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": "update-only path of sml_synchronize, one launch per step (k_update_ring)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": int(nbytes), "kernel_ms_per_launch": ms, "launches_timed": int(n)}
    # the default path: the whole spin-up in one launch
    n0 = eng.kernel_launch_count()
    eng.synchronize_all(inputs, 8)
    one_launch = eng.kernel_launch_count() - n0 <= 2    # the spin-up kernel (+ the tile-major pack the first time)
    eng.profile(True)
    eng.synchronize_all(inputs, spin_steps)
    ms_total, n = eng.sync_times()
    eng.profile(False)
    per_step = ms_total / max(1, n)
    ach = nbytes / (per_step * 1e-3) / 1e9
    out["spin_up"] = {"kernel": "k_sync_persist (time loop inside the kernel, state in shared memory, adjacency re-read from L2)"
                      if one_launch else "step launches (the shard does not qualify for k_sync_persist)",
                      "steps": int(n), "launches": 1 if one_launch else int(n), "ms_total": ms_total, "ms_per_step": per_step,
                      "algorithmic_GBps": ach, "vs_hbm_peak": ach / peak, "bound": "l2+smem (on-chip reuse across steps)",
                      "speedup_vs_step_launches": ms / per_step if per_step > 0 else None}
    return out


# ------------------------------------------------------------------------------------------ training leg
def train_leg(eng, E, torch, dist, world, regions, ws_dims, cols=2000, discard=40, batch=98, long_k=True):
    """BASELINE's second metric: training Gram FP64 TFLOP/s on USEFUL flops N(N+1)K + 2PNK (configs[2]).
    Every rank trains ONE wave of its own regions (all of them up to 144: the per-GPU shard at N = 8) for one phase;
    no collective (the path shards by region).  The aggregate is the sum of the ranks' useful flops over the slowest
    rank's Gram time.  Rank 0 also runs one long-K phase (K = 37 920, the 26-year setting of SURVEY 8(d) config 3) on a
    small wave to show the rate holds when the accumulation is long.  Runs on the engine that held the forecast model:
    the solve overwrites its W_out, so this leg is last."""
    syn = _synthetic()
    wave = regions[:144]
    rng = np.random.default_rng(1)
    cache = {}

    def series(r, ncols):
        d = ws_dims[r]
        key = (d["D"], d["S"], ncols)
        if key not in cache:   # regions of one shape class share the synthetic series (the arithmetic does not care)
            cache[key] = (syn.ar1_series(d["D"], ncols, rng), np.asfortranarray(rng.standard_normal((d["S"], ncols))))
        return cache[key]

    tds = [series(r, cols)[0] for r in wave]
    ims = [series(r, cols)[1] for r in wave]
    eng.train_set_overlap(False)               # serial schedule: the Gram kernel is timed alone
    dmma_peak = eng.dmma_probe()
    eng.train_begin(wave, batch)
    eng.train_feed(tds, ims, discard)          # warm-up phase (also the first of two accumulated phases)
    st0 = eng.train_stats()
    eng.train_feed(tds, ims, discard)
    st = eng.train_stats()
    gram_ms = st["gram_ms"] - st0["gram_ms"]
    stategen_ms = st["stategen_ms"] - st0["stategen_ms"]
    flops = st["gram_flops_useful"] - st0["gram_flops_useful"]
    info = eng.train_solve(1e-3, 1.0, True, 0.0)
    st = eng.train_stats()
    solve_ms = st["solve_ms"]
    by_chol = eng.train_solver_stats()
    eng.train_end()
    out_long = None
    if long_k and int(os.environ.get("RANK", "0")) == 0:
        K_LONG, small = 37920, wave[:8]
        tl = [series(r, K_LONG)[0] for r in small]
        il = [series(r, K_LONG)[1] for r in small]
        eng.train_begin(small, batch)
        eng.train_feed(tl, il, discard)
        s1 = eng.train_stats()
        eng.train_end()
        out_long = {"regions": len(small), "columns": K_LONG, "gram_tflops": s1["gram_flops_useful"] / (s1["gram_ms"] * 1e-3) / 1e12,
                    "gram_ms": s1["gram_ms"], "stategen_ms": s1["stategen_ms"]}
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dgemm = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    # aggregate over the ranks: total useful flops / the slowest rank's Gram time
    if world > 1:
        t = torch.tensor([flops, gram_ms, stategen_ms, solve_ms, dmma_peak, float(len(wave)), float(max(info)), float(by_chol)],
                         dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        allv = torch.stack(allv).cpu().numpy()
        flops_total, gram_ms_max = float(allv[:, 0].sum()), float(allv[:, 1].max())
        stategen_ms, solve_ms = float(allv[:, 2].max()), float(allv[:, 3].max())
        per_rank = [float(f / (m * 1e-3) / 1e12) for f, m in zip(allv[:, 0], allv[:, 1])]
        peak_total = float(allv[:, 4].sum())
        nreg_total, info_max, chol_total = int(allv[:, 5].sum()), int(allv[:, 6].max()), int(allv[:, 7].sum())
    else:
        flops_total, gram_ms_max, per_rank, peak_total = flops, gram_ms, [flops / (gram_ms * 1e-3) / 1e12], dmma_peak
        nreg_total, info_max, chol_total = len(wave), int(max(info)), int(by_chol)
    tf = flops_total / (gram_ms_max * 1e-3) / 1e12
    return {"metric": "training Gram FP64 TFLOP/s", "value": tf, "unit": "TFLOP/s", "n_gpus": world,
            "workload": f"ridge training, {nreg_total} regions ({len(wave)} per GPU, one wave) x 1 phase x {cols} columns, m=6000 "
                        f"(N=5892..6292); regions sharded as processor_decomposition, no collective",
            "flops_counted": "useful: N(N+1)K + 2PNK (symmetric half + Y*R^T)",
            "schedule": "serial (sml_train_set_overlap(0)): the Gram kernel timed alone; aggregate = total useful flops / slowest rank",
            "gram_ms": gram_ms_max, "stategen_ms": stategen_ms, "solve_ms_per_region": solve_ms / max(1, len(wave)),
            "solve_ms_wave": solve_ms, "solve_info_max": info_max, "solved_by_cholesky": chol_total,
            "per_gpu_tflops": per_rank, "long_k": out_long,
            "roofline": {"bound": "tensor", "kernel": "k_syrk_dmma (FP64 DMMA)", "achieved": tf, "peak": peak_total,
                         "unit": "TFLOP/s", "frac": tf / peak_total,
                         "peak_source": "DMMA m8n8k4 issue peak measured in this run on every GPU (sml_dmma_probe), summed; "
                                        "FP64 is not in MEASURED_PEAKS.json",
                         "cublas_dgemm_8192_tflops_rank0": dgemm, "frac_of_cublas_dgemm_rank0": per_rank[0] / dgemm}}


# ------------------------------------------------------------------------------------------ GPU arm
class OracleCheck:
    """Per-step parity inside the bench run: before every correctness step the CPU oracle (oracle/, the checker) takes
    the engine's state, feedback and local_model of a few local regions, after the step its own predict must reproduce the
    engine's new state and outvec (<= 1e-12 relative, the tolerance of tests/test_fullsize_gpu.py).  Runs on every rank."""
    TOL = 1e-12

    def __init__(self, weights):
        self.regs, self.worst, self.err = {}, 0.0, None
        try:
            from oracle import oracle_c as oc
            for r, w in weights.items():
                rc = oc.Region(R_TOTAL, r, m=M_RES, precip_bool=True, sst_bool=True, sst_bool_input=w["sst_bool_input"])
                rc.set_weights(w["rows"], w["cols"], w["vals"], None, w["wout"], w["mean"], w["std"])
                rc.set_win_compact(w["winc"], w["wcol"])
                self.regs[r] = rc
        except Exception as e:   # the oracle library is test infrastructure: its absence must not break the measurement
            self.err = f"{type(e).__name__}: {e}"

    def before(self, eng):
        for r, rc in self.regs.items():
            rc.x[:] = eng.state_get(r)
            rc.feedback[:] = eng.feedback_get(r)
            rc.local_model[:] = eng.local_model_get(r)

    def after(self, eng):
        def rel(a, b):
            return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
        for r, rc in self.regs.items():
            rc.predict()
            self.worst = max(self.worst, rel(eng.state_get(r), rc.x), rel(eng.outvec_get(r), rc.outvec))

    def result(self, torch, dist, world):
        worst, n = self.worst, len(self.regs)
        if world > 1:
            t = torch.tensor([worst, float(n)], dtype=torch.float64, device="cuda")
            tmax, tsum = t.clone(), t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            worst, n = float(tmax[0]), int(tsum[1])
        if n == 0:
            return {"ok": None, "unavailable": self.err or "no regions"}
        return {"ok": bool(worst <= self.TOL), "max_rel_err": worst, "tol": self.TOL, "regions_checked": n, "steps": CHECK_STEPS,
                "what": "CPU oracle predict (state + outvec) from the engine's own state / feedback / local_model of the first, "
                        "middle and last region of every rank, every correctness step"}


def grid_checksum(eng, torch, dist, world):
    """FP64 sum + XOR of the bit patterns of the grids THIS rank assembled; all ranks must agree"""
    g = np.concatenate([a.ravel(order="F") for a in eng.grids_get()])
    s, x = float(np.sum(g)), int(np.bitwise_xor.reduce(g.view(np.uint64)))
    agree = True
    if world > 1:
        t = torch.tensor([x & 0xFFFFFFFF, x >> 32], dtype=torch.int64, device="cuda")
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        agree = all(bool(torch.equal(v, allv[0])) for v in allv)
    return {"sum": repr(s), "xor": f"{x:016x}", "steps": CHECK_STEPS, "ranks_agree": agree,
            "what": "wholegrid4d|wholegrid2d|precip|sst after 6 sequential hybrid steps (host stub in the loop) from the fixed start"}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    E = importlib.import_module("speedy-ml_b200.engine")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    # torchrun pins OMP_NUM_THREADS to 1; the root's host-model stand-in is the only host compute of a step and may use
    # the cores the other ranks leave idle (they only enqueue)
    if rank == 0:
        torch.set_num_threads(max(1, min(8, (os.cpu_count() or 8) // max(1, world) * 2)))

    # --emulate-world W (single process): this GPU carries rank 0's shard of a W-rank run, no exchange -- isolates
    # the per-rank step cost at that scale (diagnostic; not a bench line)
    eng_world = args.emulate_world if (world == 1 and args.emulate_world > 1) else world
    eng = E.Engine(number_of_regions=R_TOTAL, irank=rank, numprocs=eng_world, device=local_rank, sst_prescribed=True,
                   stream=stream)
    my_regions = eng.region_indices

    t_setup = time.perf_counter()
    ws_dims = {}
    # regions of this rank that are also stepped by the CPU oracle during the correctness steps (first / middle / last)
    check_ids = sorted({my_regions[0], my_regions[len(my_regions) // 2], my_regions[-1]})
    check_w = {}
    workers = max(1, min(16, (os.cpu_count() or 2) // max(1, world)))
    with ThreadPoolExecutor(max_workers=workers) as ex:
        batch = 4 * workers
        for i0 in range(0, len(my_regions), batch):
            for w in ex.map(gen_region, my_regions[i0:i0 + batch]):
                eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                                  win_compact=w["winc"], win_col=w["wcol"], D=w["D"],
                                  sst_bool_input=w["sst_bool_input"])
                ws_dims[w["region"]] = {k: w[k] for k in ("n", "k", "D", "P", "S", "L")}
                if w["region"] in check_ids:
                    check_w[w["region"]] = w
    eng.finalize()
    t_setup = time.perf_counter() - t_setup
    setup = eng.setup_stats()
    F = initial_fields()
    eng.set_sst_static(F["base_sst"], F["sea_mask"])
    eng.set_sst_prescribed(F["base_sst"])
    H = importlib.import_module("speedy-ml_b200.hybrid")
    H.check_contiguous_sharding(R_TOTAL, eng_world)
    shard = H.EngineShard(eng, torch)
    if world > 1 and args.peer:
        shard.bootstrap(dist)      # sml_comm_bootstrap: every exchange of the step now runs inside the engine
    stepper = H.HybridStepper(shard, rank=rank, world=world, dist=dist if world > 1 else None)
    stepper_ovl = H.HybridStepper(shard, rank=rank, world=world, dist=dist if world > 1 else None)
    lay = E.global_layout()
    # start from climatology: G holds the "previous hybrid grid", F its host forecast
    g0 = np.concatenate([F["clim4d"].ravel(order="F"), F["clim2d"].ravel(order="F"), np.zeros(96 * 48),
                         np.maximum(F["base_sst"], 272.0).ravel(order="F"), F["tisr"].ravel(order="F")])
    shard.G.copy_(torch.from_numpy(g0))
    f4, f2 = HostStub(F["clim4d"], F["clim2d"])(F["clim4d"], F["clim2d"])
    shard.F[:lay["f_total"]].copy_(torch.from_numpy(np.concatenate([f4.ravel(order="F"), f2.ravel(order="F")])))
    eng.step_unpack_device(1)
    torch.cuda.synchronize()

    # zero-copy host path: the grids are read from the engine's pinned staging and the stand-in model writes its
    # forecast straight into the pinned upload staging; the D2H / H2D transfers themselves are unchanged
    shard.zero_copy = True
    stub_out = shard.forecast_buffers(world)
    host_model = HostStub(F["clim4d"], F["clim2d"], stub_out[0], stub_out[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks and throttle reasons are sampled from here to the end of the timed regions (a 20-step region lasts ~30 ms,
    # too short on its own for nvidia-smi's sampling period).  nvidia-smi is started NOW and given time to initialise: its
    # start-up competes with the root's host thread and stretched a 20-step e2e region by 15 % when it fell inside it
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.6)

    # ---- correctness inside the scaling run: 6 sequential hybrid steps from the fixed start, checksum of the grids
    #      and, on EVERY rank, the CPU oracle stepping three of the rank's regions from the engine's own state and inputs
    #      (checker only, outside every timed region; the exchange itself is covered by the checksum all ranks must share)
    ocheck = OracleCheck(check_w)
    for t in range(1, CHECK_STEPS + 1):
        ocheck.before(eng)
        stepper.step(t, host_model, F["tisr"])
        ocheck.after(eng)
    barrier()
    checksum = grid_checksum(eng, torch, dist, world)
    oracle_check = ocheck.result(torch, dist, world)

    device_step = stepper.device_step

    def e2e_step(t):
        (stepper_ovl if args.overlap else stepper).step(t, host_model, F["tisr"])

    def timed(fn, steps, t0_index):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.perf_counter()
        ev0.record(stream)
        for i in range(steps):
            fn(t0_index + i)
        ev1.record(stream)
        barrier()
        wall = (time.perf_counter() - wall0) * 1e3
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, wall = tt.tolist()
        return ms, wall

    # ---- e2e first (it also leaves a consistent F on the device), then the device-resident measure
    if args.overlap:
        stepper_ovl.overlap = True
        eng.set_overlap(True)
    for i in range(args.warmup):
        e2e_step(CHECK_STEPS + i + 1)
    e2e_ms, e2e_wall = timed(e2e_step, args.steps, CHECK_STEPS + args.warmup + 1)
    # where the host side of an e2e step spends its time (a few extra steps after the timed region, clocked per section)
    e2e_sections = {}
    if world == 1 or shard.comm_ready:
        nb = max(10, args.steps // 2)
        st_b = stepper_ovl if args.overlap else stepper
        base = CHECK_STEPS + args.warmup + args.steps + 1
        for i in range(nb):
            st_b.timed_step(base + i, host_model, F["tisr"], e2e_sections)
        e2e_sections = {k: round(v / nb * 1e3, 4) for k, v in e2e_sections.items()}   # ms per step
    if args.overlap:
        eng.set_overlap(False)
        stepper_ovl.overlap = False
    for i in range(args.warmup):
        device_step(i + 1)
    launches0 = eng.kernel_launch_count()
    dev_ms, dev_wall = timed(device_step, args.steps, 1)       # the headline: nothing but the step's own launches in the stream
    launches = eng.kernel_launch_count() - launches0
    # per-kernel times from a SECOND pass of the same steps with the engine's CUDA-event brackets on (seven event records
    # per step sit between the kernels and cost the step ~3 %, so they stay out of the headline pass)
    eng.profile(True)
    prof_ms, _ = timed(device_step, args.steps, 1 + args.steps)
    k_step_ms, k_fin_ms, k_cnt = eng.kernel_times()
    pack_ms, unpack_ms, ph_cnt = eng.phase_times()
    eng.profile(False)
    clocks = sampler.stop() if sampler else None

    # sanity: the model state is finite
    x = eng.state_get(my_regions[0])
    ov = eng.outvec_get(my_regions[0])
    finite = bool(np.isfinite(x).all() and np.isfinite(ov).all())
    status = eng.grid_status()

    eng_peer = eng.peer_attached()
    if eng_peer:
        eng.peer_check()
    ms_per_step = dev_ms / args.steps
    value = SIM_DAYS_PER_STEP / (ms_per_step * 1e-3)
    e2e_value = SIM_DAYS_PER_STEP / (e2e_wall / args.steps * 1e-3)  # wall clock: host work is inside
    alg_bytes = eng.predict_algorithmic_bytes()
    peak, peak_src = measured_peak()
    step_kernel_ms = k_step_ms / max(1, k_cnt)
    achieved = alg_bytes / (step_kernel_ms * 1e-3) / 1e9
    plan = eng.step_plan()
    # the committed ncu capture is of the N=1 launch (all 1152 regions); per-rank launches are proportionally smaller
    traffic, traffic_src = ncu_traffic() if world == 1 else (None, "none (the committed capture is of the N=1 launch)")

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "regions": R_TOTAL, "regions_per_gpu": len(my_regions),
                       "reservoir_m": M_RES, "degree": 6, "overlap": 1, "sim_days_per_step": SIM_DAYS_PER_STEP,
                       "sharding": f"processor_decomposition over {world} rank(s)",
                       "exchange": ("none (single rank)" if world == 1 else
                                    "inside the engine (sml_comm_bootstrap): outvecs pushed by the readout-finish kernel, "
                                    "forecast block pushed by the root's k_peer_push, device-side flags; no host collective"
                                    if shard.comm_ready else "NCCL all_gather_into_tensor of the outvec slabs + NCCL broadcast"),
                       "step_kernel": plan,
                       "l2": "per-GPU weights streamed every step (8.1 GB / n_gpus) exceed the 126 MB L2; no flush",
                       "value_path": "device-resident: predict + all-gather + scatter/clamp + feedback rebuild; "
                                     "host model excluded (F resident)",
                       "e2e_path": "sml_predict + sml_step_exchange_begin (D2H) + host model stub + "
                                   "sml_step_exchange_end (H2D), wall clock; "
                                   + ("overlapped mode: the next predict's state update and x~ readout run while the "
                                      "host model works (SURVEY.md Appendix D)" if args.overlap else "sequential mode"),
                       "state_finite": finite, "grid_status_bits": status,
                       "setup_s": round(t_setup, 1), "setup_upload_s": round(setup["upload_s"], 2),
                       "setup_weight_arena": {"bytes": setup["arena_bytes"], "device_allocations": setup["arena_chunks"]},
                       **({"emulate_world": eng_world, "note": "DIAGNOSTIC: one rank's shard of an emulated "
                           f"{eng_world}-rank run, not a whole-model number"} if eng_world != world else {})},
            "grid_checksum": checksum,
            "oracle_check": oracle_check,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((lay["f_total"] + 96 * 48) * 8),
                    "d2h_bytes_per_step": int(lay["tisr"] * 8), "ms_per_step": e2e_wall / args.steps,
                    "ms_per_step_device_events": e2e_ms / args.steps,
                    "host_sections_ms_rank0": e2e_sections},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": f"{plan['kernel']} (fused state update + readout)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms_per_launch": step_kernel_ms,
                         "finish_kernel_ms_per_launch": k_fin_ms / max(1, k_cnt), "launches_timed": k_cnt,
                         "pack_kernel_ms_per_launch": pack_ms / max(1, ph_cnt),
                         "unpack_kernel_ms_per_launch": unpack_ms / max(1, ph_cnt),
                         "profiled_pass_ms_per_step": prof_ms / args.steps,
                         "other_ms_per_step": prof_ms / args.steps - (step_kernel_ms + (k_fin_ms + pack_ms + unpack_ms) / max(1, k_cnt)),
                         "note": "kernel times come from a second pass with CUDA-event brackets between the kernels; "
                                 "ms_per_step / value come from the un-instrumented pass"},
            "clocks": clocks,
            "wall_ms_per_step": dev_wall / args.steps,
        }
        if world == 1:   # also in the emulated-shard diagnostic: the update-only kernel at 144 regions per GPU
            line["update_roofline"] = update_leg(eng, E, my_regions, peak, peak_src)
    if not args.no_train and eng_world == world:
        tr = train_leg(eng, E, torch, dist, world, my_regions, ws_dims, long_k=(world == 1))
        if rank == 0:
            line["train"] = tr
    if world > 1:
        dist.barrier()      # nobody frees its exchange block while a peer may still push into it
    eng.close()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample_baseline(args.cpu_seconds, os.cpu_count() or 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--emulate-world", type=int, default=0)
    ap.add_argument("--no-peer", dest="peer", action="store_false",
                    help="multi-GPU: NCCL host collectives instead of the exchange inside the engine")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="e2e in the sequential (reference-order) mode instead of the overlapped one")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
