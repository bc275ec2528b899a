#!/usr/bin/env python
"""bench.py -- headline benchmark of the SPEEDY-ML reservoir hot path on B200.

Workload (BASELINE.json configs[1]): the full 1152-region hybrid atmosphere forecast, T30 / 8 sigma
levels, reservoir size m=6000 (n = 5760..6160 per region class), degree 6, overlap 1, precip + SST +
TISR inputs, regions sharded over N GPUs exactly as processor_decomposition does.
A "step" is one hybrid 6-hour step: predict (state update + readout) for every region, the exchange of
the outvec slabs (NCCL all-gather when N > 1), scatter into the global grids with the clamps, and the
rebuild of every region's feedback / local_model.  1 step = 0.25 sim-day.

  value  : sim-days per wall-second with everything resident in HBM (the host model's forecast grid F
           stays the one the last e2e step left on the device)
  e2e    : the same step through the reference-facing API with HOST buffers: sendrecievegrid's
           wholegrid copy-out (D2H), the host model stub, forecast + TISR copy-in (H2D), every step
  --impl reference : the CPU oracle (port of the reference's algorithmic form) on the host cores.

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
sys.path.insert(0, os.path.join(ROOT, "tests"))

R_TOTAL = 1152
M_RES = 6000
SIM_DAYS_PER_STEP = 0.25
METRIC = "hybrid forecast sim-days/wall-sec (1152 regions)"
UNIT = "sim-days/s"
WORKLOAD = "full 1152-region hybrid atmosphere prediction, T30 8-sigma, reservoir 6000"


def sst_input_mask(region: int) -> bool:
    return region % 10 < 7  # 70 % of the regions carry an SST input slot


def gen_region(region: int, dense_win: bool = False):
    """seeded synthetic weights of one region (seed = 20251018 + region), reference construction recipe"""
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    E = importlib.import_module("speedy-ml_b200.engine")
    sst_in = sst_input_mask(region)
    d = E.region_dims(R_TOTAL, region, 1, M_RES, 6.0, True, True, sst_in, False)
    rng = np.random.default_rng(20251018 + region)
    rows, cols, vals = syn.make_adjacency(d["n"], d["k"], rng, radius=0.7, power_iters=30)
    winc, wcol = syn.make_win_compact(d["n"], d["D"], rng, sigma=0.5)
    N = d["n"] + d["S"]
    wout = np.empty((d["P"], N), order="F")
    flat = wout.reshape(-1, order="F")
    flat[:] = (rng.random(flat.size) - 0.5) * (np.sqrt(12.0) / np.sqrt(N))  # unit-variance/sqrt(N)
    mean, std = syn.make_mean_std(d["L"], rng)
    w = dict(region=region, sst_bool_input=sst_in, rows=rows, cols=cols, vals=vals, winc=winc, wcol=wcol,
             wout=wout, mean=mean, std=std, **d)
    if dense_win:
        w["win"] = syn.win_dense_from_compact(winc, wcol, d["D"])
    return w


def initial_fields(seed=7):
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    rng = np.random.default_rng(seed)
    clim4d, clim2d, tisr, base_sst, sea_mask = syn.climatology(rng)
    return dict(clim4d=clim4d, clim2d=clim2d, tisr=tisr, base_sst=base_sst, sea_mask=sea_mask)


def host_stub(w4d, w2d, clim4d, clim2d, out=None):
    """deterministic stand-in for run_model/agcm_main (SPEEDY stays on the host and is out of scope):
    forecast = 0.98*grid + 0.02*climatology, with run_model's q floor.  out=(f4, f2) reuses two F-order arrays."""
    if out is None:
        f4 = 0.98 * w4d + 0.02 * clim4d
        f2 = 0.98 * w2d + 0.02 * clim2d
    else:
        f4, f2 = out
        np.multiply(w4d, 0.98, out=f4)
        f4 += 0.02 * clim4d
        np.multiply(w2d, 0.98, out=f2)
        f2 += 0.02 * clim2d
    np.maximum(f4[3], 0.000001, out=f4[3])
    return np.asfortranarray(f4), np.asfortranarray(f2)


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
    """dram bytes per launch of the step kernel from the committed ncu --set full summary, if any"""
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("traffic_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_run(seconds_target: float, nthreads: int, steps_fixed: int | None = None, nsample: int = 64):
    """times the oracle (reference algorithmic form: COO SpMV, DENSE n x D W_in GEMV, dense W_out GEMV,
    un-standardise, tile/standardise) on a bounded sample of regions.  Returns (sim_days_per_s, info)."""
    from oracle import oracle_c as oc

    # proportional class mix: every 18th region (64 regions) keeps interior/periodic/polar shares close
    regions = list(range(0, R_TOTAL, R_TOTAL // nsample))[:nsample]
    F = initial_fields()
    regs = []
    for r in regions:
        w = gen_region(r, dense_win=True)
        rc = oc.Region(R_TOTAL, r, m=M_RES, precip_bool=True, sst_bool=True, sst_bool_input=w["sst_bool_input"])
        rc.set_weights(w["rows"], w["cols"], w["vals"], w["win"], w["wout"], w["mean"], w["std"])
        regs.append(rc)
        del w
    sst_mean = np.array([rc.view("mean", (rc.L,))[-1] for rc in regs])
    sst_std = np.array([rc.view("std", (rc.L,))[-1] for rc in regs])
    w4d, w2d = F["clim4d"].copy(order="F"), F["clim2d"].copy(order="F")
    wp = np.zeros((96, 48), order="F")
    wsst = np.maximum(F["base_sst"], 272.0)

    def one_step():
        oc.predict_all(regs, nthreads=nthreads)
        f4, f2 = host_stub(w4d, w2d, F["clim4d"], F["clim2d"])
        oc.step_scatter(regs, True, True, False, w4d, w2d, wp, wsst, f4, f2, F["tisr"], sst_mean, sst_std,
                        nthreads=nthreads)

    one_step()  # warm-up (page in the weights)
    t0 = time.perf_counter()
    one_step()
    dt1 = time.perf_counter() - t0
    steps = steps_fixed if steps_fixed is not None else max(2, int(seconds_target / max(dt1, 1e-6)))
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    region_steps_per_s = steps * len(regs) / dt
    value = region_steps_per_s / R_TOTAL * SIM_DAYS_PER_STEP
    info = {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{len(regs)} of 1152 regions (every {R_TOTAL // nsample}th id, m=6000, dense W_in as the "
                      f"reference stores it) x {steps} hybrid steps in {dt:.1f} s; scaled by 1152/{len(regs)}; "
                      f"predict + feedback/local_model rebuild, host model stub included, NetCDF excluded"}
    return value, info, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = os.cpu_count() or 1
    per_step_budget = 20.0 / max(1, args.steps + args.warmup)
    value, info, ms = cpu_oracle_run(per_step_budget * args.steps, nthreads, steps_fixed=None)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * SIM_DAYS_PER_STEP / value, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference's algorithmic form; the "
                       "Fortran/MPI/MKL reference cannot be built in this image"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ update-only leg
def update_leg(eng, E, regions, peak, peak_src, steps=40):
    """The state update alone (synchronize / spin-up, src/mod_reservoir.f90:1354-1381): all regions driven by a
    synthetic input series for `steps` steps, CUDA events around every launch inside the engine (sml_sync_times);
    algorithmic bytes = the update part of DESIGN.md section 4.1.  Runs after the timed region; the model state it
    leaves is not used again."""
    rng = np.random.default_rng(5)
    inputs = [np.asfortranarray(rng.standard_normal((eng.dims[(E.ATMO, r)]["D"], steps))) for r in regions]
    eng.synchronize_all(inputs, 3)
    eng.profile(True)
    eng.synchronize_all(inputs, steps)
    ms, n = eng.sync_times()
    eng.profile(False)
    ms /= max(1, n)
    nbytes = eng.update_algorithmic_bytes()
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "k_update_sx<2> (state update alone: synchronize)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(nbytes),
            "kernel_ms_per_launch": ms, "launches_timed": int(n)}


# ------------------------------------------------------------------------------------------ training leg
def train_leg(E, torch, nreg=16, cols=2000, discard=40, batch=98):
    """BASELINE's second metric: training Gram FP64 TFLOP/s on USEFUL flops N(N+1)K + 2PNK (configs[2]), one
    wave of full-size regions x one phase, next to the solve time and the box's cuBLAS DGEMM rate."""
    syn = importlib.import_module("speedy-ml_b200.synthetic")
    eng = E.Engine(number_of_regions=R_TOTAL, irank=1, numprocs=R_TOTAL // nreg)
    regions = eng.region_indices
    ws = {}
    for r in regions:
        w = gen_region(r)
        ws[r] = w
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], None, w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"], S=w["S"], P=w["P"])
    eng.finalize()
    rng = np.random.default_rng(1)
    tds = [syn.ar1_series(ws[r]["D"], cols, rng) for r in regions]
    ims = [np.asfortranarray(rng.standard_normal((ws[r]["S"], cols))) for r in regions]
    eng.train_set_overlap(False)               # serial schedule: the Gram kernel is timed alone
    eng.train_begin(regions, batch)
    eng.train_feed(tds, ims, discard)          # warm-up phase (also the second of two accumulated phases)
    st0 = eng.train_stats()
    eng.train_feed(tds, ims, discard)
    st = eng.train_stats()
    gram_ms = st["gram_ms"] - st0["gram_ms"]
    flops = st["gram_flops_useful"] - st0["gram_flops_useful"]
    info = eng.train_solve(1e-3, 1.0, True, 0.0)
    st = eng.train_stats()
    eng.train_end()
    eng.close()
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
    tf = flops / (gram_ms * 1e-3) / 1e12
    dmma_peak = 37.1   # tools/dmma_probe.cu on this pool's B200 (profiles/dmma_probe_r01.txt)
    return {"metric": "training Gram FP64 TFLOP/s", "value": tf, "unit": "TFLOP/s",
            "workload": f"ridge training, {nreg} regions x 1 phase x {cols} columns, m=6000 (N=5892..6292)",
            "flops_counted": "useful: N(N+1)K + 2PNK (symmetric half + Y*R^T)",
            "schedule": "serial (sml_train_set_overlap(0)): the Gram kernel timed alone",
            "gram_ms": gram_ms, "stategen_ms": st["stategen_ms"] / 2, "solve_ms_per_region": st["solve_ms"] / nreg,
            "solve_info_max": int(max(info)),
            "roofline": {"bound": "tensor", "kernel": "k_syrk_dmma (FP64 DMMA)", "achieved": tf, "peak": dmma_peak,
                         "unit": "TFLOP/s", "frac": tf / dmma_peak,
                         "peak_source": "measured DMMA issue peak (tools/dmma_probe.cu); FP64 is not in MEASURED_PEAKS.json",
                         "cublas_dgemm_8192_tflops": dgemm, "frac_of_cublas_dgemm": tf / dgemm}}


# ------------------------------------------------------------------------------------------ GPU arm
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

    # --emulate-world W (single process): this GPU carries rank 0's shard of a W-rank run, no exchange -- isolates
    # the per-rank step cost at that scale (diagnostic; not a bench line)
    eng_world = args.emulate_world if (world == 1 and args.emulate_world > 1) else world
    eng = E.Engine(number_of_regions=R_TOTAL, irank=rank, numprocs=eng_world, device=local_rank, sst_prescribed=True,
                   stream=stream)
    my_regions = eng.region_indices
    t_gen = time.perf_counter()
    workers = max(1, min(16, (os.cpu_count() or 2) // max(1, world)))
    with ThreadPoolExecutor(max_workers=workers) as ex:
        batch = 4 * workers
        for i0 in range(0, len(my_regions), batch):
            for w in ex.map(gen_region, my_regions[i0:i0 + batch]):
                eng.region_upload(w["region"], w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"],
                                  win_compact=w["winc"], win_col=w["wcol"], D=w["D"],
                                  sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    t_gen = time.perf_counter() - t_gen
    F = initial_fields()
    eng.set_sst_static(F["base_sst"], F["sea_mask"])
    eng.set_sst_prescribed(F["base_sst"])
    H = importlib.import_module("speedy-ml_b200.hybrid")
    H.check_contiguous_sharding(R_TOTAL, eng_world)
    shard = H.EngineShard(eng, torch)
    if world > 1 and args.peer:
        shard.attach_peers(dist)   # fused all-gather: peer stores from the readout kernel over NVLink
    stepper = H.HybridStepper(shard, rank=rank, world=world, dist=dist if world > 1 else None)
    stepper_ovl = H.HybridStepper(shard, rank=rank, world=world, dist=dist if world > 1 else None)
    lay = E.global_layout()
    # start from climatology: G holds the "previous hybrid grid", F its host forecast
    g0 = np.concatenate([F["clim4d"].ravel(order="F"), F["clim2d"].ravel(order="F"), np.zeros(96 * 48),
                         np.maximum(F["base_sst"], 272.0).ravel(order="F"), F["tisr"].ravel(order="F")])
    shard.G.copy_(torch.from_numpy(g0))
    f4, f2 = host_stub(F["clim4d"], F["clim2d"], F["clim4d"], F["clim2d"])
    shard.F.copy_(torch.from_numpy(np.concatenate([f4.ravel(order="F"), f2.ravel(order="F")])))
    eng.step_unpack_device(1)
    torch.cuda.synchronize()

    # zero-copy host path: the grids are read from the engine's pinned staging and the stand-in model writes its
    # forecast straight into the pinned upload staging; the D2H / H2D transfers themselves are unchanged
    shard.zero_copy = True
    stub_out = shard.forecast_buffers(world)

    def host_model(w4d, w2d, wsst):
        # stand-in for run_model (SPEEDY stays on the host); same arithmetic as the CPU arm's stub
        return host_stub(w4d, w2d, F["clim4d"], F["clim2d"], out=stub_out)

    device_step = stepper.device_step

    def e2e_step(t):
        (stepper_ovl if args.overlap else stepper).step(t, host_model, F["tisr"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
        e2e_step(i + 1)
    e2e_ms, e2e_wall = timed(e2e_step, args.steps, args.warmup + 1)
    if args.overlap:
        eng.set_overlap(False)
        stepper_ovl.overlap = False
    for i in range(args.warmup):
        device_step(i + 1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.kernel_launch_count()
    eng.profile(True)
    dev_ms, dev_wall = timed(device_step, args.steps, 1)
    k_step_ms, k_fin_ms, k_cnt = eng.kernel_times()
    pack_ms, unpack_ms, ph_cnt = eng.phase_times()
    eng.profile(False)
    launches = eng.kernel_launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    # sanity: the model state is finite
    x = eng.state_get(my_regions[0])
    ov = eng.outvec_get(my_regions[0])
    finite = bool(np.isfinite(x).all() and np.isfinite(ov).all())

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
    # the committed ncu capture is of the N=1 launch (all 1152 regions); per-rank launches are proportionally smaller
    traffic = ncu_traffic() if world == 1 else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "regions": R_TOTAL, "regions_per_gpu": len(my_regions),
                       "reservoir_m": M_RES, "degree": 6, "overlap": 1, "sim_days_per_step": SIM_DAYS_PER_STEP,
                       "sharding": f"processor_decomposition over {world} rank(s)",
                       "exchange": ("none (single rank)" if world == 1 else
                                    "fused all-gather: peer stores from the readout kernel over NVLink (CUDA IPC)"
                                    if eng_peer else "NCCL all_gather_into_tensor of the outvec slabs"),
                       "l2": "per-GPU weights streamed every step (8.1 GB / n_gpus) exceed the 126 MB L2; no flush",
                       "value_path": "device-resident: predict + all-gather + scatter/clamp + feedback rebuild; "
                                     "host model excluded (F resident)",
                       "e2e_path": "sml_predict + sml_step_exchange_begin (D2H) + host model stub + "
                                   "sml_step_exchange_end (H2D), wall clock; "
                                   + ("overlapped mode: the next predict's state update and x~ readout run while the "
                                      "host model works (SURVEY.md Appendix D)" if args.overlap else "sequential mode"),
                       "state_finite": finite, "setup_s": round(t_gen, 1),
                       **({"emulate_world": eng_world, "note": "DIAGNOSTIC: one rank's shard of an emulated "
                           f"{eng_world}-rank run, not a whole-model number"} if eng_world != world else {})},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((lay["f_total"] + 96 * 48) * 8),
                    "d2h_bytes_per_step": int(lay["tisr"] * 8), "ms_per_step": e2e_wall / args.steps,
                    "ms_per_step_device_events": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_step (fused state update + readout)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms_per_launch": step_kernel_ms,
                         "finish_kernel_ms_per_launch": k_fin_ms / max(1, k_cnt), "launches_timed": k_cnt,
                         "pack_kernel_ms_per_launch": pack_ms / max(1, ph_cnt),
                         "unpack_kernel_ms_per_launch": unpack_ms / max(1, ph_cnt),
                         "chunk_rows": eng.step_chunk_rows()},
            "clocks": clocks,
            "wall_ms_per_step": dev_wall / args.steps,
        }
        if world == 1 and eng_world == 1:
            line["update_roofline"] = update_leg(eng, E, my_regions, peak, peak_src)
        if world == 1 and not args.no_cpu_baseline:
            _, info, _ = cpu_oracle_run(args.cpu_seconds, os.cpu_count() or 1)
            line["cpu_baseline"] = info
    eng.close()
    if rank == 0:
        if world == 1 and not args.no_train:
            line["train"] = train_leg(E, torch)
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
                    help="multi-GPU: NCCL all-gather of the outvec slabs instead of the fused peer-store exchange")
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
