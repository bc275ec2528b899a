"""The hybrid step loop of src/parallelmain.f90:206-273 over one engine shard per rank.

`HybridStepper` is the host-side sequencing the reference spreads over parallelmain.f90 (region loop calling
predict, :226-251) and mpires.f90:sendrecievegrid (:218-804): predict on every local region, exchange of the
outvec slabs, assembly of the global grids with the clamps, the host model (SPEEDY's run_model, a callable
here) on rank 0, and the rebuild of every region's feedback / local_model.

Ranks are one process per GPU.  The data path has ONE exchange per step -- the all-gather of the outvec
slabs (1152 x 136 doubles in total) -- plus the distribution of rank 0's host-model forecast and the date's TISR
field; every rank then rebuilds the whole grid and gathers its own halo'd inputs locally.  Regions are sharded
as processor_decomposition (src/res_domain.f90:31-62) does; with number_of_regions divisible by the rank count
the shards are contiguous ascending blocks, so rank-major all-gather order IS region order (`slab_rows`).

Two transports, same arithmetic (bit-identical grids):
  * library (default on a GPU box, `EngineShard.bootstrap`): after sml_comm_bootstrap the engine does the whole
    exchange itself over NVLink peer stores -- outvecs from the readout-finish kernel, ocean slabs and the root's
    forecast block from push kernels -- and this file issues NO collective: every rank just calls
    predict / exchange_begin / exchange_end, exactly like the Fortran shim's sendrecievegrid.
  * host collectives (`dist` given, not bootstrapped): NCCL / gloo all-gather and broadcast on the engine's buffers
    (the CPU tests with the oracle-backed shard, and the A/B of the fused path).

The stepper only needs the small `shard` protocol below, so the multi-rank sequencing is testable on CPU with
the gloo backend (tests/test_multirank_gloo.py drives it with an oracle-backed shard); on a GPU box the shard
is `EngineShard`, a thin view of speedy-ml_b200.engine.Engine whose buffers are engine-owned device memory.

shard protocol:  predict() . ocean_predict() . pack(t) . unpack(t) . exchange_begin(t) -> (w4d, w2d, wp, wsst)
                 exchange_end(t, f4d, f2d, tisr) . load_forecast(f4d, f2d, tisr) [rank 0, stages H2D]
                 tensors: slab, gathered, F, tisr_dev, (ocean_slab, ocean_gathered) or None
"""
from __future__ import annotations

import numpy as np

from . import engine as E


def slab_rows(number_of_regions: int, numprocs: int):
    """region id of every row of the all-gathered outvec buffer (rank-major order)."""
    rows = []
    for r in range(numprocs):
        rows += E.processor_decomposition(r, numprocs, number_of_regions)
    return rows


def check_contiguous_sharding(number_of_regions: int, numprocs: int):
    """the engine indexes the gathered buffer by region id: require rank-major order == region order"""
    rows = slab_rows(number_of_regions, numprocs)
    if rows != list(range(number_of_regions)):
        raise ValueError(f"{number_of_regions} regions over {numprocs} ranks is not a contiguous sharding "
                         "(processor_decomposition puts the remainder regions out of order)")
    return number_of_regions // numprocs


def ocean_step_due(t: int, timestep: int = 6, timestep_slab: int = 168) -> bool:
    """src/parallelmain.f90:238: the ocean reservoirs step when mod(t*timestep, timestep_slab) == 0"""
    return (t * timestep) % timestep_slab == 0


class EngineShard:
    """the stepper's view of one Engine (one rank's regions on one B200)"""

    def __init__(self, eng: "E.Engine", torch_module, ocean: bool = False):
        self.eng = eng
        torch = torch_module
        # the host collectives below are ordered against torch's CURRENT stream only, so the engine must launch on it:
        # otherwise an all-gather would read outvecs the readout has not written yet, silently
        eng.set_stream(torch.cuda.current_stream())
        self._stream = torch.cuda.current_stream()
        self.comm_ready = False
        bufs = eng.exchange_buffers()
        self.slab = torch.as_tensor(bufs["outvec_slab"], device="cuda")
        self.gathered = torch.as_tensor(bufs["gathered"], device="cuda")
        self.F = torch.as_tensor(bufs["F"], device="cuda")
        self.G = torch.as_tensor(bufs["G"], device="cuda")
        lay = E.global_layout()
        self.lay = lay
        self.tisr_dev = self.G[lay["tisr"]:]
        self.ocean_slab = self.ocean_gathered = None
        if ocean:
            ob = eng.ocean_exchange_buffers()
            self.ocean_slab = torch.as_tensor(ob["ocean_slab"], device="cuda")
            self.ocean_gathered = torch.as_tensor(ob["ocean_gathered"], device="cuda")
        self._pin_f = torch.empty(lay["f_total"], dtype=torch.float64).pin_memory()
        self._pin_t = torch.empty(E.XGRID * E.YGRID, dtype=torch.float64).pin_memory()
        self._torch = torch
        self._pin_f_np = self._pin_f.numpy()
        self._pin_t_np = self._pin_t.numpy()
        self.reuse_grids = False    # True: exchange_begin returns the same host arrays every step
        self.zero_copy = False      # True: exchange_begin returns views of the pinned staging (no host copy)

    @property
    def peer_attached(self):
        return self.eng.peer_attached()

    def bootstrap(self, dist):
        """sml_comm_bootstrap over torch.distributed: afterwards the engine runs the whole multi-rank exchange itself"""
        def allgather_bytes(blob):
            out = [None] * dist.get_world_size()
            dist.all_gather_object(out, blob)
            return out
        self.eng.comm_bootstrap(allgather_bytes)
        self.comm_ready = True

    def attach_peers(self, dist):
        """exchange the CUDA IPC handles of the ranks' exchange blocks and switch the atmosphere all-gather to
        peer stores from the readout kernel (ranks of one node)"""
        mine = self.eng.peer_export()
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, mine)
        self.eng.peer_attach(handles)
        dist.barrier()

    def predict(self):
        self.eng.predict()

    def ocean_predict(self):
        self.eng.predict(kind=E.OCEAN)

    def pack(self, t):
        self.eng.step_pack_device(t)

    def unpack(self, t):
        self.eng.step_unpack_device(t)

    def exchange_device(self, t):
        self.eng.step_exchange_device(t)

    def exchange_begin(self, t, copy_out=True):
        if not copy_out:
            return self.eng.step_exchange_begin(t, copy_out=False)   # only enqueues the grid assembly
        if self.zero_copy:
            return self.eng.step_exchange_begin_view(t)      # views of the engine's pinned staging
        return self.eng.step_exchange_begin(t, reuse=self.reuse_grids)

    def forecast_buffers(self, world):
        """(forecast_4d, forecast_2d) arrays the host model may write in place: the engine's pinned upload staging on a
        single rank, this shard's pinned broadcast staging otherwise"""
        if world == 1 or self.comm_ready:
            return self.eng.forecast_staging()[:2]
        lay = self.lay
        return (self._pin_f_np[:lay["w2d"]].reshape((4, E.XGRID, E.YGRID, E.ZGRID), order="F"),
                self._pin_f_np[lay["w2d"]:].reshape((E.XGRID, E.YGRID), order="F"))

    def exchange_end(self, t, f4d, f2d, tisr):
        self.eng.step_exchange_end(t, f4d, f2d, tisr)

    def load_forecast(self, f4d, f2d, tisr):
        """rank 0, multi-rank runs: stage the host model's output for the broadcast (pinned -> device, async)"""
        lay = self.lay
        # the pinned staging is free: the previous step's copy out of it precedes this step's grid assembly in
        # stream order, and exchange_begin has just waited for the copy-out of those grids
        if not np.shares_memory(f4d, self._pin_f_np):        # else: the host model wrote into the staging in place
            np.copyto(self._pin_f_np[:lay["w2d"]], np.asarray(f4d).reshape(-1, order="F"))
            np.copyto(self._pin_f_np[lay["w2d"]:], np.asarray(f2d).reshape(-1, order="F"))
        self.F.copy_(self._pin_f, non_blocking=True)
        if tisr is not None:
            np.copyto(self._pin_t_np, np.asarray(tisr).reshape(-1, order="F"))
            self.tisr_dev.copy_(self._pin_t, non_blocking=True)

    # overlapped mode
    def set_overlap(self, on):
        self.eng.set_overlap(on)

    def set_tisr(self, tisr):
        self.eng.set_tisr(tisr)

    def predict_ahead(self, t):
        self.eng.step_predict_ahead(t)


class HybridStepper:
    """one hybrid step = parallelmain's region loop + sendrecievegrid, over `world` ranks"""

    def __init__(self, shard, rank: int = 0, world: int = 1, dist=None, timestep: int = 6, timestep_slab: int = 168,
                 overlap: bool = False):
        self.s, self.rank, self.world, self.dist = shard, rank, world, dist
        self.timestep, self.timestep_slab = timestep, timestep_slab
        if world > 1 and dist is None:
            raise ValueError("multi-rank stepping needs torch.distributed")
        # overlap: the next step's state update and x~ readout run while the host model works (Appendix D)
        self.overlap = overlap
        if overlap:
            self.s.set_overlap(True)

    # ---- the exchange of the outvec slabs (the path's only data collective)
    def _gather_outvecs(self, ocean_stepped: bool):
        if self.world == 1:
            return
        if not getattr(self.s, "peer_attached", False):   # else: pushed by the readout kernel over NVLink
            self.dist.all_gather_into_tensor(self.s.gathered, self.s.slab)
        if ocean_stepped and self.s.ocean_slab is not None:
            self.dist.all_gather_into_tensor(self.s.ocean_gathered, self.s.ocean_slab)

    def _predict(self, t):
        self.s.predict()
        stepped = self.s.ocean_slab is not None and ocean_step_due(t, self.timestep, self.timestep_slab)
        if stepped:
            self.s.ocean_predict()
        return stepped or (t == 1 and self.s.ocean_slab is not None)  # first step publishes the seeded ocean outvecs

    def step(self, t: int, host_model, tisr_grid):
        """the reference-facing step: host buffers, the host model (run_model) between begin and end.
        host_model(w4d, w2d, wsst) -> (forecast_4d, forecast_2d); tisr_grid is the date's global TISR field."""
        if self.world > 1 and getattr(self.s, "comm_ready", False):
            return self._step_library(t, host_model, tisr_grid)
        stepped = self._predict(t)
        self._gather_outvecs(stepped)
        if self.overlap:
            self.s.set_tisr(tisr_grid)          # every rank holds the TISR table (get_tisr_by_date)
        if self.rank == 0:
            w4d, w2d, wp, wsst = self.s.exchange_begin(t)   # overlap: also launches the next predict's ML part
            f4d, f2d = host_model(w4d, w2d, wsst)
            if self.world == 1:
                self.s.exchange_end(t, f4d, f2d, None if self.overlap else tisr_grid)
                return (w4d, w2d, wp, wsst)
            self.s.load_forecast(f4d, f2d, None if self.overlap else tisr_grid)
        else:
            self.s.pack(t)
            if self.overlap:
                self.s.predict_ahead(t)
            w4d = w2d = wp = wsst = None
        self.dist.broadcast(self.s.F, 0)
        if not self.overlap:
            self.dist.broadcast(self.s.tisr_dev, 0)
        self.s.unpack(t)
        return (w4d, w2d, wp, wsst) if self.rank == 0 else None

    def timed_step(self, t, host_model, tisr_grid, acc):
        """step() with the host-side sections clocked into acc (a dict of seconds): where the root's chain spends its
        time -- predict call, TISR upload, exchange_begin (grid assembly + copy-out wait), the host model, exchange_end"""
        import time
        c = time.perf_counter
        t0 = c()
        self._predict(t)
        t1 = c()
        if self.overlap:
            self.s.set_tisr(tisr_grid)
        t2 = c()
        root = self.rank == 0
        grids = self.s.exchange_begin(t) if root else self.s.exchange_begin(t, copy_out=False)
        t3 = c()
        f4d = f2d = None
        if root:
            f4d, f2d = host_model(grids[0], grids[1], grids[3])
        t4 = c()
        if root:
            self.s.exchange_end(t, f4d, f2d, None if self.overlap else tisr_grid)
        else:
            self.s.exchange_end(t, None, None, None)
        t5 = c()
        for k, v in (("predict", t1 - t0), ("set_tisr", t2 - t1), ("exchange_begin", t3 - t2), ("host_model", t4 - t3),
                     ("exchange_end", t5 - t4)):
            acc[k] = acc.get(k, 0.0) + v
        return grids

    def _step_library(self, t, host_model, tisr_grid):
        """multi-rank step with the exchange inside the engine (after EngineShard.bootstrap): the call sequence of
        parallelmain.f90:226-262 on every rank, no collective here.  Ranks other than the root never wait on the host."""
        self._predict(t)
        if self.overlap:
            self.s.set_tisr(tisr_grid)
        if self.rank == 0:
            grids = self.s.exchange_begin(t)
            f4d, f2d = host_model(grids[0], grids[1], grids[3])
            self.s.exchange_end(t, f4d, f2d, None if self.overlap else tisr_grid)
            return grids
        self.s.exchange_begin(t, copy_out=False)
        self.s.exchange_end(t, None, None, None)
        return None

    def device_step(self, t: int):
        """everything resident on the device; the forecast buffer F keeps what the last host step left"""
        stepped = self._predict(t)
        if self.world > 1 and not getattr(self.s, "comm_ready", False):
            self._gather_outvecs(stepped)
            self.s.pack(t)
            self.s.unpack(t)
        elif hasattr(self.s, "exchange_device"):
            self.s.exchange_device(t)        # scatter + gather in one cooperative launch
        else:
            self.s.pack(t)
            self.s.unpack(t)
