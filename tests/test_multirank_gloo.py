"""Multi-rank sequencing of the hybrid step on CPU: world_size 2, gloo backend, no GPU.

The product's HybridStepper (speedy-ml_b200/hybrid.py) is driven with an oracle-backed shard: each rank owns
processor_decomposition's block of regions, computes its outvec slab with the CPU oracle, the stepper
all-gathers the slabs, every rank rebuilds the global grids through the product's flattened scatter table
(sml_region_maps, host integer code of the C-ABI library -- no CUDA involved) and its own feedback vectors.
Rank 0 checks every step against a single-process oracle run of the same model: grids bit-identical."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import c_region, initial_grids, oc, region_weights

E = importlib.import_module("speedy-ml_b200.engine")
H = importlib.import_module("speedy-ml_b200.hybrid")

R, M = 1152, 300
NSTEPS = 3


def test_sharding_is_contiguous_for_power_of_two_ranks():
    for world in (1, 2, 4, 8):
        assert H.check_contiguous_sharding(R, world) == R // world
        assert H.slab_rows(R, world) == list(range(R))
    with pytest.raises(ValueError):     # 1152 = 5*230 + 2: remainder regions land out of order
        H.check_contiguous_sharding(R, 5)
    # the engine's table agrees with the oracle's restatement of processor_decomposition
    for world in (2, 3, 5, 8):
        for rank in range(world):
            assert E.processor_decomposition(rank, world, R) == oc.processor_decomposition(rank, world, R)


def test_ocean_step_schedule():
    due = [t for t in range(1, 60) if H.ocean_step_due(t)]
    assert due == [28, 56]              # mod(t*6, 168) == 0, src/parallelmain.f90:238


class OracleShard:
    """CPU stand-in for EngineShard: same protocol, arithmetic by the oracle, index work by the product library"""

    def __init__(self, rank, world, grids):
        self.ids = E.processor_decomposition(rank, world, R)
        self.ws = [region_weights(R, r, m=M, with_dense_win=False) for r in self.ids]
        self.regs = [c_region(w) for w in self.ws]
        self.P = self.ws[0]["P"]
        self.slab = torch.zeros(len(self.ids) * self.P, dtype=torch.float64)
        self.gathered = self.slab if world == 1 else torch.zeros(R * self.P, dtype=torch.float64)
        lay = E.global_layout()
        self.lay = lay
        self.G = np.zeros(lay["g_total"])
        self.F = torch.zeros(lay["f_total"], dtype=torch.float64)
        self.tisr_dev = torch.zeros(96 * 48, dtype=torch.float64)
        self.ocean_slab = self.ocean_gathered = None
        self.grids = grids
        # the product's scatter table for ALL regions (what sml_finalize uploads as out_dst)
        self.out_dst = np.concatenate([E.region_maps(R, r, 1, True, False)["output_map"] for r in range(R)])
        self.sst_mean = np.array([w["mean"][-1] for w in self.ws])
        self.sst_std = np.array([w["std"][-1] for w in self.ws])
        rng = np.random.default_rng(5)
        fbs = [(rng.standard_normal(576), rng.standard_normal(132)) for _ in range(R)]   # same stream on every rank
        for w, rc in zip(self.ws, self.regs):
            fb, lm = fbs[w["region"]]
            rc.feedback[:] = fb[:w["D"]]
            rc.local_model[:] = lm[:w["S"]]

    def predict(self):
        oc.predict_all(self.regs, nthreads=2)
        for i, rc in enumerate(self.regs):
            self.slab[i * self.P:(i + 1) * self.P] = torch.from_numpy(rc.outvec.copy())

    def pack(self, t):
        lay, G = self.lay, self.G
        G[:lay["sst"]] = 0.0
        G[self.out_dst] = self.gathered.numpy()
        q = G[3:lay["w2d"]:4]
        q[q < 0.000001] = 0.000001
        p = G[lay["precip"]:lay["sst"]]
        p[p < 0.00001] = 0.0
        sst = self.grids["base_sst"].ravel(order="F").copy()        # prescribed-SST mode of the bench
        G[lay["sst"]:lay["tisr"]] = np.maximum(sst, 272.0)

    def _grids(self):
        lay, G = self.lay, self.G
        return (G[:lay["w2d"]].reshape((4, 96, 48, 8), order="F"), G[lay["w2d"]:lay["precip"]].reshape((96, 48), order="F"),
                G[lay["precip"]:lay["sst"]].reshape((96, 48), order="F"), G[lay["sst"]:lay["tisr"]].reshape((96, 48), order="F"))

    def exchange_begin(self, t):
        self.pack(t)
        return tuple(np.asfortranarray(g.copy()) for g in self._grids())

    def load_forecast(self, f4d, f2d, tisr):
        self.F[:self.lay["w2d"]] = torch.from_numpy(np.asarray(f4d).ravel(order="F").copy())
        self.F[self.lay["w2d"]:] = torch.from_numpy(np.asarray(f2d).ravel(order="F").copy())
        self.tisr_dev[:] = torch.from_numpy(np.asarray(tisr).ravel(order="F").copy())

    def unpack(self, t):
        w4d, w2d, wp, wsst = self._grids()
        F = self.F.numpy()
        f4d = F[:self.lay["w2d"]].reshape((4, 96, 48, 8), order="F")
        f2d = F[self.lay["w2d"]:].reshape((96, 48), order="F")
        tisr = self.tisr_dev.numpy().reshape((96, 48), order="F")
        oc.step_scatter(self.regs, True, True, False, w4d, w2d, wp, wsst, f4d, f2d, tisr, self.sst_mean, self.sst_std,
                        nthreads=2)

    def exchange_end(self, t, f4d, f2d, tisr):
        self.load_forecast(f4d, f2d, tisr)
        self.unpack(t)


def _run_model(grids):
    def host_model(w4d, w2d, wsst):
        return oc.host_stub(w4d, w2d, grids["clim4d"], grids["clim2d"])
    return host_model


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grids = initial_grids()
        shard = OracleShard(rank, world, grids)
        stepper = H.HybridStepper(shard, rank=rank, world=world, dist=dist)
        outs = []
        for t in range(1, NSTEPS + 1):
            g = stepper.step(t, _run_model(grids), grids["tisr"])
            if rank == 0:
                outs.append([a.copy() for a in g])
        fb = {r: rc.feedback.copy() for r, rc in zip(shard.ids, shard.regs) if r in (0, 575, 576, 1151)}
        lm = {r: rc.local_model.copy() for r, rc in zip(shard.ids, shard.regs) if r in (0, 575, 576, 1151)}
        np.save(f"{out_path}.rank{rank}.npy", np.array([outs if rank == 0 else None, fb, lm], dtype=object),
                allow_pickle=True)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_step_matches_single_process(tmp_path):
    out = str(tmp_path / "mr")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    # single-process reference run of the same model (world = 1 path of the same stepper)
    grids = initial_grids()
    shard = OracleShard(0, 1, grids)
    stepper = H.HybridStepper(shard)
    ref = [[a.copy() for a in stepper.step(t, _run_model(grids), grids["tisr"])] for t in range(1, NSTEPS + 1)]
    r0 = np.load(out + ".rank0.npy", allow_pickle=True)
    r1 = np.load(out + ".rank1.npy", allow_pickle=True)
    for t in range(NSTEPS):
        for a, b in zip(r0[0][t], ref[t]):
            assert np.array_equal(a, b)
    full = {w["region"]: rc for w, rc in zip(shard.ws, shard.regs)}
    for res in (r0, r1):
        for r, v in res[1].items():
            assert np.array_equal(v, full[r].feedback)
        for r, v in res[2].items():
            assert np.array_equal(v, full[r].local_model)
    assert set(r0[1]) == {0, 575} and set(r1[1]) == {576, 1151}
    # and the scatter table agrees with the oracle's own slice-by-slice gather of the same outvecs
    gc = oc.step_gather(shard.regs, True, False, None, None)
    shard.gathered[:] = torch.from_numpy(np.concatenate([rc.outvec for rc in shard.regs]))
    shard.pack(0)
    ge = shard._grids()
    for a, b in zip(ge[:3], gc[:3]):
        assert np.array_equal(a, b)
