"""Multi-GPU parity check, launched by torchrun (one rank per GPU; 2, 4 or 8 ranks):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mr_gpu_check.py
Runs the same closed hybrid loop (small reservoirs on the full 1152-region tiling) with
  (a) host collectives (NCCL all-gather of the outvec slabs, NCCL broadcast of the forecast),
  (b) the exchange INSIDE the engine after sml_comm_bootstrap: outvecs pushed by the readout-finish kernel, the root's
      forecast block pushed by k_peer_push, consumers waiting on device flags -- no collective issued by the host,
  (c) (b) in the overlapped mode,
and requires: (a) == (b) bit for bit on rank 0, (c) within 1e-11 of (a), (a) within 1e-10 of the single-process CPU
oracle, the grids assembled on EVERY rank bit-identical to rank 0's, and run_speedy = .false. reaching every rank.
Then the coupled model (ocean reservoirs; their slabs pushed by the engine as well) against the oracle.
Prints MULTIGPU_OK on success (tests/test_multigpu.py looks for it)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from helpers import c_region, initial_grids, oc, region_weights, rel_inf  # noqa: E402

E = importlib.import_module("speedy-ml_b200.engine")
H = importlib.import_module("speedy-ml_b200.hybrid")
R, M, NSTEPS = 1152, 300, 6


def reset_engine(eng, ws):
    rng = np.random.default_rng(5)
    draws = [(rng.standard_normal(576), rng.standard_normal(132), 0.1 * rng.standard_normal(700)) for _ in range(R)]
    for r, w in ws.items():
        fb, lm, x0 = draws[r]
        eng.feedback_set(r, fb[:w["D"]])
        eng.local_model_set(r, lm[:w["S"]])
        eng.state_set(r, x0[:w["n"]])
    return draws


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    G = initial_grids()
    eng = E.Engine(number_of_regions=R, irank=rank, numprocs=world, device=local, sst_prescribed=True, stream=stream)
    ws = {r: region_weights(R, r, m=M, with_dense_win=False) for r in eng.region_indices}
    for r, w in ws.items():
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    eng.finalize()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    eng.set_sst_prescribed(G["base_sst"])
    shard = H.EngineShard(eng, torch)

    def reset():
        return reset_engine(eng, ws)

    def host_model(w4d, w2d, wsst):
        return oc.host_stub(w4d, w2d, G["clim4d"], G["clim2d"])

    def run(overlap):
        reset()
        st = H.HybridStepper(shard, rank=rank, world=world, dist=dist, overlap=overlap)
        out = []
        for t in range(1, NSTEPS + 1):
            g = st.step(t, host_model, G["tisr"])
            if rank == 0:
                out.append([a.copy() for a in g])
        torch.cuda.synchronize()
        fbs = {r: eng.feedback_get(r) for r in list(ws)[:3]}
        mine = eng.grids_get()          # what THIS rank assembled in the last step (every rank rebuilds the whole grid)
        if overlap:
            eng.set_overlap(False)
        dist.barrier()
        return out, fbs, mine

    def same_on_every_rank(grids):
        """the last step's grids of every rank against rank 0's, bit for bit"""
        flat = torch.from_numpy(np.concatenate([a.ravel(order="F") for a in grids])).cuda()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        return bool(torch.equal(flat.view(torch.int64), ref.view(torch.int64)))

    nccl, fb_nccl, g_nccl = run(False)
    assert not eng.peer_attached() and not shard.comm_ready
    shard.bootstrap(dist)               # sml_comm_bootstrap: from here on the host issues no collective in a step
    assert eng.peer_attached() and shard.comm_ready
    peer, fb_peer, g_peer = run(False)
    ovl, fb_ovl, g_ovl = run(True)
    eng.peer_check()
    ok = True
    ok &= same_on_every_rank(g_nccl) and same_on_every_rank(g_peer) and same_on_every_rank(g_ovl)
    for r in fb_nccl:
        ok &= np.array_equal(fb_nccl[r], fb_peer[r]) and rel_inf(fb_ovl[r], fb_nccl[r]) < 1e-11
    # run_speedy: the root refuses the next SPEEDY step; the flag travels with the forecast to every rank
    st = H.HybridStepper(shard, rank=rank, world=world, dist=dist)
    ok &= eng.run_speedy() is True
    if rank == 0:
        eng.set_run_speedy(False)
    st.step(NSTEPS + 1, host_model, G["tisr"])
    ok &= eng.run_speedy() is False
    if rank == 0:
        eng.set_run_speedy(True)
    st.step(NSTEPS + 2, host_model, G["tisr"])
    ok &= eng.run_speedy() is True
    dist.barrier()
    if rank == 0:
        for t in range(NSTEPS):
            for a, b, c in zip(nccl[t], peer[t], ovl[t]):
                ok &= np.array_equal(a, b)
                ok &= rel_inf(c, a) < 1e-11
        # single-process oracle of the whole model
        all_ws = [region_weights(R, r, m=M, with_dense_win=False) for r in range(R)]
        rcs = [c_region(w) for w in all_ws]
        draws = reset()
        for w, rc in zip(all_ws, rcs):
            fb, lm, x0 = draws[w["region"]]
            rc.feedback[:], rc.local_model[:], rc.x[:] = fb[:w["D"]], lm[:w["S"]], x0[:w["n"]]
        sst_mean = np.array([w["mean"][-1] for w in all_ws])
        sst_std = np.array([w["std"][-1] for w in all_ws])
        has = np.ones(R, dtype=np.int32)
        oo = np.zeros((R, 4))
        for i in range(R):
            xs, xe, ys, ye, *_ = oc.getxyresextent(R, i)
            oo[i] = G["base_sst"][xs - 1:xe, ys - 1:ye].ravel(order="F")
        worst = 0.0
        for t in range(NSTEPS):
            oc.predict_all(rcs, nthreads=8)
            gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
            for a, b in zip(nccl[t], gc):
                worst = max(worst, rel_inf(a, b))
            f4, f2 = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
            oc.step_scatter(rcs, True, True, False, *gc, f4, f2, G["tisr"], sst_mean, sst_std, nthreads=8)
        ok &= worst < 1e-10
        print(f"rank0 of {world}: host collectives == engine exchange bitwise, overlap within 1e-11, every rank's grid "
              f"identical, run_speedy broadcast, oracle worst rel err {worst:.2e}, ok={ok}")
    dist.barrier()
    eng.close()

    # ---- coupled model: ocean reservoirs on the 'ocean' regions; their slabs are pushed to every rank by the engine
    # after each ocean step (and once for the seeded outvecs), the atmosphere slabs by the readout kernel, the forecast
    # by the root's push kernel.  Rank 0 compares the grids with a single-process oracle.
    from helpers import c_ocean, ocean_weights, sst_input_mask
    eng = E.Engine(number_of_regions=R, irank=rank, numprocs=world, device=local, sst_prescribed=False, stream=stream)
    for r, w in ws.items():
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], sst_bool_input=w["sst_bool_input"])
    wos = {r: ocean_weights(R, r, m=M, mean=ws[r]["mean"], std=ws[r]["std"], with_dense_win=False)
           for r in ws if sst_input_mask(r)}
    for r, w in wos.items():
        eng.region_upload(r, w["rows"], w["cols"], w["vals"], w["wout"], w["mean"], w["std"], win_compact=w["winc"],
                          win_col=w["wcol"], D=w["D"], kind=E.OCEAN)
    eng.finalize()
    eng.set_sst_static(G["base_sst"], G["sea_mask"])
    shard2 = H.EngineShard(eng, torch, ocean=True)
    shard2.bootstrap(dist)
    rng = np.random.default_rng(9)
    odraw = {r: (rng.standard_normal(128), 285.0 + 5.0 * rng.random(8)) for r in range(R)}   # same stream on every rank
    draws = reset_engine(eng, ws)
    for r, w in wos.items():
        eng.feedback_set(r, odraw[r][0][:w["D"]], kind=E.OCEAN)
        eng.outvec_set(r, odraw[r][1], kind=E.OCEAN)
    NS2 = 30
    st = H.HybridStepper(shard2, rank=rank, world=world, dist=dist)
    cgrids = []
    for t in range(1, NS2 + 1):
        g = st.step(t, host_model, G["tisr"])
        if rank == 0 and t in (1, 2, 27, 28, 29, 30):
            cgrids.append((t, [a.copy() for a in g]))
    torch.cuda.synchronize()
    eng.peer_check()
    if rank == 0:
        all_ws = [region_weights(R, r, m=M, with_dense_win=False) for r in range(R)]
        all_wos = {r: ocean_weights(R, r, m=M, mean=all_ws[r]["mean"], std=all_ws[r]["std"], with_dense_win=False)
                   for r in range(R) if sst_input_mask(r)}
        rcs = [c_region(w) for w in all_ws]
        cos = {r: c_ocean(w) for r, w in all_wos.items()}
        for w, rc in zip(all_ws, rcs):
            fb, lm, x0 = draws[w["region"]]
            rc.feedback[:], rc.local_model[:], rc.x[:] = fb[:w["D"]], lm[:w["S"]], x0[:w["n"]]
        for r, co in cos.items():
            co.feedback[:] = odraw[r][0][:all_wos[r]["D"]]
            co.outvec[:] = odraw[r][1]
            co.x[:] = 0.0
        sst_mean = np.array([w["mean"][-1] for w in all_ws])
        sst_std = np.array([w["std"][-1] for w in all_ws])
        has = np.array([1 if r in cos else 0 for r in range(R)], dtype=np.int32)
        want = dict(cgrids)
        worst2 = 0.0
        for t in range(1, NS2 + 1):
            oc.predict_all(rcs, nthreads=8)
            if (t * 6) % 168 == 0:
                for co in cos.values():
                    co.predict()
            oo = np.zeros((R, 4))
            for r, co in cos.items():
                oo[r] = co.outvec[:4]
            gc = oc.step_gather(rcs, True, True, G["base_sst"], G["sea_mask"], ocean_out=oo, has_ocean=has)
            if t in want:
                for a, b in zip(want[t], gc):
                    worst2 = max(worst2, rel_inf(a, b))
            f4, f2 = oc.host_stub(gc[0], gc[1], G["clim4d"], G["clim2d"])
            oc.step_scatter(rcs, True, True, False, *gc, f4, f2, G["tisr"], sst_mean, sst_std, nthreads=8)
            for r, co in cos.items():
                co.build_feedback(rcs[r], t, gc[3])
        ok &= worst2 < 1e-10
        print(f"rank0: coupled {world}-rank loop (every exchange inside the engine) vs oracle: {worst2:.2e}, ok={ok}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()      # nobody frees its exchange block while a peer may still push into it
    eng.close()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("MULTIGPU_FAILED")
    if rank == 0:
        print("MULTIGPU_OK")


if __name__ == "__main__":
    main()
