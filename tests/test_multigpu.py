"""Multi-GPU parity (skipped where the box has fewer GPUs; the builder runs it with gpurun --gpus 2 / 8):

* tests/mr_gpu_check.py under torchrun at 2, 4 and 8 ranks: the exchange inside the engine (sml_comm_bootstrap, peer
  stores over NVLink, no host collective) against NCCL host collectives bit for bit, against the CPU oracle, identical
  grids on every rank, the run_speedy flag, and the coupled atmosphere + ocean model;
* the compiled C++ replay driver at world = 2 (two processes connected through sml_comm_bootstrap over a POSIX
  shared-memory all-gather -- no Python, no torch, no NCCL anywhere): grids bit-identical to its single-rank run.
"""
import importlib
import os
import socket
import subprocess
import sys
import uuid

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_engine_exchange_matches_nccl_and_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(here, "mr_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "MULTIGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("overlap", [False, True])
def test_cpp_replay_driver_two_ranks(tmp_path, overlap):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from case_io import N2, N4, read_output, write_case
    from helpers import initial_grids, region_weights, syn

    B = importlib.import_module("speedy-ml_b200.build")
    exe = B.build_host_driver()
    NSTEPS, SYNC, R = 5, 3, 1152
    ws = [region_weights(R, r, m=300, with_dense_win=False) for r in range(R)]
    G = initial_grids()
    rng = np.random.default_rng(77)
    x0 = [0.1 * rng.standard_normal(w["n"]) for w in ws]
    fb0 = [rng.standard_normal(w["D"]) for w in ws]
    lm0 = [rng.standard_normal(w["S"]) for w in ws]
    sync = [syn.ar1_series(w["D"], SYNC, rng) for w in ws]
    case = str(tmp_path / "case.bin")
    write_case(case, ws, x0, fb0, lm0, sync, G, NSTEPS)
    extra = ["--batched-sync"] + (["--overlap"] if overlap else [])

    one = str(tmp_path / "one.bin")
    p = subprocess.run([exe, case, one] + extra, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "program finished correctly" in p.stdout, p.stdout + p.stderr
    ref_steps, ref_ov, ref_fb = read_output(one, ws, NSTEPS)

    shm = "smlreplay_" + uuid.uuid4().hex[:12]
    outs = [str(tmp_path / f"two_{r}.bin") for r in range(2)]
    procs = [subprocess.Popen([exe, case, outs[r], "--rank", str(r), "--world", "2", "--shm", shm] + extra,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    for r, pr in enumerate(procs):
        so, se = pr.communicate(timeout=600)
        assert pr.returncode == 0 and "program finished correctly" in so, f"rank {r}: {so}{se}"

    per = N4 + 3 * N2
    half = R // 2
    raw0 = np.fromfile(outs[0], dtype=np.float64)
    raw1 = np.fromfile(outs[1], dtype=np.float64)
    # rank 0: the grids of every step, bit-identical to the single-rank run (sharding does not change any region's
    # arithmetic and the exchange is pure data movement)
    for t in range(NSTEPS):
        got = raw0[t * per:(t + 1) * per]
        want = np.concatenate([a.ravel(order="F") for a in ref_steps[t]])
        assert np.array_equal(got, want), f"step {t + 1}: two-rank grids differ from the single-rank run"
    # both ranks appended the grids THEY assembled in the last step
    last0 = raw0[NSTEPS * per:(NSTEPS + 1) * per]
    last1 = raw1[:per]
    want = np.concatenate([a.ravel(order="F") for a in ref_steps[-1]])
    assert np.array_equal(last0, want) and np.array_equal(last1, want)
    # final outvec / feedback of each rank's regions
    for rank, raw, pos in ((0, raw0, (NSTEPS + 1) * per), (1, raw1, per)):
        for i in range(rank * half, (rank + 1) * half):
            w = ws[i]
            assert np.array_equal(raw[pos:pos + w["P"]], ref_ov[i]); pos += w["P"]
            assert np.array_equal(raw[pos:pos + w["D"]], ref_fb[i]); pos += w["D"]
        assert pos == raw.size
