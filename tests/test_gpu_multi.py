"""Multi-GPU parity: N x-slabs with the NCCL halo exchange (overlapped and not) against the
1-rank CPU oracle.  Needs >= 2 GPUs on the box (skipped otherwise); run with -m gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
import parity_util as util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the cylinder straddles the face between slabs 0 and 1 of a 2-slab run (x = 64)
CASES = {
    "even": O.Case(nx=128, ny=48, cylinder_x=0.5, cylinder_radius=0.2, output_frequency=6, inlet_velocity=0.04),
    # odd ny: scalar bulk kernel, halo fused into the fix-up kernel instead of the bulk launch
    "periodic": O.Case(nx=96, ny=40, cylinder_x=0.02, cylinder_radius=0.2, output_frequency=6, inlet_velocity=0.04),  # flags=1 in the worker
    "odd": O.Case(nx=128, ny=47, cylinder_x=0.5, cylinder_radius=0.2, output_frequency=6, inlet_velocity=0.04),
}


def n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("overlap", ["1", "0", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("seed,case_name", [(0, "even"), (5, "even"), (5, "odd")])
def test_slabs_match_single_rank_oracle(tmp_path, world, overlap, seed, case_name):
    CASE = CASES[case_name]
    if n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    steps = 37
    # "1": edge kernel + halo fused over peer memory when CUDA IPC is available (else the NCCL path);
    # "nccl": the overlapped NCCL send/recv path, forced; "0": exchange in stream order, no overlap
    env = dict(os.environ, LBM_B200_OVERLAP="0" if overlap == "0" else "1", LBM_B200_P2P="0" if overlap == "nccl" else "1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29800 + os.getpid() % 150 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(tmp_path),
           str(steps), str(seed), case_name]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    if seed:
        state = util.random_state(CASE, seed)
        o = util.oracle_with_state(CASE, state)
    else:
        o = O.Oracle(CASE)
    rows, bad = o.run(steps)
    parts = [np.load(tmp_path / ("slab%d.npz" % k)) for k in range(world)]
    for key, want in (("f_next", o.f_next[1:-1, 1:-1]), ("f_current", o.f_current[1:-1, 1:-1]), ("rho", o.rho), ("ux", o.ux), ("uy", o.uy)):
        got = np.concatenate([p[key] for p in parts], axis=1)
        assert np.array_equal(got, want), "%s differs: max %.3e" % (key, np.abs(got - want).max())
    assert np.array_equal(parts[0]["g_rho"], o.rho) and np.array_equal(parts[0]["g_ux"], o.ux) and np.array_equal(parts[0]["g_uy"], o.uy)
    assert all(int(p["bad"]) == bad == -1 for p in parts)
    if overlap == "nccl":
        assert all(int(p["halo_p2p"]) == 0 for p in parts)
    print("halo_p2p:", [int(p["halo_p2p"]) for p in parts])
    total = sum(p["rows"][:, 1:3] for p in parts)
    assert np.array_equal(parts[0]["rows"][:, 0], rows[:, 0]) and np.abs(total - rows[:, 1:3]).max() <= 1e-13
    ofx, ofy = o.forces()
    assert abs(parts[1]["forces_total"][0] - ofx) <= 1e-13 and abs(parts[1]["forces_total"][1] - ofy) <= 1e-13
    assert abs(float(parts[0]["maxvel"][0]) - o.max_velocity()) <= 1e-15


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_periodic_slabs_match_single_gpu(tmp_path, p2p):
    """Periodic-x channel over 2 slabs: every rank has a neighbour on BOTH sides (the same peer), the
    situation of the middle ranks of a longer chain.  Reference: the same engine on one GPU."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import lbm_b200

    steps, seed = 29, 3
    env = dict(os.environ, LBM_B200_P2P=p2p)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29950 + os.getpid() % 40), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(tmp_path),
           str(steps), str(seed), "periodic"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    case = CASES["periodic"]
    s = lbm_b200.Solver(util.case_to_params(case, flags=1))
    s.initialise()
    s.upload_f(util.random_state(case, seed), 0)
    rows, bad = s.run(steps)
    parts = [np.load(tmp_path / ("slab%d.npz" % k)) for k in range(2)]
    got, want = np.concatenate([p["f_next"] for p in parts], axis=1), s.f_next()[1:-1, 1:-1]
    bad_cells = np.argwhere((got != want).any(axis=2))
    assert len(bad_cells) == 0, "%d cells differ, x range %d..%d, first %s" % (
        len(bad_cells), bad_cells[:, 1].min(), bad_cells[:, 1].max(), bad_cells[:5].tolist())
    assert np.array_equal(np.concatenate([p["f_current"] for p in parts], axis=1), s.f_current()[1:-1, 1:-1])
    rho, ux, uy = s.macros()
    assert np.array_equal(parts[0]["g_rho"], rho) and np.array_equal(parts[0]["g_ux"], ux) and np.array_equal(parts[0]["g_uy"], uy)
    print("halo_p2p:", [int(p["halo_p2p"]) for p in parts])
    s.close()
