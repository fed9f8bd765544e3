"""world_size-2 (and 4) checks of the x-slab partition on CPU, gloo backend.

Each rank drives one slab of the CPU oracle and exchanges, after every collision, exactly what
the engine's NCCL halo exchange sends (csrc/lbm_engine.cu `exchange`): three populations per
face, rows 0..ny-1, no corners.  The joined interiors must equal the 1-rank oracle bit for bit,
which is the property the multi-GPU path relies on (and which the reference's own multi-rank
2-D decompositions do not have)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, case_kw, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    import lbm_b200

    case = O.Case(**case_kw)
    slab = lbm_b200.Slab(rank, world, case.nx, case.ny)
    assert lbm_b200.slabs.env_rank_world() == (rank, world, rank)
    o = O.Oracle(case, slab.x_start, slab.lnx)
    ny = case.ny
    fx_rows = []
    for t in range(steps):
        o.collide()
        if t % case.output_frequency == 0:
            f = torch.tensor(o.forces(), dtype=torch.float64)
            dist.all_reduce(f)  # MPI_Reduce(SUM) of include/LBMIO.h:167-168
            fx_rows.append((t, float(f[0]), float(f[1])))
        o.edge_ghosts()
        reqs, recv = [], {}
        for east, peer, pops in ((True, slab.east, lbm_b200.slabs.EAST_GOING), (False, slab.west, lbm_b200.slabs.WEST_GOING)):
            if peer < 0:
                continue
            full = o.get_halo(east).reshape(ny, 9)
            send = torch.from_numpy(np.ascontiguousarray(full[:, list(pops)]))  # 3 populations only
            recv[east] = torch.empty(ny, 3, dtype=torch.float64)
            reqs.append(dist.isend(send, peer, tag=int(east)))
            reqs.append(dist.irecv(recv[east], peer, tag=int(not east)))
        for r in reqs:
            r.wait()
        for east, buf in recv.items():
            # what arrives from the east neighbour moves west, and vice versa; the six other
            # populations of the ghost column are never pulled: leave them at 0
            pops = lbm_b200.slabs.WEST_GOING if east else lbm_b200.slabs.EAST_GOING
            full = np.zeros((ny, 9))
            full[:, list(pops)] = buf.numpy()
            o.put_halo(east, full.reshape(-1))
        o.stream()
        o.boundaries()
        ok = torch.tensor([1.0 if o.check_stability() else 0.0], dtype=torch.float64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        assert ok.item() == 1.0
    np.savez(os.path.join(out_dir, "slab%d.npz" % rank), f_current=o.f_current[1:-1, 1:-1], f_next=o.f_next[1:-1, 1:-1],
             rho=o.rho, ux=o.ux, uy=o.uy, forces=np.array(fx_rows), halo_bytes=slab.halo_bytes_per_step())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case_kw,steps", [
    (2, dict(nx=64, ny=32, cylinder_x=0.5, cylinder_radius=0.2, output_frequency=5), 40),   # cylinder straddles the slab face
    (4, dict(nx=96, ny=24, cylinder_x=0.3, cylinder_radius=0.15, output_frequency=7), 30),
])
def test_slabs_equal_the_single_rank_run(tmp_path, world, case_kw, steps):
    from oracle import oracle as O

    port = 29600 + (os.getpid() % 300) + world
    mp.spawn(_worker, args=(world, port, case_kw, steps, str(tmp_path)), nprocs=world, join=True)
    case = O.Case(**case_kw)
    one = O.Oracle(case)
    rows, bad = one.run(steps)
    assert bad == -1
    parts = [np.load(tmp_path / ("slab%d.npz" % r)) for r in range(world)]
    for key, want in (("f_current", one.f_current[1:-1, 1:-1]), ("f_next", one.f_next[1:-1, 1:-1]), ("rho", one.rho),
                      ("ux", one.ux), ("uy", one.uy)):
        got = np.concatenate([p[key] for p in parts], axis=1)
        assert np.array_equal(got, want), key
    # forces: per-slab partial sums added in rank order differ from the serial sum by rounding only
    f = parts[0]["forces"]
    assert np.array_equal(f[:, 0], rows[:, 0])
    assert np.abs(f[:, 1:3] - rows[:, 1:3]).max() <= 1e-14
    assert int(parts[0]["halo_bytes"]) == 3 * case.ny * 8 and (world < 3 or int(parts[1]["halo_bytes"]) == 2 * 3 * case.ny * 8)


def test_slab_rules():
    import lbm_b200

    s = lbm_b200.Slab(1, 8, 32768, 8192)
    assert (s.lnx, s.x_start, s.west, s.east, s.has_inlet, s.has_outlet) == (4096, 4096, 0, 2, False, False)
    assert s.halo_bytes_per_step() == 2 * 3 * 8192 * 8
    assert s.owner_of_column(int(0.2 * 32768)) == 1  # the cylinder centre of BASELINE config 5 lives on slab 1
    assert lbm_b200.Slab(0, 8, 32768, 8192).has_inlet and lbm_b200.Slab(7, 8, 32768, 8192).has_outlet
    p = lbm_b200.Slab(0, 4, 64, 8, periodic_x=True)
    assert (p.west, p.east, p.has_inlet) == (3, 1, False)
    with pytest.raises(ValueError):
        lbm_b200.Slab(0, 3, 64, 8)


# ---- the temporally blocked passes over slabs, one slab per PROCESS, wide halo over gloo --------------------------
def _tb_worker(rank, world, port, case_kw, depths, halo_w, out_dir):
    """Each rank runs the sm_100a kernel's own thread program for its slab on the host (oracle/prototypes/tb_emul.cpp)
    and ships what the last stage stored for its neighbours -- the wide halo the GPUs push over NVLink -- over gloo."""
    import ctypes as C

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O

    case = O.Case(**case_kw)
    o = O.Oracle(case)
    rng = np.random.default_rng(21)  # (same seed on every rank: the global start state)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    init = o.f_current.copy()
    f = w * (1.0 + 0.05 * rng.standard_normal((case.ny + 2, case.nx + 2, 9)))
    for sl in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
        f[sl] = init[sl]
    o.f_current[...] = f
    o.run(1)
    state = np.ascontiguousarray(o.f_next.copy())
    solid = np.zeros((case.ny + 2, case.nx + 2), dtype=np.uint8)
    solid[1:-1, 1:-1] = o.solid
    L = C.CDLL(os.path.join(ROOT, "oracle", "prototypes", "libtbemul.so"))
    L.tbs_create.restype = C.c_void_p
    L.tbs_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    L.tbs_pass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    for name in ("tbs_get_outbox", "tbs_put_inbox"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.tbs_get_interior.argtypes = [C.c_void_p, C.c_void_p]
    L.tbs_destroy.argtypes = [C.c_void_p]
    h = L.tbs_create(state.ctypes.data, solid.ctypes.data, case.nx, case.ny, case.tau, case.inlet_velocity, rank, world, halo_w, 0)
    lnx, ny, it = case.nx // world, case.ny, 1
    west, east = (rank - 1 if rank > 0 else -1), (rank + 1 if rank < world - 1 else -1)
    sent = 0
    for depth in depths:
        bad = L.tbs_pass(h, depth, 16, 5, max(halo_w, 3), it)
        assert bad == 0x7fffffff
        it += depth
        reqs, inbox = [], {}
        for to_east, peer in ((1, east), (0, west)):
            if peer < 0:
                continue
            box = np.empty((halo_w, 9, ny))
            L.tbs_get_outbox(h, to_east, box.ctypes.data)
            sent += int(np.isfinite(box).sum()) * 8
            inbox[to_east] = torch.empty(halo_w, 9, ny, dtype=torch.float64)
            reqs.append(dist.isend(torch.from_numpy(box), peer, tag=to_east))
            reqs.append(dist.irecv(inbox[to_east], peer, tag=1 - to_east))
        for r in reqs:
            r.wait()
        for side, buf in inbox.items():  # what arrived from the neighbour on that side
            L.tbs_put_inbox(h, side, np.ascontiguousarray(buf.numpy()).ctypes.data)
    out = np.empty((ny, lnx, 9))
    L.tbs_get_interior(h, out.ctypes.data)
    L.tbs_destroy(h)
    np.savez(os.path.join(out_dir, "tb%d.npz" % rank), f_next=out, sent=sent, passes=len(depths))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,depths,halo_w", [(2, (2, 2, 1, 2), 2), (4, (2, 1, 2), 2), (2, (3, 2, 3), 3)])
def test_temporally_blocked_slabs_with_the_wide_halo_over_gloo(tmp_path, world, depths, halo_w):
    from oracle import oracle as O

    case_kw = dict(nx=32 * world, ny=24, cylinder_x=1.0 / world if world > 1 else 0.5, cylinder_radius=0.2, output_frequency=5,
                   inlet_velocity=0.04)
    port = 29300 + (os.getpid() % 200) + 7 * world + halo_w
    mp.spawn(_tb_worker, args=(world, port, case_kw, depths, halo_w, str(tmp_path)), nprocs=world, join=True)
    case = O.Case(**case_kw)
    o = O.Oracle(case)
    rng = np.random.default_rng(21)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    init = o.f_current.copy()
    f = w * (1.0 + 0.05 * rng.standard_normal((case.ny + 2, case.nx + 2, 9)))
    for sl in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
        f[sl] = init[sl]
    o.f_current[...] = f
    o.run(1 + sum(depths))
    parts = [np.load(tmp_path / ("tb%d.npz" % r)) for r in range(world)]
    got = np.concatenate([p["f_next"] for p in parts], axis=1)
    assert np.array_equal(got, o.f_next[1:-1, 1:-1])
    # the wire traffic of a pass: per face 3 population columns for the farthest halo column, 6 for the next,
    # 9 beyond (lbm_tb.cuh tb_push) -- 9 columns per face at depth 2, 18 at depth 3
    per_face = sum(3 if d == halo_w - 1 else (6 if d == halo_w - 2 else 9) for d in range(halo_w)) * case.ny * 8
    assert int(parts[0]["sent"]) == per_face * len(depths)           # an end slab: one face
    if world > 2:
        assert int(parts[1]["sent"]) == 2 * per_face * len(depths)   # a middle slab: two faces
