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
