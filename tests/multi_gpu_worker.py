"""Worker of tests/test_gpu_multi.py: one x-slab per GPU under torchrun; writes its slab's state."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_dir, steps, seed, case_name = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    import torch.distributed as dist

    dist.init_process_group("gloo")  # only to hand the NCCL id around; the halo is the engine's own NCCL
    import lbm_b200
    import parity_util as util
    from test_gpu_multi import CASES

    CASE = CASES[case_name]

    params = util.case_to_params(CASE, flags=1 if case_name == "periodic" else 0)
    s, slab = lbm_b200.create_slab_solver(params, dist)
    s.initialise()
    if seed:
        state = util.random_state(CASE, seed)  # the global state; this slab uploads its columns
        local = np.ascontiguousarray(state[:, slab.x_start:slab.x_start + slab.lnx + 2, :])
        s.upload_f(local, iteration=0)
    rows, bad = s.run(steps)
    fx, fy = s.forces()
    tot = s.allreduce([fx, fy], 0)
    rho, ux, uy = s.macros()
    g = s.gather_macros()
    np.savez(os.path.join(out_dir, "slab%d.npz" % slab.rank), f_next=s.f_next()[1:-1, 1:-1], f_current=s.f_current()[1:-1, 1:-1],
             rho=rho, ux=ux, uy=uy, rows=rows, bad=bad, forces_total=tot, halo_p2p=s.info().halo_p2p, maxvel=s.allreduce([s.max_velocity()], 2),
             **({"g_rho": g[0], "g_ux": g[1], "g_uy": g[2]} if g is not None else {}))
    s.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
