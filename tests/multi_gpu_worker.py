"""Worker of tests/test_gpu_multi.py: one x-slab per GPU under torchrun; writes its slab's state."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_dir, steps, seed, case_name = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    mode = sys.argv[5] if len(sys.argv) > 5 else "run"
    import torch.distributed as dist

    dist.init_process_group("gloo")  # only to hand the NCCL id around; the halo is the engine's own NCCL
    import lbm_b200
    import parity_util as util
    from test_gpu_multi import CASES

    CASE = CASES[case_name]

    params = util.case_to_params(CASE, flags=(1 if case_name == "periodic" else 0) | (16 if os.environ.get("LBM_TEST_AA") == "1" else 0))
    s, slab = lbm_b200.create_slab_solver(params, dist)
    s.initialise()
    if seed:
        state = util.random_state(CASE, seed)  # the global state; this slab uploads its columns
        local = np.ascontiguousarray(state[:, slab.x_start:slab.x_start + slab.lnx + 2, :])
        s.upload_f(local, iteration=0)
    if mode == "skewed":
        # observers in the middle of a run, ranks deliberately out of step
        import time

        out = {}
        for k, n in enumerate((1, 3, 4, 7, 2, 5)):
            s.step(n)
            if slab.rank % 2 == 1:
                time.sleep(0.3)
            rho, ux, uy = s.macros()
            out["rho_%d" % k], out["ux_%d" % k], out["uy_%d" % k] = rho.copy(), ux.copy(), uy.copy()
            if slab.rank % 2 == 0:
                time.sleep(0.3)
        np.savez(os.path.join(out_dir, "skew%d.npz" % slab.rank), f_next=s.f_next()[1:-1, 1:-1], **out)
        s.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    if mode == "silent":
        # rank 1 never steps: rank 0 must get an error, not a hang
        import time

        code, msg, t0 = 0, b"", time.time()
        if slab.rank == 0:
            try:
                s.step(6)
                s.sync()
            except lbm_b200.LbmError as e:
                code, msg = e.code, str(e).encode()
        secs = time.time() - t0
        dist.barrier()  # rank 1 waits here, alive (its memory stays mapped), without ever launching a step
        np.savez(os.path.join(out_dir, "silent%d.npz" % slab.rank), code=code, message=np.frombuffer(msg, dtype=np.uint8),
                 seconds=secs, halo_p2p=s.info().halo_p2p)
        os._exit(0)  # no collective teardown with a failed exchange
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
