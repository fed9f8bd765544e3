#!/usr/bin/env python
"""BASELINE config 3 (cylinder flow at Re = 200 on 8192 x 2048, one B200) with DENSE force sampling: the momentum-
exchange forces of every second iteration (output_frequency = 2: each temporally blocked pass ends on an output step)
through the fixed parallel reduction tree (lbm_set_force_mode(LBM_FORCES_TREE), ~3 us per sample instead of ~25 us for
the reference-ordered sum), and the Strouhal number from those samples -- no longer limited to the 140-step grid of
the reference's forces.csv.

  python tools/c3_dense.py [--steps 280000] [--start 140000] [--out gpurun_out/c3_dense]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=8192)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=280000)
    ap.add_argument("--start", type=int, default=140000)
    ap.add_argument("--of", type=int, default=2)
    ap.add_argument("--ordered", action="store_true", help="the reference-ordered sum instead of the tree")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c3_dense"))
    a = ap.parse_args()
    import lbm_b200
    import strouhal

    u, tau = 0.1333, 0.9095  # Re = u D / nu = 0.1333 * 204.8 / 0.13650 = 200.0 (profiles/r01_c3_re200.md)
    p = lbm_b200.SimulationParams(nx=a.nx, ny=a.ny, tau=tau, inlet_velocity=u, output_frequency=a.of)
    s = lbm_b200.Solver(p)
    s.set_force_mode(0 if a.ordered else 1)
    s.initialise()
    info = s.info()
    rows_all = []
    t0 = time.time()
    done = 0
    while done < a.steps:
        n = min(20000, a.steps - done)
        rows, bad = s.run(n)
        assert bad == -1, "unstable at %d" % bad
        rows_all.append(rows)
        done += n
    s.sync()
    wall = time.time() - t0
    rows = np.concatenate(rows_all)
    s.close()
    D = 2.0 * p.get_cylinder_radius_cells()
    forces = {"timestep": rows[:, 0], "drag_force": rows[:, 1], "lift_force": rows[:, 2], "drag_coeff": rows[:, 3], "lift_coeff": rows[:, 4]}
    res = strouhal.analyse(forces, u, D, start=a.start)
    # the same analysis on the reference's 140-step grid, from the same run
    coarse = {k: v[(rows[:, 0] % 140) == 0] for k, v in forces.items()}
    res140 = strouhal.analyse(coarse, u, D, start=a.start)
    out = {"case": "cylinder flow Re = 200, %dx%d, tau %.4f, u_in %.4f, %d iterations" % (a.nx, a.ny, tau, u, a.steps),
           "force_mode": "ordered" if a.ordered else "tree", "output_frequency": a.of, "samples": int(len(rows)),
           "pass_depth": int(info.pass_depth), "kernel_variant": int(info.kernel_variant),
           "wall_s": wall, "mlups_wall_clock_forces_included": a.nx * a.ny * a.steps / wall / 1e6,
           "dense": res, "on_the_140_step_grid": res140}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out + ".json", "w"), indent=1)
    # lift history from `start` on, every 10th sample, for the record
    sel = rows[:, 0] >= a.start
    np.savetxt(a.out + "_lift.csv.gz", rows[sel][::10][:, [0, 3, 4]], delimiter=",", header="timestep,drag_coeff,lift_coeff", fmt=["%d", "%.8f", "%.8f"], comments="")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
