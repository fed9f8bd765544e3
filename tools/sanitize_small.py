"""Small-grid tour of every kernel for compute-sanitizer (one tool per gpurun call):
  compute-sanitizer --tool memcheck python tools/sanitize_small.py
A-B and AA, channel / periodic / forced modes, odd ny, upload, every observer, lbm_run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lbm_b200 as L

rng = np.random.default_rng(0)
w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
n = 0
for aa in (0, 16):
    for nx, ny, flags, kw in ((40, 24, 0, {}), (33, 17, 0, {}), (32, 16, 1 | 2 | 4 | 8, {}), (32, 16, 1 | 4, {"body_force_x": 1e-6}),
                              (24, 12, 2, {}), (6, 4, 0, {}), (1, 1, 0, {})):
        for variant in (0, 1):
            p = L.SimulationParams(nx=nx, ny=ny, output_frequency=3, cylinder_radius=0.2, cylinder_x=0.3, flags=flags | aa, **kw)
            s = L.Solver(p)
            s.set_kernel_variant(variant)
            s.initialise()
            s.step(5)
            s.f_next(); s.f_current(); s.macros(); s.forces(); s.max_velocity(); s.check_stability()
            st = np.zeros((ny + 2, nx + 2, 9))
            st[:] = w * (1 + 0.03 * rng.standard_normal((ny + 2, nx + 2, 9)))
            s.upload_f(st, 0)
            rows, bad = s.run(8)
            s.f_next(); s.f_current(); s.macros()
            s.close()
            n += 1
print("sanitize tour ok:", n, "solver configurations")
