"""tools/strouhal.py (the headless scripts/lift.py) on a synthetic lift signal with a known period."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_known_period_is_recovered(tmp_path):
    import strouhal

    U, D, period = 0.1333, 50.0, 1693.5
    t = np.arange(0, 120000, 140.0)
    cl = 1.25 * np.sin(2 * np.pi * t / period + 0.3) * (1 - np.exp(-t / 8000.0))
    cd = 3.77 + 0.2 * np.sin(4 * np.pi * t / period)
    forces = {"timestep": t, "drag_force": cd, "lift_force": cl, "drag_coeff": cd, "lift_coeff": cl}
    res = strouhal.analyse(forces, U, D)
    want = D / (period * U)
    assert abs(res["strouhal_lift_py"] / want - 1) < 3e-3        # limited by the 140-step sampling grid
    assert abs(res["strouhal_refined"] / want - 1) < 3e-4
    assert abs(res["strouhal_fft"] / want - 1) < 2e-3
    assert abs(res["cl_amplitude"] - 1.25) < 0.01 and abs(res["cd_mean_from_start"] - 3.77) < 0.01
    assert res["peaks"] == int((120000 - 30000) / period) or res["peaks"] == int((120000 - 30000) / period) + 1
    # file round trip through the csv loaders
    p = tmp_path / "forces.csv"
    with open(p, "w") as f:
        f.write("timestep,drag_force,lift_force,drag_coeff,lift_coeff\n")
        for a, b, c in zip(t, cd, cl):
            f.write("%d,%.8f,%.8f,%.8f,%.8f\n" % (a, b, c, b, c))
    q = tmp_path / "simulation_params.csv"
    q.write_text("parameter,value\ninlet_velocity,0.13330000\ncylinder_radius,25\n")
    got = strouhal.analyse(strouhal.load_forces(str(p)), strouhal.load_params(str(q))["inlet_velocity"], 50.0)
    assert abs(got["strouhal_refined"] / want - 1) < 3e-4
