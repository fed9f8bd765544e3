#!/usr/bin/env python
"""Headless force-coefficient analysis: the Strouhal number exactly as the reference's
scripts/lift.py:71-98 defines it (peaks of lift_coeff for timestep >= 30000 found with
scipy.signal.find_peaks(prominence=0.5); St = D / (mean peak spacing * U), D = 2*cylinder_radius
cells from simulation_params.csv), without its matplotlib/seaborn plotting, plus two estimates
that are not limited by the output_frequency sampling grid (parabolic peak refinement and the FFT
peak), and the mean / range of C_D and C_L as IOManager::calculate_time_averaged_drag prints them
(reference include/LBMIO.h:367-413, rows with timestep > 1000).

  python tools/strouhal.py [forces.csv] [simulation_params.csv] [--start 30000] [--json]
"""
from __future__ import annotations

import argparse
import json
import sys

import numpy as np


def load_forces(path):
    a = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    return {"timestep": a[:, 0], "drag_force": a[:, 1], "lift_force": a[:, 2], "drag_coeff": a[:, 3], "lift_coeff": a[:, 4]}


def load_params(path):
    out = {}
    for line in open(path).read().splitlines()[1:]:
        k, v = line.split(",")
        out[k] = float(v)
    return out


def analyse(forces, U, D, start=30000, prominence=0.5):
    from scipy.signal import find_peaks

    t, cl, cd = forces["timestep"], forces["lift_coeff"], forces["drag_coeff"]
    sel = t >= start
    ts, ls, ds = t[sel], cl[sel], cd[sel]
    res = {"U": U, "D": D, "start": start, "samples": int(sel.sum())}
    peaks, _ = find_peaks(ls, prominence=prominence)
    res["peaks"] = int(len(peaks))
    if len(peaks) >= 2:
        period = float(np.mean(np.diff(ts[peaks])))  # scripts/lift.py:85-92
        res["period_lift_py"] = period
        res["strouhal_lift_py"] = D / (period * U)
        # the same peaks, each refined with a parabola through its three samples
        inner = peaks[(peaks > 0) & (peaks < len(ls) - 1)]
        y0, y1, y2 = ls[inner - 1], ls[inner], ls[inner + 1]
        dt = ts[1] - ts[0]
        tp = ts[inner] + 0.5 * (y0 - y2) / (y0 - 2 * y1 + y2) * dt
        if len(tp) >= 2:
            res["period_refined"] = float((tp[-1] - tp[0]) / (len(tp) - 1))
            res["strouhal_refined"] = D / (res["period_refined"] * U)
    if len(ls) >= 16:
        dt = float(ts[1] - ts[0])
        x = (ls - ls.mean()) * np.hanning(len(ls))
        n = 1 << int(np.ceil(np.log2(len(x) * 16)))
        spec = np.abs(np.fft.rfft(x, n))
        k = int(np.argmax(spec[1:]) + 1)
        if 0 < k < len(spec) - 1:  # parabolic interpolation of the spectral peak
            a, b, c = spec[k - 1], spec[k], spec[k + 1]
            k = k + 0.5 * (a - c) / (a - 2 * b + c)
        f = k / (n * dt)
        res["strouhal_fft"] = f * D / U
    res["cl_amplitude"] = float(0.5 * (ls.max() - ls.min())) if len(ls) else None
    res["cd_mean_from_start"] = float(ds.mean()) if len(ds) else None
    late = t > 1000  # include/LBMIO.h:375,391
    if late.any():
        res["summary_t_gt_1000"] = {
            "mean_cd": float(cd[late].mean()), "cd_range": [float(cd[late].min()), float(cd[late].max())],
            "mean_cl": float(cl[late].mean()), "cl_range": [float(cl[late].min()), float(cl[late].max())],
            "samples": int(late.sum())}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("forces", nargs="?", default="forces.csv")
    ap.add_argument("params", nargs="?", default="simulation_params.csv")
    ap.add_argument("--start", type=int, default=30000)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    p = load_params(a.params)
    res = analyse(load_forces(a.forces), p["inlet_velocity"], 2.0 * p["cylinder_radius"], a.start)
    if a.json:
        print(json.dumps(res))
        return 0
    print("Strouhal Number Calculation:")
    print("  Inlet Velocity (U): %.4f (lattice units)" % res["U"])
    print("  Cylinder Diameter (D): %.1f (lattice units)" % res["D"])
    print("  Steady-state analysis from timestep: %d" % res["start"])
    print("  Number of peaks found: %d" % res["peaks"])
    if "strouhal_lift_py" in res:
        print("  Average Period (T): %.2f (timesteps)" % res["period_lift_py"])
        print("  Strouhal Number (St = f*D/U): %.4f   [scripts/lift.py definition]" % res["strouhal_lift_py"])
        if "strouhal_refined" in res:
            print("  refined peaks: St = %.5f, FFT: St = %.5f" % (res["strouhal_refined"], res.get("strouhal_fft", float("nan"))))
    print("  C_L amplitude: %s, mean C_D (t >= start): %s" % (res["cl_amplitude"], res["cd_mean_from_start"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
