"""Exploration (not a test): periodic shear-wave decay and body-force Poiseuille on the GPU."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lbm_b200 as L

# 1. shear wave
for ny in (64, 128):
    p = L.SimulationParams(nx=96, ny=ny, tau=0.6, inlet_velocity=0.01, output_frequency=0, flags=1 | 2 | 4 | 8)
    s = L.Solver(p); s.initialise()
    rho0, ux0, uy0 = s.macros()
    m0 = rho0.sum()
    N = 2000
    s.step(N)
    rho, ux, uy = s.macros()
    # macros lag one step: they are the moments of f_current(N-1) -> time N-1
    y = np.arange(ny)
    k = 2 * np.pi / ny
    nu = (0.6 - 0.5) / 3
    amp = 2 * (ux.mean(axis=1) * np.sin(k * y)).mean()
    print("shear ny=%d: amp/u0=%.8f analytic(N-1)=%.8f analytic(N)=%.8f  mass drift=%.3e  max|uy|=%.2e xvar=%.2e" % (
        ny, amp / 0.01, np.exp(-nu * k * k * (N - 1)), np.exp(-nu * k * k * N), abs(rho.sum() - m0) / m0, np.abs(uy).max(),
        np.abs(ux - ux.mean(axis=1, keepdims=True)).max()))
    print("  stable", s.check_stability())
    s.close()

# 2. Poiseuille: periodic x, walls, body force
for (nx, ny, F, steps) in ((64, 32, 1e-6, 60000), (64, 64, 1e-7, 250000)):
    p = L.SimulationParams(nx=nx, ny=ny, tau=0.6, inlet_velocity=0.0, output_frequency=0, flags=1 | 4, body_force_x=F)
    s = L.Solver(p); s.initialise()
    t0 = time.time(); s.step(steps); s.sync(); dt = time.time() - t0
    rho, ux, uy = s.macros()
    u = ux.mean(axis=1)
    nu = (0.6 - 0.5) / 3
    d2 = u[:-2] - 2 * u[1:-1] + u[2:]
    print("poiseuille %dx%d F=%g steps=%d (%.1fs): umax=%.6e  d2u interior mean=%.6e  -F/nu=%.6e -F/(tau nu)=%.6e" % (
        nx, ny, F, steps, dt, u.max(), d2[2:-2].mean(), -F / nu, -F / (0.6 * nu)))
    print("  u[0..3]=", u[:4], " sym err=", np.abs(u - u[::-1]).max(), "xvar", np.abs(ux - u[:, None]).max())
    # best-fit wall offset: u = a*(y - y0)*(y1 - y), fit quadratic
    yy = np.arange(ny)
    c = np.polyfit(yy[1:-1], u[1:-1], 2)
    roots = np.roots(c)
    print("  parabola roots (wall positions):", roots, " curvature*nu=", 2 * c[0] * nu)
    for (w0, w1, name) in ((0.0, ny - 1.0, "walls at nodes"), (-0.5, ny - 0.5, "half-way")):
        G = -2 * c[0]
        ana = 0.5 * G * (yy - w0) * (w1 - yy)
        print("  RMSE/umax vs %s: %.5f" % (name, np.sqrt(((u - ana) ** 2).mean()) / u.max()))
    s.close()
