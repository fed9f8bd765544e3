"""Shared helpers of the parity tests: tolerances and oracle <-> engine glue."""
import numpy as np

from oracle import oracle as O

# north_star: "population fields within a stated relative tolerance (<= 1e-12 after N steps in fp64)".
RTOL_F = 1e-12
# rho is a sum of nine O(0.1) populations: same relative bound.  Velocities are DIFFERENCES of
# populations divided by rho, so they carry an absolute error of a few ulps of the populations:
# absolute bound in lattice units (u is O(0.01..0.2)).
ATOL_U = 1e-13


def case_to_params(case, **kw):
    import lbm_b200

    return lbm_b200.SimulationParams(
        tau=case.tau, inlet_velocity=case.inlet_velocity, nx=case.nx, ny=case.ny, output_frequency=case.output_frequency,
        cylinder_x=case.cylinder_x, cylinder_y=case.cylinder_y, cylinder_radius=case.cylinder_radius, **kw)


def assert_close_f(got, want, what):
    err = np.abs(got - want)
    bad = err > RTOL_F * np.abs(want)
    if bad.any():
        idx = np.argwhere(bad)[0]
        raise AssertionError("%s: %d of %d values beyond %.0e relative; first at %s: got %r want %r" % (
            what, int(bad.sum()), bad.size, RTOL_F, tuple(idx), got[tuple(idx)], want[tuple(idx)]))


def assert_close_u(got, want, what):
    err = np.abs(got - want).max()
    assert err <= ATOL_U, "%s: max abs error %.3e > %.0e" % (what, err, ATOL_U)


def compare_state(solver, oracle, what, exact=False, macros_exact=None):
    """Every observable the reference exposes (Grid accessors, include/LBMGrid.h:115-129)."""
    got = {"f_next": solver.f_next(), "f_current": solver.f_current()}
    got["rho"], got["ux"], got["uy"] = solver.macros()
    report = {}
    for k in ("f_next", "f_current", "rho", "ux", "uy"):
        want = getattr(oracle, k)
        report[k] = bool(np.array_equal(got[k], want))
        if (exact and k.startswith("f_")) or (exact if macros_exact is None else macros_exact) and not k.startswith("f_"):
            assert report[k], "%s %s: not bit-identical, max abs diff %.3e" % (what, k, np.abs(got[k] - want).max())
        elif k in ("ux", "uy"):
            assert_close_u(got[k], want, "%s %s" % (what, k))
        else:
            assert_close_f(got[k], want, "%s %s" % (what, k))
    return report


def random_state(case, seed, amplitude=0.05):
    """A seeded, strictly positive, non-equilibrium f_current on the padded grid (AoS)."""
    rng = np.random.default_rng(seed)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    f = w * (1.0 + amplitude * rng.standard_normal((case.ny + 2, case.nx + 2, 9)))
    # the ghost ring of f_current is never written after Grid::initialise (SURVEY.md Appendix A
    # step 6): a reachable state always holds eq(1, u_in, 0) there, and lbm_upload_f ignores it
    o = O.Oracle(case)  # keep alive: f_current is a view of its memory
    init = o.f_current.copy()
    for sl in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
        f[sl] = init[sl]
    return np.ascontiguousarray(f)


def oracle_with_state(case, state, x_start=0, lnx=None):
    o = O.Oracle(case, x_start, lnx)
    o.f_current[...] = state
    return o
