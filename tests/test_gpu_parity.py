"""GPU parity tests: the CUDA path, called through the C-ABI (include/lbm_b200.h), against the
pinned CPU oracle on the same inputs.  Run with `-m gpu` on a B200.

Tolerance (written here as the contract demands): populations and rho <= 1e-12 relative,
velocities <= 1e-13 absolute (tests/util.py).  Because the kernels are compiled without FMA
contraction and keep the reference's operation order, the results are in fact bit-identical to
the strict-IEEE oracle; `test_bit_identical_*` asserts that stronger property separately.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
import parity_util as util

pytestmark = pytest.mark.gpu

CASES = {
    "64x32": O.Case(nx=64, ny=32, cylinder_x=0.3, cylinder_radius=0.2, output_frequency=3),
    "96x48": O.Case(nx=96, ny=48, cylinder_radius=0.15, output_frequency=7, inlet_velocity=0.05),
    "130x34": O.Case(nx=130, ny=34, cylinder_x=0.25, cylinder_y=0.4, cylinder_radius=0.12, output_frequency=11, tau=0.7),
    "70x33_odd_ny": O.Case(nx=70, ny=33, cylinder_x=0.3, cylinder_radius=0.15, output_frequency=5),
    "256x64": O.Case(nx=256, ny=64, output_frequency=140),
    "cyl_on_wall": O.Case(nx=80, ny=40, cylinder_x=0.5, cylinder_y=0.1, cylinder_radius=0.2, output_frequency=4),
    # obstacle cut by the inlet column / by the outlet-top corner; a one-cell obstacle; degenerate lattices
    "cyl_at_inlet": O.Case(nx=72, ny=36, cylinder_x=0.03, cylinder_y=0.5, cylinder_radius=0.25, output_frequency=5),
    "cyl_at_outlet_corner": O.Case(nx=72, ny=36, cylinder_x=0.97, cylinder_y=0.9, cylinder_radius=0.3, output_frequency=5),
    "single_solid_cell": O.Case(nx=40, ny=24, cylinder_radius=0.0, output_frequency=3),
    "tiny_6x4": O.Case(nx=6, ny=4, cylinder_radius=0.3, output_frequency=2),
    "tall_8x96": O.Case(nx=8, ny=96, cylinder_x=0.5, cylinder_radius=0.02, output_frequency=9),
}


def make_solver(case, variant=None, **kw):
    import lbm_b200

    s = lbm_b200.Solver(util.case_to_params(case, **kw))
    if variant is not None:
        s.set_kernel_variant(variant)
    s.initialise()
    return s


def test_initial_state_matches_grid_initialise():
    for name, case in CASES.items():
        s, o = make_solver(case), O.Oracle(case)
        assert np.array_equal(s.solid(), o.solid), name
        assert s.info().solid_cells == o.n_solid
        util.compare_state(s, o, name + " init", exact=True)
        s.close()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_steps_match_oracle(name, variant):
    case = CASES[name]
    s, o = make_solver(case, variant), O.Oracle(case)
    done = 0
    for n in (1, 2, 3, 10, 100):
        s.step(n - done)
        o.run(n - done)
        done = n
        util.compare_state(s, o, "%s v%d N=%d" % (name, variant, n))
    ok, bad = s.check_stability()
    assert ok and bad == -1
    s.close()


@pytest.mark.parametrize("name", ["64x32", "130x34", "70x33_odd_ny"])
def test_bit_identical_to_strict_oracle(name):
    case = CASES[name]
    s, o = make_solver(case), O.Oracle(case)
    s.step(57)
    o.run(57)
    util.compare_state(s, o, name + " N=57", exact=True)
    s.close()


@pytest.mark.parametrize("name", ["64x32", "96x48", "cyl_on_wall"])
def test_run_forces_rows_and_csv(name):
    case = CASES[name]
    s, o = make_solver(case), O.Oracle(case)
    rows, bad = s.run(60)
    want, obad = o.run(60)
    assert bad == obad == -1
    assert rows.shape == want.shape
    assert np.array_equal(rows[:, 0], want[:, 0])
    # k_forces adds the link terms in the reference's serial (y, x, i) order: same bits
    assert np.array_equal(rows, want)
    assert O.format_forces_csv(rows) == O.format_forces_csv(want)  # forces.csv, byte for byte
    # record_forces on demand (IOManager::record_forces for the current f_next)
    fx, fy = s.forces()
    ofx, ofy = o.forces()
    assert fx == ofx and fy == ofy
    assert abs(s.max_velocity() - o.max_velocity()) <= 1e-13
    s.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("name", ["64x32", "130x34", "70x33_odd_ny", "cyl_on_wall"])
def test_seeded_random_state(name, seed):
    """Upload a seeded non-equilibrium f_current, advance, compare: exercises every population of
    every cell (a uniform initial state cannot reveal a mis-indexed pull)."""
    case = CASES[name]
    state = util.random_state(case, seed)
    s = make_solver(case)
    s.upload_f(state, iteration=0)
    o = util.oracle_with_state(case, state)
    for n in (1, 1, 5):
        s.step(n)
        o.run(n)
        util.compare_state(s, o, "%s seed %d" % (name, seed))
    s.close()


@pytest.mark.parametrize("name", ["64x32", "cyl_on_wall", "cyl_at_inlet"])
@pytest.mark.parametrize("aa", [0, 16])
def test_list_driven_solid_reset_path(name, aa, monkeypatch):
    """LBM_B200_COLSKIP=0 disables the per-column solid-run table: every column then takes the path
    meant for obstacles whose columns hold several solid runs (bulk kernels process the solid cells,
    the fix-up list resets them every step).  Same bits."""
    monkeypatch.setenv("LBM_B200_COLSKIP", "0")
    case = CASES[name]
    s, o = make_solver(case, flags=aa), O.Oracle(case)
    for n in (1, 1, 1, 20):
        s.step(n)
        o.run(n)
        util.compare_state(s, o, "%s loose aa=%d" % (name, aa), exact=True, macros_exact=(aa == 0))
    rows, bad = s.run(30)
    want, obad = o.run(30)
    assert bad == obad == -1 and np.array_equal(rows, want)
    s.close()


def test_variants_are_bit_identical_to_each_other():
    case = CASES["256x64"]
    state = util.random_state(case, 7)
    outs = []
    for v in (0, 1, 2):
        s = make_solver(case, v)
        s.upload_f(state, 0)
        s.step(25)
        outs.append((s.f_next(), s.f_current(), s.macros()))
        s.close()
    for a in outs[1:]:
        assert np.array_equal(a[0], outs[0][0]) and np.array_equal(a[1], outs[0][1])
        for k in range(3):
            assert np.array_equal(a[2][k], outs[0][2][k])


def test_instability_reported_at_reference_timestep():
    case = O.Case(nx=512, ny=128, tau=0.52, inlet_velocity=0.1, output_frequency=50)
    s, o = make_solver(case), O.Oracle(case)
    rows, bad = s.run(400)
    want, obad = o.run(400)
    assert obad >= 0 and bad == obad
    assert rows.shape == want.shape and np.array_equal(rows[:, :3], want[:, :3])
    ok, first = s.check_stability()
    assert not ok and first == obad
    s.close()


@pytest.mark.parametrize("name", ["sampled_256x64", "sampled_2048x512_default"])
def test_golden_fixtures_of_the_fast_math_reference(golden_dir, name):
    """Against outputs of the reference built with ITS OWN flags (-O3 -ffast-math -mfma)."""
    man = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    case = O.Case(**man["case"])
    gx, gy, m = g["gx"], g["gy"], g["interior"]
    s = make_solver(case)
    done = 0
    for n in man["steps"]:
        s.step(n - done)
        done = n
        fn, fc = s.f_next(), s.f_current()
        rho, ux, uy = s.macros()
        util.assert_close_f(fc[gy, gx, :], g[f"fast_N{n}_f_current"], f"{name} f_current N={n}")
        util.assert_close_f(fn[gy, gx, :], g[f"fast_N{n}_f_next"], f"{name} f_next N={n}")
        util.assert_close_f(rho[gy[m] - 1, gx[m] - 1], g[f"fast_N{n}_rho"], f"{name} rho N={n}")
        util.assert_close_u(ux[gy[m] - 1, gx[m] - 1], g[f"fast_N{n}_ux"], f"{name} ux N={n}")
        util.assert_close_u(uy[gy[m] - 1, gx[m] - 1], g[f"fast_N{n}_uy"], f"{name} uy N={n}")
        assert np.array_equal(fc[gy, gx, :], g[f"strict_N{n}_f_current"]), "not bit-identical to the strict build"
        inner = fc[1:-1, 1:-1, :]
        sums = np.array([inner.sum(), (inner ** 2).sum(), rho.sum(), ux.sum()])
        assert np.allclose(sums, g[f"fast_N{n}_sums"], rtol=1e-12, atol=0)
    rows, bad = make_solver(case).run(man["steps"][-1])
    # forces.csv: byte for byte against the strict-IEEE build of the reference (same summation
    # order, same bits); against the -ffast-math build every digit agrees too, but a lift that
    # cancels to +-1e-16 can print as "-0.00000000" there (its compiler reassociates the sum).
    last = man['steps'][-1]
    csv = O.format_forces_csv(rows).encode()
    assert csv == g[f"strict_N{last}_forces_csv"].tobytes()
    assert csv.replace(b"-0.00000000", b"0.00000000") == \
        g[f"fast_N{last}_forces_csv"].tobytes().replace(b"-0.00000000", b"0.00000000")
    s.close()


def test_quirks_f3_f4():
    """Solid cells keep w in f_next, W/E ghost columns are 0, S/N ghost rows keep eq(1,u_in,0)."""
    case = CASES["64x32"]
    s = make_solver(case)
    s.step(20)
    fn, fc = s.f_next(), s.f_current()
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    ys, xs = np.nonzero(s.solid())
    assert np.array_equal(fn[ys + 1, xs + 1, :], np.broadcast_to(w, (len(ys), 9)))
    assert np.all(fn[1:-1, 0, :] == 0.0) and np.all(fn[1:-1, -1, :] == 0.0)
    assert fc[1, 1, 6] == 0.0
    o = O.Oracle(case)
    assert np.array_equal(fn[0, :, :], o.f_next[0, :, :]) and np.array_equal(fn[-1, :, :], o.f_next[-1, :, :])
    s.close()


def test_error_behaviour():
    import lbm_b200

    with pytest.raises(lbm_b200.LbmError):
        lbm_b200.Solver(lbm_b200.SimulationParams(nx=0, ny=8))
    with pytest.raises(lbm_b200.LbmError):
        lbm_b200.Solver(lbm_b200.SimulationParams(nx=8, ny=8, tau=0.5))
    s = lbm_b200.Solver(lbm_b200.SimulationParams(nx=8, ny=8))
    with pytest.raises(lbm_b200.LbmError):
        s.step(1)  # not initialised
    s.close()


def test_full_size_slab_against_oracle():
    """BASELINE size: the 4096 x 8192 per-GPU slab of the weak-scaling config, seeded random state,
    two iterations, every population of every cell compared with the CPU oracle."""
    case = O.Case(nx=4096, ny=8192)
    state = util.random_state(case, 11)
    s = make_solver(case)
    s.upload_f(state, 0)
    o = util.oracle_with_state(case, state)
    del state
    s.step(2)
    o.run(2)
    fn = s.f_next()
    util.assert_close_f(fn, o.f_next, "4096x8192 f_next")
    assert np.array_equal(fn, o.f_next)
    del fn
    rho, ux, uy = s.macros()
    util.assert_close_f(rho, o.rho, "rho")
    util.assert_close_u(ux, o.ux, "ux")
    util.assert_close_u(uy, o.uy, "uy")
    fx, fy = s.forces()
    ofx, ofy = o.forces()
    assert fx == ofx and fy == ofy  # 7904 links, reference summation order
    s.close()


@pytest.mark.parametrize("aa", [0, 16])
@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 2), (3, 1), (1, 5), (2, 3), (5, 2), (4, 7)])
def test_degenerate_lattices(nx, ny, aa):
    """Lattices where the inlet column is the outlet column, or the bottom wall row is the top one
    (the reference applies both rules, in its serial order), down to one cell; 1 x 5 goes unstable
    at the same timestep as the reference."""
    case = O.Case(nx=nx, ny=ny, cylinder_radius=0.0, cylinder_x=0.9, cylinder_y=0.9, output_frequency=2)
    s, o = make_solver(case, flags=aa), O.Oracle(case)
    rows, bad = s.run(12)
    want, obad = o.run(12)
    assert bad == obad
    assert np.array_equal(rows, want)
    if bad == -1:
        util.compare_state(s, o, "%dx%d aa=%d" % (nx, ny, aa), exact=True, macros_exact=(aa == 0))
    s.close()


def test_written_f_next_of_solid_cells_and_ghost_rows_takes_effect_like_in_the_reference():
    """Grid::f_next is writable in the reference (include/LBMGrid.h:119-121).  At an iteration boundary a write only
    matters in solid cells and S/N ghost rows (every later streaming pulls them); fluid cells are overwritten by the
    next collision.  lbm_upload_f_next reproduces exactly that, bit for bit."""
    case = CASES["96x48"]
    s, o = make_solver(case, 2), O.Oracle(case)
    s.step(6)
    o.run(6)
    fn = s.f_next()
    assert np.array_equal(fn, o.f_next)
    ys, xs = np.nonzero(o.solid)
    edit = fn.copy()
    edit[ys + 1, xs + 1, :] *= 1.0 + 0.01 * np.arange(9)          # every solid cell
    edit[0, :, :] *= 0.97                                          # S ghost row (and its corners)
    edit[-1, 5:20, 3] += 0.002                                     # part of the N ghost row
    edit[10:20, 30:40, :] = 123.0                                  # fluid cells: dead values, must change nothing
    fluid_block_is_fluid = not o.solid[9:19, 29:39].any()
    s.upload_f_next(edit)
    of = o.f_next
    of[ys + 1, xs + 1, :] = edit[ys + 1, xs + 1, :]
    of[0, :, :] = edit[0, :, :]
    of[-1, :, :] = edit[-1, :, :]
    assert fluid_block_is_fluid
    for n in (1, 1, 2, 9):
        s.step(n)
        o.run(n)
        util.compare_state(s, o, "after writing f_next +%d" % n, exact=True)
    s.close()
