"""GPU parity of the in-place AA variant (LBM_FLAG_AA, one population buffer, csrc/lbm_aa.cu).

Populations (f_next, f_current), forces rows and the stability verdict must equal the oracle
BIT FOR BIT at both step parities; rho/ux/uy are recovered from collision invariants and are
held to the stated tolerance instead (<= 1e-12 relative for rho, <= 1e-13 absolute for u)."""
import numpy as np
import pytest

from oracle import oracle as O
import parity_util as util
from test_gpu_parity import CASES

pytestmark = pytest.mark.gpu
AA = 16


def make(case, variant=None, flags=0, **kw):
    import lbm_b200

    s = lbm_b200.Solver(util.case_to_params(case, flags=AA | flags, **kw))
    if variant is not None:
        s.set_kernel_variant(variant)
    s.initialise()
    return s


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("variant", [0, 1])
def test_aa_steps_match_oracle(name, variant):
    case = CASES[name]
    s, o = make(case, variant), O.Oracle(case)
    util.compare_state(s, o, name + " init", exact=True)
    done = 0
    for n in (1, 2, 3, 4, 11, 100, 101):
        s.step(n - done)
        o.run(n - done)
        done = n
        util.compare_state(s, o, "%s AA v%d N=%d" % (name, variant, n), exact=True, macros_exact=False)
        fx, fy = s.forces()
        assert (fx, fy) == o.forces(), n
        assert abs(s.max_velocity() - o.max_velocity()) <= 1e-13
    ok, bad = s.check_stability()
    assert ok and bad == -1
    s.close()


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("name", ["64x32", "130x34", "70x33_odd_ny", "cyl_on_wall"])
def test_aa_seeded_random_state(name, seed):
    case = CASES[name]
    state = util.random_state(case, seed)
    s = make(case)
    s.upload_f(state, iteration=0)
    o = util.oracle_with_state(case, state)
    assert np.array_equal(s.f_current(), o.f_current)
    for n in (1, 1, 1, 4):
        s.step(n)
        o.run(n)
        util.compare_state(s, o, "%s AA seed %d" % (name, seed), exact=True, macros_exact=False)
    s.close()


@pytest.mark.parametrize("name", ["64x32", "96x48", "cyl_on_wall"])
def test_aa_run_rows_are_the_reference_rows(name):
    case = CASES[name]
    s, o = make(case), O.Oracle(case)
    rows, bad = s.run(61)
    want, obad = o.run(61)
    assert bad == obad == -1
    assert np.array_equal(rows, want)
    s.close()


def test_aa_instability_timestep():
    case = O.Case(nx=512, ny=128, tau=0.52, inlet_velocity=0.1, output_frequency=50)
    s, o = make(case), O.Oracle(case)
    rows, bad = s.run(400)
    want, obad = o.run(400)
    assert obad >= 0 and bad == obad
    assert np.array_equal(rows[:, :3], want[:, :3])
    s.close()


@pytest.mark.parametrize("flags,kw", [
    (1 | 2 | 4 | 8, dict()),                                # fully periodic shear wave
    (1 | 4, dict(body_force_x=1e-6)),                       # Poiseuille: periodic x, walls, body force
    (2, dict()),                                            # periodic y only, channel with cylinder
    (1 | 2, dict()),                                        # both periodic, cylinder kept
])
def test_aa_equals_ab_in_the_extended_modes(flags, kw):
    import lbm_b200

    case = O.Case(nx=96, ny=40, cylinder_x=0.05, cylinder_y=0.95, cylinder_radius=0.15, output_frequency=4, inlet_velocity=0.03)
    rng = np.random.default_rng(9)
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
    st = np.zeros((case.ny + 2, case.nx + 2, 9))
    st[1:-1, 1:-1] = w * (1 + 0.04 * rng.standard_normal((case.ny, case.nx, 9)))
    out = []
    for aa in (0, AA):
        s = lbm_b200.Solver(util.case_to_params(case, flags=flags | aa, **kw))
        s.initialise()
        s.upload_f(st, 0)
        snaps = []
        for n in (1, 1, 1, 10, 11):
            s.step(n)
            snaps.append((s.f_next()[1:-1, 1:-1].copy(), s.f_current()[1:-1, 1:-1].copy(), s.forces(), s.macros()))
        rows, bad = s.run(20)
        snaps.append((s.f_next()[1:-1, 1:-1].copy(), s.f_current()[1:-1, 1:-1].copy(), s.forces(), s.macros()))
        out.append((snaps, rows, bad))
        s.close()
    (a, rows_a, bad_a), (b, rows_b, bad_b) = out
    assert bad_a == bad_b == -1 and np.array_equal(rows_a, rows_b)
    for k, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x[0], y[0]), ("f_next", k)
        assert np.array_equal(x[1], y[1]), ("f_current", k)
        assert x[2] == y[2], ("forces", k)
        assert np.abs(x[3][0] - y[3][0]).max() <= 1e-12 and np.abs(x[3][1] - y[3][1]).max() <= 1e-13
