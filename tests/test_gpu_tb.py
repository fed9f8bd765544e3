"""GPU parity of the temporally blocked kernels (kernel variant 2, csrc/lbm_tb.cuh): passes of depth 1, 2 and 3
through the C-ABI against the pinned CPU oracle -- populations, f_current, rho / u (emitted by the pass, rebuilt by
a store-less re-run of the pass, or derived from the previous buffer), forces rows, the stability verdict.
Everything bit-identical: the per-cell arithmetic is the same code as every other kernel's."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
import parity_util as util
from test_gpu_parity import CASES

pytestmark = pytest.mark.gpu


def make(case, depth, **kw):
    import lbm_b200

    s = lbm_b200.Solver(util.case_to_params(case, **kw))
    s.set_kernel_variant(lbm_b200.VARIANT_TB)
    s.set_pass_depth(depth)
    s.initialise()
    assert s.info().pass_depth == depth
    return s


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("depth", [1, 2, 3])
def test_passes_match_oracle_bit_for_bit(name, depth):
    case = CASES[name]
    s, o = make(case, depth), O.Oracle(case)
    done = 0
    for n in (1, 2, 3, 4, 10, 11, 57, 100):  # odd and even distances: every mix of pass depths and observers
        s.step(n - done)
        o.run(n - done)
        done = n
        util.compare_state(s, o, "%s depth %d N=%d" % (name, depth, n), exact=True)
    ok, bad = s.check_stability()
    assert ok and bad == -1
    s.close()


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("depth", [2, 3])
@pytest.mark.parametrize("name", ["64x32", "130x34", "70x33_odd_ny", "cyl_on_wall", "cyl_at_outlet_corner"])
def test_seeded_random_state(name, depth, seed):
    case = CASES[name]
    state = util.random_state(case, seed)
    s = make(case, depth)
    s.upload_f(state, iteration=0)
    o = util.oracle_with_state(case, state)
    for n in (1, 2, 5, 6):
        s.step(n)
        o.run(n)
        util.compare_state(s, o, "%s depth %d seed %d +%d" % (name, depth, seed, n), exact=True)
    s.close()


@pytest.mark.parametrize("depth", [2, 3])
@pytest.mark.parametrize("of", [1, 2, 5, 7, 140])
def test_run_rows_and_macros_at_output_steps(depth, of):
    """lbm_run: forces rows for every output step whatever the pass depth (an output iteration ends its pass), and
    rho / u right after it come from the moments the pass emitted."""
    case = O.Case(nx=96, ny=48, cylinder_radius=0.15, output_frequency=of, inlet_velocity=0.05)
    s, o = make(case, depth), O.Oracle(case)
    for n in (33, 1, 40, 2):
        rows, bad = s.run(n)
        want, obad = o.run(n)
        assert bad == obad == -1
        assert np.array_equal(rows, want), (depth, of, n)
        util.compare_state(s, o, "run depth %d of %d +%d" % (depth, of, n), exact=True)
    fx, fy = s.forces()
    ofx, ofy = o.forces()
    assert fx == ofx and fy == ofy
    assert s.max_velocity() == o.max_velocity()
    s.close()


@pytest.mark.parametrize("depth", [2, 3])
def test_instability_reported_at_reference_timestep(depth):
    case = O.Case(nx=512, ny=128, tau=0.52, inlet_velocity=0.1, output_frequency=50)
    s, o = make(case, depth), O.Oracle(case)
    rows, bad = s.run(400)
    want, obad = o.run(400)
    assert obad >= 0 and bad == obad
    assert rows.shape == want.shape and np.array_equal(rows[:, :3], want[:, :3])
    s.close()


@pytest.mark.parametrize("flags", [1 | 4 | 8, 1 | 2 | 4 | 8, 1, 1 | 2])
def test_periodic_modes_agree_with_the_single_step_kernels(flags):
    """The reference has no periodic mode: variant 2 (wrapped addressing inside the pass) against
    variant 1 (ghost copies, one iteration per launch) of this engine, bit for bit."""
    import lbm_b200

    case = O.Case(nx=96, ny=64, cylinder_x=0.4, cylinder_radius=0.12, output_frequency=9, inlet_velocity=0.03)
    outs = []
    for variant, depth in ((1, 1), (2, 1), (2, 2), (2, 3)):
        s = lbm_b200.Solver(util.case_to_params(case, flags=flags, body_force_x=1e-6 if flags == 1 else 0.0))
        s.set_kernel_variant(variant)
        s.set_pass_depth(depth)
        s.initialise()
        got = []
        for n in (1, 6, 10):
            s.step(n)
            got.append((s.f_next()[1:-1, 1:-1].copy(), s.f_current()[1:-1, 1:-1].copy()) + tuple(a.copy() for a in s.macros()))
        outs.append(got)
        s.close()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


@pytest.mark.parametrize("depth", [2, 3])
def test_full_size_slab_two_passes(depth):
    """BASELINE size (the 4096 x 8192 per-GPU slab): seeded random state, 1 + T + T iterations, every population of
    every cell against the CPU oracle."""
    case = O.Case(nx=4096, ny=8192)
    state = util.random_state(case, 12)
    s = make(case, depth)
    s.upload_f(state, 0)
    o = util.oracle_with_state(case, state)
    del state
    s.step(1 + 2 * depth)
    o.run(1 + 2 * depth)
    fn = s.f_next()
    assert np.array_equal(fn, o.f_next)
    del fn
    rho, ux, uy = s.macros()
    assert np.array_equal(rho, o.rho) and np.array_equal(ux, o.ux) and np.array_equal(uy, o.uy)
    fx, fy = s.forces()
    ofx, ofy = o.forces()
    assert fx == ofx and fy == ofy
    s.close()


def test_tree_force_reduction_equals_the_ordered_sum_to_rounding():
    """LBM_FORCES_TREE: the same link terms through a fixed parallel tree -- deterministic from run to run, within
    1e-14 of the reference-ordered sum (relative to the largest partial sum), rows for every second iteration."""
    case = O.Case(nx=384, ny=192, cylinder_radius=0.12, output_frequency=2, inlet_velocity=0.05)
    runs = []
    for tree in (0, 1, 1):
        s = make(case, 2)
        s.set_force_mode(tree)
        rows, bad = s.run(80)
        assert bad == -1 and len(rows) == 40
        runs.append(rows)
        s.close()
    o = O.Oracle(case)
    want, _ = o.run(80)
    assert np.array_equal(runs[0], want)            # ordered mode: the reference's bits
    assert np.array_equal(runs[1], runs[2])          # tree mode: deterministic
    scale = np.abs(want[:, 1:3]).max()
    assert np.abs(runs[1][:, 1:3] - want[:, 1:3]).max() <= 1e-14 * max(scale, 1.0)
    assert np.array_equal(runs[1][:, 0], want[:, 0])


def test_shared_reciprocal_division_is_the_ieee_division():
    """u = j / rho: the kernels divide both momentum components by the density through ONE refined reciprocal
    (csrc/lbm_cell.cuh div_pair); bit for bit the IEEE quotient on 200 million random operand triples, special values
    included."""
    import lbm_b200

    s = lbm_b200.Solver(lbm_b200.SimulationParams(nx=64, ny=32))
    for seed in (1, 2, 3, 4):
        assert s.selftest_division(50_000_000, seed) == 0
    s.close()


def test_fused_later_stages_give_the_same_bits():
    """k_tb<3,128,0,1,FUSED> (LBM_B200_TB_FUSED=1, read once per process: hence a child process): the parity cases of
    this file with the arithmetic of stages 3 and 2 written side by side."""
    env = dict(os.environ, LBM_B200_TB_FUSED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider", "-k",
                        "test_passes_match_oracle_bit_for_bit or test_seeded_random_state or test_run_rows_and_macros"],
                       env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], r.stdout[-500:]
