"""Host logic of the temporally blocked path (no GPU): which passes lbm_step launches.  An iteration whose collision is
an output step ends its pass (IOManager::record_forces reads ITS populations, reference include/LBMSolver.h:52-54), the
first iteration after initialise / upload stands alone, everything else pairs up (or triples, at depth 3)."""
import pytest

import lbm_b200


def iterations_of(passes, start):
    out, t = [], start
    for d in passes:
        out.append(list(range(t, t + d)))
        t += d
    return out


@pytest.mark.parametrize("of", [1, 2, 3, 7, 140, 0])
@pytest.mark.parametrize("depth", [1, 2, 3])
@pytest.mark.parametrize("start,fresh", [(0, True), (1, False), (5, False), (140, False), (141, False)])
def test_every_output_iteration_ends_its_pass(of, depth, start, fresh):
    n = 300
    passes = lbm_b200.plan_passes(start, n, of, depth, fresh)
    assert sum(passes) == n and all(1 <= d <= depth for d in passes)
    groups = iterations_of(passes, start)
    if fresh:
        assert groups[0] == [start]  # collides f_current in place: no pull, so it cannot feed a second stage
    for g in groups:
        for t in g[:-1]:
            assert of == 0 or t % of != 0, (g, of)  # an output iteration is never in the middle of a pass


def test_the_reference_cadence_is_seventy_pairs_per_output_period():
    """output_frequency = 140 (reference include/LBMConfig.h:41): after the first iteration every period is 70 passes of 2."""
    passes = lbm_b200.plan_passes(0, 1 + 140 * 3, 140, 2, True)
    assert passes[0] == 1 and passes[1:] == [2] * (70 * 3)
    # an odd period costs one single pass per period
    passes = lbm_b200.plan_passes(1, 7 * 4, 7, 2, False)
    assert passes == [2, 2, 2, 1] * 4
    # depth 3 with the reference cadence: 140 = 46 * 3 + 2
    passes = lbm_b200.plan_passes(1, 140, 140, 3, False)
    assert passes == [3] * 46 + [2]


def test_bad_arguments():
    with pytest.raises(lbm_b200.LbmError):
        lbm_b200.plan_passes(0, 10, 140, 4)
