"""Pin oracle/demo_oracle.py (CEM planner draws, elite refit, augmentation, demonstration rows) against the golden produced by the
unmodified reference with its random draws recorded (tests/golden/demo_golden.npz)."""
import numpy as np
import pytest

from oracle import demo_oracle as dm
from oracle import env_oracle as eo
from oracle.mt19937 import LegacyMT19937


@pytest.fixture(scope="module")
def g():
    from conftest import load_golden
    return load_golden("demo_golden.npz")


@pytest.mark.parametrize("k", [0, 1])
def test_planner_draws_and_stream_consumption(g, k):
    rng = LegacyMT19937(int(g["seed_%d" % k]))
    goal, region, _ = eo.set_init_and_goal(rng)
    assert (goal == g["goal_%d" % k]).all() and (region == g["region_%d" % k]).all()
    eo.random_init_state(rng, region)                          # robot-learning.py:22 reset
    start = eo.random_init_state(rng, region)                  # the planner's own start draw (environment.py:151)
    assert (start.astype(np.float32) == g["start_%d" % k].astype(np.float32)).all()
    a0 = dm.draw_iteration_actions(rng, 0)
    assert (a0[:4] == g["it0_actions_%d" % k]).all()
    # iteration 1 with the reference's own mean / std: draws bit-exact, stored float32
    mean, std = g["it1_mean_%d" % k], g["it1_std_%d" % k]
    a1 = dm.draw_iteration_actions(rng, 1, mean, std)
    assert (a1[:2] == g["it1_draws_%d" % k].astype(np.float32)).all()
    # the remaining two iterations consume 2 x 100 x 200 x 2 normals whatever their parameters: the stream ends where the reference's does
    for it in (2, 3):
        dm.draw_iteration_actions(rng, it, mean, std)
    assert rng.random_double() == float(g["uniform_after_plan_%d" % k])


def test_whole_planner_reproduces_the_reference_demonstration(g):
    """Closed loop (float64 states like the reference): the oracle's plan IS the reference's demonstration."""
    k = 1
    speed, angle = eo.synthetic_maps(0)
    rng = LegacyMT19937(int(g["seed_%d" % k]))
    goal, region, _ = eo.set_init_and_goal(rng)
    eo.random_init_state(rng, region)
    states, actions = dm.plan(rng, speed, angle, goal, region)
    np.testing.assert_array_equal(actions, g["demo_actions_%d" % k])
    np.testing.assert_array_equal(states, g["demo_states_%d" % k])


@pytest.mark.parametrize("k", [0, 1])
def test_augmentation_and_rows(g, k):
    S, A = g["demo_states_%d" % k], g["demo_actions_%d" % k]
    rng = LegacyMT19937(int(g["seed_%d" % k]) + 1)
    aug = dm.augment(rng, S, A)
    ref = g["aug_states_%d" % k]
    assert ref.shape[0] == 3 * (199 * 6 + 1) and int(g["n_demo_states_%d" % k]) == 200 + ref.shape[0]
    np.testing.assert_array_equal(aug, ref)
    assert rng.random_double() == float(g["uniform_after_aug_%d" % k])
    rows = dm.demonstration_rows(S, A, g["goal_%d" % k])
    assert len(rows) == 199 == len(g["rows_reward_%d" % k])
    np.testing.assert_array_equal(np.array([r[2] for r in rows], dtype=np.float64), g["rows_reward_%d" % k])
    assert [r[4] for r in rows] == list(g["rows_done_%d" % k])
