"""Pin oracle/driver_oracle.py (the restated update(dt) of robot-learning.py) against the golden of the WHOLE reference run:
training until the money is gone, refused purchases, the switch to testing, the test phase (tests/golden/loop_golden.npz)."""
import numpy as np

from oracle import driver_oracle as do
from oracle.robot_oracle import RobotOracle


class _NoRng:
    """Reset draws and exploration noise of the golden run come from numpy's global stream, interleaved with the planner's and the
    sampler's draws: this test is teacher-forced (states and actions from the golden), so the oracle's own draws are never used."""
    def uniform(self, lo, hi):
        return 0.5 * (lo + hi)


def test_driver_oracle_vs_whole_reference_run(loop_golden):
    g = loop_golden
    kinds = g["kinds"]
    assert (np.bincount(kinds, minlength=6)[[do.SWITCH, do.SKIP, do.TEST]] > 0).all()     # the run exercises every branch
    d = do.DriverOracle(None, None, g["goal"], g["region"], RobotOracle(g["goal"]), _NoRng(), tick_seconds=float(g["tick_seconds"]))
    d.state = np.zeros(2)            # robot-learning.py:22 (value irrelevant until the first reset: three 'demo' ticks come first)
    step = 0
    for t in range(kinds.shape[0]):
        assert not d.finished
        training = d.mode == "training"
        money_before = d.money()
        kind = d.begin_tick()
        assert kind == kinds[t], "tick %d: kind %d, reference %d" % (t, kind, kinds[t])
        if training:
            assert money_before == g["money"][t], t                     # the same float64 expression, bit for bit
        if kind in (do.RESET, do.SWITCH):
            d.state = g["states"][t].copy()                               # teacher-forced reset draw
        elif kind == do.STEP:
            d.finish_tick(kind, g["actions"][t], g["states"][t])
            assert d.last_done == g["step_dones"][step]
            step += 1
        elif kind == do.TEST:
            d.finish_tick(kind, g["actions"][t], g["states"][t])
    assert d.finished and d.success == bool(g["success"]) and d.penalty == bool(g["penalty"])
    assert (d.demos_bought, d.resets_bought, d.steps_bought) == (int(g["demos_bought"]), int(g["resets_bought"]), int(g["steps_bought"]))
    assert d.ticks == int(g["training_ticks"]) and d.test_ticks == int(g["test_ticks"])
    assert d.test_best_distance == float(g["test_best_distance"])
    assert d.robot.updates == int(g["n_updates"])


def test_driver_oracle_runs_on_the_oracle_world():
    """Closed loop on the oracle's own world (no teacher forcing): terminates, spends the budget, ends in testing."""
    from oracle import env_oracle as eo
    from oracle import td3_oracle as to
    from oracle.mt19937 import LegacyMT19937
    speed, angle = eo.synthetic_maps(0)
    rng = LegacyMT19937(123)
    goal, region, _ = eo.set_init_and_goal(rng)
    w = to.kaiming_uniform_params(np.random.RandomState(0), 2, 16, 2, 2)
    d = do.DriverOracle(speed, angle, goal, region, RobotOracle(goal, w, hidden=16, layers=2), rng, LegacyMT19937(7), tick_seconds=0.1)
    d.reset_env()
    n = 0
    while not d.finished and n < 5000:
        d.tick()
        n += 1
    assert d.finished and d.mode == "testing" and d.demos_bought == 3 and d.money() < 0.05
    assert d.test_ticks >= 1 and np.isfinite(d.test_best_distance)
