"""CPU restatement of the reference's driver, robot-learning.py:19-117 (the body of update(dt)), for ONE env.  TEST INFRASTRUCTURE.

Follows /root/reference/robot-learning.py:
  calculate_remaining_money                                        :45-50
  training branch: get_next_action_type, the switch to testing
  once the money is gone, the three purchase gates                 :66-103
  testing branch: success within TEST_DISTANCE_THRESHOLD, best
  distance, time-out                                               :104-117
over oracle/env_oracle.py (world) and oracle/robot_oracle.py (agent hooks).  Like the golden run it is pinned against
(tests/golden/loop_golden.npz, produced by the unmodified reference classes under exactly this loop), the wall-clock money term is
`ticks x tick_seconds` and the test time-out is counted in ticks.  The learner is not part of it: `robot.updates` counts the
td3_update calls, the actor's weights are whatever the caller installs in `robot.actor_flat`.
"""
import numpy as np

from . import env_oracle as eo
from .robot_oracle import RobotOracle

STARTING_MONEY, COST_PER_STEP, COST_PER_CPU_SECOND, COST_PER_DEMO, COST_PER_RESET = 100, 0.01, 0.03, 20, 5   # constants.py:43-47
TEST_DISTANCE_THRESHOLD, TEST_TIMEOUT, UPDATE_RATE = 5, 100, 10                                                # constants.py:50, 53, 31

STEP, DEMO, RESET, SWITCH, SKIP, TEST, IDLE = range(7)          # tick kinds (RTD3_TICK_TYPE_* of include/rtd3.h)
_KIND = {"step": STEP, "demo": DEMO, "reset": RESET}


class DriverOracle:
    def __init__(self, speed, angle, goal, region, robot: RobotOracle, env_rng, noise_rng=None, tick_seconds=1.0 / UPDATE_RATE,
                 gates=True, noise_fn=None):
        """env_rng / noise_rng: LegacyMT19937 streams for Environment.reset draws and the exploration noise (the reference draws
        both from numpy's one global stream; pass the same object twice for that).  noise_fn (optional): callable returning the two
        unit normals of a 'step' tick instead of noise_rng (the throughput mode's counter-based generator, oracle/philox.py).
        gates=False: the training branch alone with every purchase going through (money is not looked at)."""
        self.speed, self.angle = speed, angle
        self.goal, self.region = np.asarray(goal, np.float64), np.asarray(region, np.float64)
        self.robot, self.env_rng, self.noise_rng = robot, env_rng, noise_rng if noise_rng is not None else env_rng
        self.tick_seconds = tick_seconds
        self.gates, self.noise_fn = gates, noise_fn
        self.test_timeout_ticks = max(1, int(round(TEST_TIMEOUT / tick_seconds))) if tick_seconds > 0 else 1000
        self.mode = "training"
        self.demos_bought = self.resets_bought = self.steps_bought = 0
        self.ticks = self.test_ticks = 0
        self.test_best_distance = np.inf
        self.penalty = self.success = self.finished = False
        self.state = None
        self.last_reward = self.last_done = None

    def reset_env(self):                                          # environment.py:130-137
        self.state = eo.random_init_state(self.env_rng, self.region)
        return self.state

    def money(self):                                              # robot-learning.py:45-50
        spent = (self.demos_bought * COST_PER_DEMO + self.resets_bought * COST_PER_RESET + self.steps_bought * COST_PER_STEP
                 + (self.ticks * self.tick_seconds) * COST_PER_CPU_SECOND)
        return STARTING_MONEY - spent

    def begin_tick(self):
        """Everything of update(dt) up to (not including) the action of a 'step' / test tick.  Returns the tick kind."""
        if self.finished:
            return IDLE
        if self.mode == "testing":
            return TEST
        action_type = self.robot.get_next_action_type()           # robot-learning.py:68
        money = self.money() if self.gates else np.inf
        self.ticks += 1
        if money < 0:                                             # :70-80
            self.penalty = bool(money < -1.0)
            self.reset_env()
            self.mode = "testing"
            return SWITCH
        if action_type == "reset":                                # :82-87
            if money >= COST_PER_RESET:
                self.reset_env()
                self.resets_bought += 1
                return RESET
            return SKIP
        if action_type == "demo":                                 # :88-94 (the demonstration itself is the caller's business)
            if money >= COST_PER_DEMO:
                self.demos_bought += 1
                return DEMO
            return SKIP
        return STEP if money >= COST_PER_STEP else SKIP           # :95-96

    def action(self, kind):
        """get_next_action_training (two legacy normals, x then y) / get_next_action_testing (robot.py:541-595)."""
        if kind == STEP:
            z = self.noise_fn() if self.noise_fn is not None else [self.noise_rng.gauss(), self.noise_rng.gauss()]
            return self.robot.act(self.state, z)
        return self.robot.act(self.state)

    def finish_tick(self, kind, action, next_state):
        """The rest of a 'step' tick (robot-learning.py:98-101) or of a test tick (:106-117) given the action taken and the state
        the environment returned."""
        next_state = np.asarray(next_state, np.float64)
        if kind == STEP:
            self.last_reward, self.last_done = self.robot.process_transition(self.state, action, next_state)
            self.state = next_state
            self.steps_bought += 1
            return
        assert kind == TEST
        distance = np.linalg.norm(next_state - self.goal)
        self.state = next_state
        self.test_ticks += 1
        if distance <= TEST_DISTANCE_THRESHOLD:
            self.success = self.finished = True
        if distance < self.test_best_distance:
            self.test_best_distance = distance
        if self.test_ticks >= self.test_timeout_ticks:
            self.finished = True

    def tick(self):
        """One whole update(dt) on the oracle world.  Returns (kind, action)."""
        kind = self.begin_tick()
        act = np.zeros(2)
        if kind in (STEP, TEST):
            act = self.action(kind)
            self.finish_tick(kind, act, eo.step_scalar(self.speed, self.angle, self.state, act))
        return kind, act
