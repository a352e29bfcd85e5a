"""Pin oracle/robot_oracle.py against the golden trace produced by the unmodified reference Robot."""
import numpy as np

from oracle.mt19937 import LegacyMT19937
from oracle.robot_oracle import RobotOracle


def test_act_training_and_testing(robot_golden):
    g = robot_golden
    r = RobotOracle(g["goal"], g["actor_w"])
    r.current_noise_scale = 0.75
    m = LegacyMT19937(99)                                  # np.random.seed(99) in the golden script
    train = np.array([r.act(s, [m.gauss(), m.gauss()]) for s in g["act_states"]])
    test = np.array([r.act(s) for s in g["act_states"]])
    np.testing.assert_allclose(train, g["act_train"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(test, g["act_test"], rtol=0, atol=2e-5)
    assert ((np.abs(g["act_train"]) == 5) == (np.abs(train) == 5)).all()


def test_transition_trace(robot_golden):
    g = robot_golden
    r = RobotOracle(g["goal"])
    r.demonstration_states = list(g["demo_states"])
    r.demo_flag = True
    r.path_length = int(g["trace_path_length"])
    for t in range(g["trace_s"].shape[0]):
        assert r.plan_index == g["trace_plan"][t]
        rew, done = r.process_transition(g["trace_s"][t], g["trace_a"][t], g["trace_s2"][t])
        np.testing.assert_allclose(rew, g["trace_reward"][t], rtol=1e-12)
        assert done == g["trace_done"][t] and r.stuck_flag == g["trace_stuck"][t] and r.goal_reached == g["trace_reached"][t]
        if r.plan_index == r.path_length - 1 or r.goal_reached or r.stuck_flag:
            r.plan_index, r.goal_reached, r.stuck_flag = 0, False, False
        else:
            r.plan_index += 1
    assert g["trace_stuck"].any() and g["trace_reached"].any() and g["trace_done"].any()   # the trace exercises all three


def test_state_machine(robot_golden):
    g = robot_golden
    r = RobotOracle(g["goal"])
    code = {"step": 0, "demo": 1, "reset": 2}
    for t in range(400):
        if t in (200, 300):
            r.stuck_flag = True
        if t == 250:
            r.goal_reached = True
        assert code[r.get_next_action_type()] == g["sm_types"][t]
        assert r.num_episodes == g["sm_episodes"][t] and r.path_length == g["sm_path_len"][t]
        assert r.current_noise_scale == g["sm_noise"][t]
    assert r.updates == int(g["sm_updates"])
    assert list(g["sm_types"][:5]) == [1, 1, 1, 2, 0]       # exactly three demos are bought, then a reset (SURVEY a-11)
