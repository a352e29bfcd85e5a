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


def test_candidate_lists_contain_the_nearest_state_for_every_query():
    """The rule behind rtd3_demo_lists (restated in the oracle): a state that another state beats at all four corners of a cell
    cannot be nearest inside it.  Checked against the reference's expression - the min over ALL states (robot.py:753) - on a
    clustered path with augmentation-like noise (some states outside the world) plus uniform states, for queries that include every
    cell corner, cell edges and random interior points."""
    from oracle.robot_oracle import demo_candidate_lists, nearest_demo_distance
    rs = np.random.RandomState(0)
    t = np.linspace(0, 1, 500)[:, None]
    pts = np.concatenate([np.array([[5.0, 80.0]]) * (1 - t) + np.array([[90.0, 15.0]]) * t + rs.normal(0, 2.5, (500, 2)),
                          rs.uniform(-3, 103, (200, 2))])
    grid = 40                                              # a 40 x 40 corner of the world keeps the pure-numpy build short
    lists = demo_candidate_lists(pts, grid=grid)
    sizes = np.array([len(l) for l in lists])
    assert sizes.min() >= 1 and sizes.mean() < 12
    gx, gy = np.meshgrid(np.arange(grid, dtype=np.float64), np.arange(grid, dtype=np.float64), indexing="ij")
    queries = np.concatenate([np.stack([gx.ravel(), gy.ravel()], 1),                                   # cell corners
                              np.stack([gx.ravel() + 0.5, gy.ravel()], 1),                             # edge midpoints
                              rs.uniform(0, grid, (3000, 2)),
                              np.float32(rs.uniform(0, grid, (1000, 2))).astype(np.float64)])          # float32 states, as on the device
    queries = queries[(queries < grid).all(axis=1)]
    for q in queries:
        assert nearest_demo_distance(q, pts, lists, grid=grid) == nearest_demo_distance(q, pts)          # bit-identical float64
