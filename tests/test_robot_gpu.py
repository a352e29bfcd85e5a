"""Parity of the CUDA Robot hooks (act / transition / episode state machine) with the reference golden trace and the oracle.

Flags (done, goal reached, stuck) and action types bit-exact; actions and rewards to float32 round-off of float64 references.
"""
import numpy as np
import pytest
import torch

from oracle.mt19937 import LegacyMT19937
from oracle.robot_oracle import RobotOracle

pytestmark = pytest.mark.gpu


def test_act_training_and_testing_vs_reference(pkg, robot_golden):
    g = robot_golden
    robot = pkg.Robot(g["goal"])
    robot.td3_agent.actor_network.load_flat(g["actor_w"])
    robot._noise_scale.fill_(0.75)
    np.random.seed(99)                                     # the single-env robot draws from numpy's global stream (robot.py:640)
    train = np.array([robot.get_next_action_training(s, 100.0) for s in g["act_states"]])
    test = np.array([robot.get_next_action_testing(s) for s in g["act_states"]])
    np.testing.assert_allclose(train, g["act_train"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(test, g["act_test"], rtol=0, atol=5e-5)
    assert ((np.abs(g["act_train"]) == 5) == (np.abs(train) == 5)).all()
    assert train.dtype == np.float64 and train.shape == (64, 2)
    # numpy's stream advanced by exactly the draws the reference makes
    m = LegacyMT19937(99)
    for _ in range(64 * 2):
        m.gauss()
    assert np.random.normal() == m.gauss()
    res = robot.residual_action(g["act_states"][0] - g["goal"])
    np.testing.assert_allclose(res, g["act_residual"][0], rtol=1e-4, atol=1e-4)


def test_act_batched_matches_oracle(pkg, robot_golden):
    g = robot_golden
    n = 3000
    rs = np.random.RandomState(1)
    goals = rs.uniform(5, 95, (n, 2))
    states = rs.uniform(0, 98.9999, (n, 2)).astype(np.float32)
    robot = pkg.Robot(torch.from_numpy(goals).cuda(), seed=11)
    robot.td3_agent.actor_network.load_flat(g["actor_w"])
    z = rs.normal(size=(2, n))
    act = robot.get_next_action_training(torch.from_numpy(states).cuda(), None, noise=torch.from_numpy(z).cuda()).cpu().numpy()
    for i in range(0, n, 37):
        o = RobotOracle(goals[i], g["actor_w"])
        ref = o.act(states[i].astype(np.float64), z[:, i])
        np.testing.assert_allclose(act[i], ref, rtol=0, atol=5e-5)
    # default noise path: per-env legacy streams seeded seed+i
    a2 = robot.get_next_action_training(torch.from_numpy(states).cuda(), None).cpu().numpy()
    for i in (0, 5, 2999):
        m = LegacyMT19937(11 + i)
        o = RobotOracle(goals[i], g["actor_w"])
        np.testing.assert_allclose(a2[i], o.act(states[i].astype(np.float64), [m.gauss(), m.gauss()]), rtol=0, atol=5e-5)


def test_transition_trace_vs_reference(pkg, robot_golden):
    g = robot_golden
    robot = pkg.Robot(g["goal"], seed=0)
    robot.set_demonstration_states(g["demo_states"])
    robot._demo_flag.fill_(1)
    robot._path_length.fill_(int(g["trace_path_length"]))
    T = g["trace_s"].shape[0]
    for t in range(T):
        assert robot.plan_index == g["trace_plan"][t]
        robot.process_transition(g["trace_s"][t], g["trace_a"][t], g["trace_s2"][t], 100.0)
        # the kernel sees float32 states; the reference float64 ones: rewards agree to float32 round-off of ~100-unit distances
        np.testing.assert_allclose(float(robot._reward64[0]), g["trace_reward"][t], rtol=0, atol=2e-4)
        assert bool(robot._done[0]) == g["trace_done"][t]
        assert robot.stuck_flag == g["trace_stuck"][t] and robot.goal_reached == g["trace_reached"][t]
        if robot.plan_index == robot.path_length - 1 or robot.goal_reached or robot.stuck_flag:
            robot._plan_index.zero_(); robot._goal_reached.zero_(); robot._stuck_flag.zero_()
        else:
            robot._plan_index += 1
    rb = robot.memory
    assert len(rb) == T and rb.position == T
    np.testing.assert_allclose(rb.r[:T].cpu().numpy(), g["trace_reward"], rtol=0, atol=2e-4)
    assert ((rb.notdone[:T].cpu().numpy() < 0.5) == g["trace_done"]).all()
    assert (rb.s[:T].cpu().numpy() == g["trace_s"].astype(np.float32)).all()
    assert (rb.s2[:T].cpu().numpy() == g["trace_s2"].astype(np.float32)).all()


def test_transition_batched_flags_bit_exact_vs_oracle(pkg):
    """Many envs, float32-representable inputs so that the float64 oracle sees exactly what the kernel sees."""
    n, T = 512, 24
    rs = np.random.RandomState(3)
    goals = rs.uniform(5, 95, (n, 2))
    demos = rs.uniform(0, 99, (700, 2))
    robot = pkg.Robot(torch.from_numpy(goals).cuda(), seed=5, buffer_size=20000)
    robot.demo_grid_min_points = 1                          # exercise the candidate lists against the oracle's cdist-style min
    robot.set_demonstration_states(demos)
    assert robot._demo_cells is not None
    robot._demo_flag.fill_(1)
    robot._path_length.fill_(9)
    orcs = []
    for i in range(n):
        o = RobotOracle(goals[i])
        o.demonstration_states, o.demo_flag, o.path_length = list(demos), True, 9
        orcs.append(o)
    cur = rs.uniform(10, 90, (n, 2)).astype(np.float32)
    for t in range(T):
        step = (rs.uniform(-5, 5, (n, 2)) * np.where(rs.rand(n, 1) < 0.5, 0.1, 1.0)).astype(np.float32)
        nxt = np.clip(cur + step, 0, 98.9999).astype(np.float32)
        near = rs.rand(n) < 0.05                           # some envs land inside / on the edge of the goal radius
        nxt[near] = (goals[near] + rs.uniform(-3.6, 3.6, (near.sum(), 2))).astype(np.float32)
        robot.process_transition(torch.from_numpy(cur).cuda(), torch.from_numpy(step).cuda(), torch.from_numpy(nxt).cuda(), None)
        rew = robot._reward64.cpu().numpy(); done = robot._done.cpu().numpy().astype(bool)
        stuck = robot._stuck_flag.cpu().numpy().astype(bool); reached = robot._goal_reached.cpu().numpy().astype(bool)
        for i in range(n):
            o = orcs[i]
            r_ref, d_ref = o.process_transition(cur[i].astype(np.float64), step[i].astype(np.float64), nxt[i].astype(np.float64))
            assert done[i] == d_ref and stuck[i] == o.stuck_flag and reached[i] == o.goal_reached, (t, i)
            np.testing.assert_allclose(rew[i], r_ref, rtol=1e-12, atol=1e-9)
        # episode bookkeeping on both sides through the real state machine (td3_update is not the subject here)
        robot.td3_agent.td3_update = lambda mem: None
        types = robot.get_next_action_type(None, None).cpu().numpy()
        for i in range(n):
            assert ("step", "demo", "reset")[types[i]] == orcs[i].get_next_action_type(), (t, i)
        cur = nxt
    assert len(robot.memory) == n * T


def test_state_machine_vs_reference(pkg, robot_golden):
    g = robot_golden
    robot = pkg.Robot(g["goal"], seed=0)
    calls = []
    robot.td3_agent.td3_update = lambda mem: calls.append(1)
    code = {"step": 0, "demo": 1, "reset": 2}
    for t in range(400):
        if t in (200, 300):
            robot._stuck_flag.fill_(1)
        if t == 250:
            robot._goal_reached.fill_(1)
        ty = robot.get_next_action_type(np.zeros(2), 100.0)
        assert code[ty] == g["sm_types"][t], t
        assert robot.num_episodes == g["sm_episodes"][t] and robot.path_length == g["sm_path_len"][t]
        assert robot.current_noise_scale == g["sm_noise"][t]
    assert len(calls) == int(g["sm_updates"])


def test_demonstration_pipeline_single_env(pkg, env_golden):
    """get_demonstration -> process_demonstration: shapes, counts and RNG consumption as in the reference (SURVEY a-8, 3.4)."""
    g = env_golden
    np.random.seed(1707366464)
    env = pkg.Environment(maps=(g["speed"], g["angle"]))
    state = env.reset()
    robot = pkg.Robot(env.goal_state)
    demo_s, demo_a = env.get_demonstration()
    assert demo_s.shape == (200, 2) and demo_a.shape == (200, 2) and demo_s.dtype == np.float32 and demo_a.dtype == np.float32
    # the planner's end point is close to the goal compared with where it started
    d0 = np.linalg.norm(demo_s[0] - env.goal_state)
    last = env.dynamics(demo_s[-1], demo_a[-1])
    assert np.linalg.norm(last - env.goal_state) < 0.5 * d0
    robot.process_demonstration(demo_s, demo_a, 100.0)
    assert len(robot.memory) == 199
    assert len(robot.demonstration_states) == 200 + 3 * (199 * 6 + 1)      # 3 785 per demo (SURVEY: 11 355 after 3 demos)
    assert len(robot.paths_to_draw) == 4
    assert not robot.goal_reached
    assert env.robot_state is state                                          # planning did not move the robot (same object, as in the reference)


def test_batched_trainer_loop_and_masked_push(pkg, env_golden):
    """The lock-step loop of robot-learning.py:66-101 over 300 envs: per-env counters follow the reference's state machine,
    only stepping envs push replay rows, and an update runs when episodes end."""
    g = env_golden
    n = 300
    env = pkg.Environment(num_envs=n, seed=21, maps=(g["speed"], g["angle"]))
    robot = pkg.Robot(env.goal_state, hidden=64, layers=2, seed=3, buffer_size=40000)
    robot.td3_agent.num_epochs = 4
    robot.set_demonstration_states(np.random.RandomState(0).uniform(0, 99, (300, 2)))
    tr = pkg.BatchedTrainer(env, robot)
    orcs = [RobotOracle(np.zeros(2)) for _ in range(3)]
    pushed = 0
    for t in range(60):
        types = tr.tick().cpu().numpy()
        pushed += int((types == 0).sum())
        assert len(robot.memory) == min(pushed, robot.memory.capacity)
    # first four ticks: demo, demo, demo, reset for every env (SURVEY a-11), then stepping
    assert pushed > 0 and robot.num_updates >= 1
    assert int(tr.resets_bought.min()) >= 2                 # the leaving-demo reset and the first time-out (path length 50)
    st = env.robot_state
    assert float(st.min()) >= 0 and float(st.max()) < 100
    rows = robot.memory.s[:len(robot.memory)]
    assert torch.isfinite(rows).all() and torch.isfinite(robot.memory.r[:len(robot.memory)]).all()
    assert torch.isfinite(robot.td3_agent.params).all()


def test_philox_sampler_range_and_determinism(pkg):
    rb = pkg.ReplayBuffer(100000, seed=9)
    k = torch.arange(70000, dtype=torch.float32, device="cuda")
    rb.push(torch.stack([k, k], 1), torch.stack([k, k], 1), k, torch.stack([k, k], 1), torch.zeros(70000, dtype=torch.bool, device="cuda"))
    rb.sampler = "philox"
    a = rb.sample_indices(8192, 6)
    assert a.shape == (6, 8192) and int(a.min()) >= 0 and int(a.max()) < 70000
    assert float(a.float().mean()) == pytest.approx(35000, rel=0.02)
    rb2 = pkg.ReplayBuffer(100000, seed=9)
    rb2.push(torch.stack([k, k], 1), torch.stack([k, k], 1), k, torch.stack([k, k], 1), torch.zeros(70000, dtype=torch.bool, device="cuda"))
    rb2.sampler = "philox"
    assert torch.equal(a, rb2.sample_indices(8192, 6))
    assert not torch.equal(a, rb.sample_indices(8192, 6))   # the counter advances


def test_graphed_trainer_matches_eager_trainer(pkg, env_golden):
    """The CUDA-graph tick replays exactly the eager tick (same kernels, same RNG streams): identical states and replay rows."""
    g = env_golden
    outs = []
    for use_graph in (False, True):
        env = pkg.Environment(num_envs=256, seed=5, maps=(g["speed"], g["angle"]))
        robot = pkg.Robot(env.goal_state, hidden=64, layers=2, seed=3, buffer_size=40000)
        torch.manual_seed(0)
        robot.td3_agent.actor_network.load_flat(np.random.RandomState(0).normal(0, 0.05, robot.td3_agent.actor_network.count()).astype(np.float32))
        robot.episodes_per_update = 10 ** 9                 # no learner update: compare the env/robot pipeline only
        tr = pkg.BatchedTrainer(env, robot, graph=use_graph, check_interval=4)
        for _ in range(70):
            tr.tick()
        rows = len(robot.memory)
        outs.append((env.robot_state.clone(), rows, robot.memory.s[:rows].sum().item(), robot.memory.r[:rows].sum().item(),
                     tr.steps_bought.clone(), tr.resets_bought.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert outs[0][2] == pytest.approx(outs[1][2], rel=1e-6) and outs[0][3] == pytest.approx(outs[1][3], rel=1e-6)
    assert torch.equal(outs[0][4], outs[1][4]) and torch.equal(outs[0][5], outs[1][5])


@pytest.mark.parametrize("m", [64, 700, 11355])
def test_demo_grid_search_is_bit_identical_to_the_full_sweep(pkg, m):
    """robot.py:753: the proximity term is the min over ALL demonstration states.  The candidate lists (rtd3_demo_lists: per
    1 x 1 cell the states that can be nearest somewhere in the cell) must give the very same float64 as the full
    sweep: clustered demo paths (the reference's 3 785 states per demonstration), augmentation noise pushing points outside
    [0,100), queries on cell corners / edges / world borders / far from every demo / outside the world (those sweep the set)."""
    n = 32768
    rs = np.random.RandomState(m)
    t = np.linspace(0, 1, m // 3 + 1)[:, None]
    paths = [np.array([[5.0, 80.0]]) * (1 - t) + np.array([[90.0, 15.0]]) * t + rs.normal(0, s, (t.shape[0], 2)) for s in (0.5, 2.5, 6.0)]
    demos = np.concatenate(paths)[:m]
    demos[:5] = [[-3.0, 50.0], [104.5, 20.0], [50.0, -7.0], [131.0, 140.0], [99.99, 99.99]]      # outside the world / the grid
    goals = np.tile(np.array([[500.0, 500.0]]), (n, 1))                                           # never reached: the demo term always counts
    nxt = rs.uniform(0, 98.9999, (n, 2)).astype(np.float32)
    nxt[:64, 0] = np.arange(64, dtype=np.float32) * 4.0 % 100                                     # on vertical cell edges
    nxt[64:128, 1] = np.float32(98.9999)                                                          # world border
    nxt[128:192] = np.float32(0.0)
    nxt[192:256] = (demos[rs.randint(0, m, 64)]).clip(0, 98.9999).astype(np.float32)              # on top of demo states
    gx, gy = np.meshgrid(np.arange(100, dtype=np.float32), np.arange(100, dtype=np.float32), indexing="ij")
    nxt[256:10256] = np.stack([gx.ravel(), gy.ravel()], 1)                                        # every cell corner
    nxt[10256:10320] = np.nextafter(np.float32(rs.randint(1, 100, (64, 2))), np.float32(0))       # just below a cell corner
    nxt[10320:10336] = [[-0.5, 50.0], [100.0, 50.0], [50.0, -2.0], [50.0, 100.25]] * 4            # outside the grid: full sweep
    nxt[10336:10400] = np.float32(99.99999)                                                       # the largest reset state
    cur = nxt.copy()
    act = np.zeros((n, 2), np.float32)
    out = []
    for min_points in (10 ** 9, 1):                       # full sweep, then the grid
        robot = pkg.Robot(torch.from_numpy(goals).cuda(), seed=5, buffer_size=40000)
        robot.demo_grid_min_points = min_points
        robot.set_demonstration_states(demos)
        assert (robot._demo_cells is None) == (min_points > m)
        robot._demo_flag.fill_(1)
        robot.process_transition(torch.from_numpy(cur).cuda(), torch.from_numpy(act).cuda(), torch.from_numpy(nxt).cuda(), None)
        out.append(robot._reward64.cpu().numpy().copy())
    assert np.array_equal(out[0], out[1])
    # and against cdist-style numpy on a few rows (the reference's own arithmetic, float64)
    for i in (0, 70, 130, 200, 4095, 256, 5000, 10255, 10260, 10320, 10323, 10399):
        p = nxt[i].astype(np.float64)
        d = np.sqrt(((demos - p) ** 2).sum(axis=1)).min()
        ref = -np.linalg.norm(p - goals[i]) + 10 * (-d)
        np.testing.assert_allclose(out[1][i], ref, rtol=1e-13)


def test_demo_candidate_lists_are_small_and_complete(pkg):
    """Structure of rtd3_demo_lists' output for a reference-like set (3 demonstration paths + 3 augmentations each, 11 355 states):
    every cell has at least one candidate, the lists are short (the point of the structure), and every list entry is one of the
    demonstration states."""
    rs = np.random.RandomState(0)
    pts = []
    for _ in range(3):
        a, b = rs.uniform(5, 95, 2), rs.uniform(5, 95, 2)
        t = np.linspace(0, 1, 200)[:, None]
        path = a * (1 - t) + b * t + np.cumsum(rs.normal(0, 0.3, (200, 2)), axis=0)
        pts.append(path)
        for _ in range(3):
            pts.append(np.repeat(path[:-1], 6, axis=0) + rs.normal(0, 2.5, (199 * 6, 2)))
            pts.append(path[-1:] + rs.normal(0, 2.5, (1, 2)))
    demos = np.concatenate(pts)
    assert demos.shape == (11355, 2)
    robot = pkg.Robot(torch.zeros((4, 2), dtype=torch.float64, device="cuda"), seed=5, buffer_size=100)
    robot.set_demonstration_states(demos)
    start = robot._demo_cells.cpu().numpy()
    sizes = np.diff(start)
    assert start[0] == 0 and start[-1] == robot._demo_list.shape[0] and sizes.min() >= 1
    assert sizes.mean() < 20 and sizes.max() < 400
    have = {tuple(r) for r in demos}
    assert all(tuple(r) in have for r in robot._demo_list.cpu().numpy())
