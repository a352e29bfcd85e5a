"""The fused tick (csrc/rtd3_tick.cu: rtd3_tick_pre -> actor forward -> rtd3_tick_post) against the one-launch-per-hook tick
of `BatchedTrainer` (which tests/test_robot_gpu.py pins against the reference trace and the oracle): every array must end
up bit-identical - robot-learning.py:66-101 per env.  Only the ORDER of the rows in the replay ring may differ (the
compacted push takes its slots from an atomic counter in both forms), so rows are compared as sorted sets."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(pkg, g, n, demos, fused, noise="mt19937", graph=False, check_interval=1, hidden=64, grid_min=4096):
    env = pkg.Environment(num_envs=n, seed=5, maps=(g["speed"], g["angle"]))
    robot = pkg.Robot(env.goal_state, hidden=hidden, layers=2, seed=3, buffer_size=max(40000, 140 * n))     # the ring never wraps in these tests: which rows a
                                                                      # wrap overwrites depends on the (atomic) push order
    robot.td3_agent.actor_network.load_flat(np.random.RandomState(0).normal(0, 0.05, robot.td3_agent.actor_network.count()).astype(np.float32))
    robot.episodes_per_update = 10 ** 9                     # no learner update: the replay row order differs, so would the minibatches
    robot.demo_grid_min_points = grid_min
    if demos is not None:
        robot.set_demonstration_states(demos)
    tr = pkg.BatchedTrainer(env, robot, noise=noise, graph=graph, check_interval=check_interval, fused=fused)
    return env, robot, tr


def _snapshot(env, robot, tr):
    rows = len(robot.memory)
    rb = robot.memory
    table = torch.cat([rb.s[:rows], rb.a[:rows], rb.r[:rows, None], rb.s2[:rows], rb.notdone[:rows, None]], dim=1).double().cpu().numpy()
    table = table[np.lexsort(table.T[::-1])]
    arrays = {
        "state": env._state, "state64": env._state64, "env_pos": env._bank.pos, "noise_pos": robot._bank.pos,
        "num_episodes": robot._num_episodes, "noise_scale": robot._noise_scale, "path_length": robot._path_length,
        "plan_index": robot._plan_index, "hist": robot._hist, "hist_count": robot._hist_count, "hist_head": robot._hist_head,
        "goal_reached": robot._goal_reached, "demo_flag": robot._demo_flag, "stuck_flag": robot._stuck_flag, "type": robot._type,
        "update": robot._update, "any_update": robot._any_update, "reward": robot._reward, "reward64": robot._reward64,
        "done": robot._done, "action": robot._action, "prev": tr._prev, "steps": tr.steps_bought, "resets": tr.resets_bought,
    }
    return {k: v.clone() for k, v in arrays.items()}, rows, table


def _demo_path(m, seed=0):
    rs = np.random.RandomState(seed)
    t = np.linspace(0, 1, m)[:, None]
    return np.array([[5.0, 80.0]]) * (1 - t) + np.array([[90.0, 15.0]]) * t + rs.normal(0, 2.5, (m, 2))


@pytest.mark.parametrize("n,m,grid_min", [(3, 0, 4096), (33, 300, 4096), (512, 512, 4096), (5000, 0, 4096), (4096, 6000, 1), (777, 14000, 1)])
def test_fused_tick_is_bit_identical_to_the_hook_by_hook_tick(pkg, env_golden, n, m, grid_min):
    """m = 0: no demonstrations; m below grid_min: full sweep from shared memory; above: the grid search with the points staged
    in shared memory (<= 13 000) or read from L2 (14 000).  130 ticks cover the three 'demo' ticks, the leaving-demo reset, the
    first time-outs (path length 50, then 70) and goal / stuck resets."""
    demos = _demo_path(m) if m else None
    snaps = []
    for fused in (False, True):
        env, robot, tr = _build(pkg, env_golden, n, demos, fused, grid_min=grid_min)
        mid = None
        for k in range(130):
            tr.tick()
            if k == 60:
                mid = _snapshot(env, robot, tr)
        snaps.append((mid, _snapshot(env, robot, tr)))
    for which in (0, 1):
        (a, rows_a, tab_a), (b, rows_b, tab_b) = snaps[0][which], snaps[1][which]
        for k in a:
            assert torch.equal(a[k], b[k]), (k, which)
        assert rows_a == rows_b and rows_a > 0
        assert np.array_equal(tab_a, tab_b)
    assert int(snaps[1][1][0]["resets"].min()) >= 2          # every env left the demo phase and timed out at least once


def test_fused_tick_testing_mode_without_noise(pkg, env_golden):
    """RTD3_TICK_NOISE_NONE is get_next_action_testing's action (robot.py:575-595): compose without a noise term."""
    n = 300
    env, robot, tr = _build(pkg, env_golden, n, None, True)
    robot._num_episodes.fill_(10)
    robot._demo_flag.fill_(1)                                # past the demo phase: every env steps
    state = env.robot_state.clone()
    expect = robot.get_next_action_testing(state).clone()
    nxt = env.dynamics(state, expect)
    L, lib = pkg._lib.lib(), pkg._lib
    t = tr._tick_state()
    lib.check(L.rtd3_tick_pre(lib.ctypes.byref(t), lib.stream_ptr(env.device)))
    res = robot.td3_agent.forward(pkg.learner.NET_ACTOR, robot._base)
    lib.check(L.rtd3_tick_post(env._handle, lib.ctypes.byref(t), lib.ptr(res), None, lib.TICK_NOISE_NONE, lib.stream_ptr(env.device)))
    assert torch.equal(robot._action.t(), expect)
    assert torch.equal(env.robot_state, nxt)


def test_philox_noise_is_standard_normal_and_counter_based(pkg, env_golden):
    n = 65536
    acts = []
    for seed in (11, 11, 12):
        env, robot, tr = _build(pkg, env_golden, n, None, True, noise="philox")
        tr.philox_seed = seed
        robot._num_episodes.fill_(10)
        robot._demo_flag.fill_(1)
        robot._noise_scale.fill_(0.2)                        # noise = z * 0.2 * 5 = z
        robot.td3_agent.actor_network.load_flat(np.zeros(robot.td3_agent.actor_network.count(), np.float32))   # residual 0
        robot._goal.copy_(env._state.double())               # baseline 0, so action = clip(z, +-5)
        tr.tick()
        first = robot._action.clone()
        robot._goal.copy_(env._state.double())
        robot._plan_index.zero_(); robot._goal_reached.zero_(); robot._stuck_flag.zero_()
        tr.tick()
        acts.append((first, robot._action.clone()))
    z = acts[0][0].double()
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z.abs() < 1).double().mean()) - 0.6827) < 0.005
    assert abs(float((z[0] * z[1]).mean())) < 0.01           # the two components of an env are uncorrelated
    assert torch.equal(acts[0][0], acts[1][0]) and torch.equal(acts[0][1], acts[1][1])     # same seed: same noise
    assert not torch.equal(acts[0][0], acts[0][1])           # the tick counter advances
    assert not torch.equal(acts[0][0], acts[2][0])           # another seed


def test_multi_tick_graph_replays_the_eager_fused_ticks(pkg, env_golden):
    """`run(ticks)` with graph=True replays check_interval fused ticks from ONE graph; Philox noise is a function of the device
    tick counter, so the replay must reproduce the eager ticks exactly."""
    n = 2048
    snaps = []
    for graph in (False, True):
        env, robot, tr = _build(pkg, env_golden, n, _demo_path(400), True, noise="philox", graph=graph, check_interval=8)
        tr.run(8 * 9 + 3)                                   # nine replays of the 8-tick graph + three single ticks
        snaps.append(_snapshot(env, robot, tr))
        assert tr.ticks == 75
        if graph:
            assert tr._graph_k is not None and tr._graph_k_launches == 8 * 3
    (a, rows_a, tab_a), (b, rows_b, tab_b) = snaps
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert rows_a == rows_b and np.array_equal(tab_a, tab_b)


def test_fused_tick_with_learner_updates_runs(pkg, env_golden):
    n = 1024
    env, robot, tr = _build(pkg, env_golden, n, _demo_path(300), True, noise="philox", graph=True, check_interval=8, hidden=128)
    robot.episodes_per_update = n
    robot.td3_agent.batch_size = 256
    robot.td3_agent.num_epochs = 4
    robot.memory.sampler = "philox"
    tr.run(160)
    assert robot.num_updates >= 1
    assert torch.isfinite(robot.td3_agent.params).all()
    st = env.robot_state
    assert float(st.min()) >= 0 and float(st.max()) < 100


def test_tick_argument_errors(pkg, env_golden):
    env, robot, tr = _build(pkg, env_golden, 64, None, True)
    L, lib = pkg._lib.lib(), pkg._lib
    t = tr._tick_state()
    assert L.rtd3_tick_post(env._handle, lib.ctypes.byref(t), None, None, 0, None) == -1          # null residual
    res = torch.zeros((64, 2), device="cuda")
    assert L.rtd3_tick_post(env._handle, lib.ctypes.byref(t), lib.ptr(res), None, lib.TICK_NOISE_GIVEN, None) == -1   # noise missing
    assert L.rtd3_tick_post(env._handle, lib.ctypes.byref(t), lib.ptr(res), None, 7, None) == -1
    t.rp_total = None
    assert L.rtd3_tick_pre(lib.ctypes.byref(t), None) == -1
    with pytest.raises(ValueError):
        pkg.BatchedTrainer(env, robot, noise="philox", fused=False)


@pytest.mark.parametrize("n,m,hidden", [(2048, 400, 64), (8192, 11355, 256), (20000, 0, 128), (333, 700, 96)])
def test_multi_tick_kernel_matches_the_three_launch_tick(pkg, env_golden, n, m, hidden):
    """rtd3_tick_run_f16 (check_interval ticks in ONE launch, actor forward inside the kernel) against the fused three-launch tick
    with the same f16 forward: the same device functions run per env in the same order, so every array must be bit-identical.
    20 000 envs: more tiles than SMs (a CTA runs all ticks of one tile, then of the next; the trainer itself only picks the kernel
    for a single wave of tiles, so the test forces it); 333: a ragged last tile."""
    snaps = []
    for multi in (False, True):
        env, robot, tr = _build(pkg, env_golden, n, _demo_path(m) if m else None, True, noise="philox", graph=False, check_interval=8,
                                hidden=hidden, grid_min=64)
        robot.td3_agent.precision = "f16"
        tr.multi_tick_kernel = multi
        if multi and n > 148 * 128:
            tr._multi_tick_ok = lambda: True                # force the kernel beyond one wave of tiles
        assert tr._multi_tick_ok() == multi
        before = pkg._lib.launch_count()
        tr.run(8 * 9 + 3)                                   # nine blocks of 8 ticks + three single ticks
        launches = pkg._lib.launch_count() - before
        snaps.append(_snapshot(env, robot, tr))
        assert tr.ticks == 75 and int(tr._tick_counter[0].item()) == 75
        if multi:
            assert launches <= 9 + 3 * 3 + 2                # one launch per block (+ the one-off weight copies)
    (a, rows_a, tab_a), (b, rows_b, tab_b) = snaps
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert rows_a == rows_b and rows_a > 0 and np.array_equal(tab_a, tab_b)


def test_philox_noise_matches_the_oracle(pkg, env_golden):
    """The normals generated inside rtd3_tick_post against oracle/philox.py (pinned to Philox4x32-10 by the Random123 known-answer
    vectors): with a zero baseline, a zero residual and noise_scale * 5 == 1 the action IS the clipped normal of (seed, tick 1, env)."""
    from oracle.philox import normal2
    n = 4096
    env, robot, tr = _build(pkg, env_golden, n, None, True, noise="philox")
    tr.philox_seed = 0x123456789A                            # both key words in use
    robot._num_episodes.fill_(10)
    robot._demo_flag.fill_(1)
    robot._noise_scale.fill_(0.2)
    robot.td3_agent.actor_network.load_flat(np.zeros(robot.td3_agent.actor_network.count(), np.float32))
    robot._goal.copy_(env._state.double())
    tr.tick()
    got = robot._action.cpu().numpy()                        # [2, n]
    want = np.array([normal2(0x123456789A, 1, e) for e in range(n)]).T
    np.testing.assert_allclose(got, np.clip(want, -5, 5).astype(np.float32), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("multi", [False, True])
def test_captured_ticks_follow_a_replaced_demonstration_set(pkg, env_golden, multi):
    """A captured tick holds the device pointers of the demonstration set: after `set_demonstration_states` the graphs must be
    re-captured (single-tick graph via tick(), eight-tick graph via run()) - compared with eager fused ticks."""
    snaps = []
    for graph in (False, True):
        env, robot, tr = _build(pkg, env_golden, 512, _demo_path(300), True, noise="philox", graph=graph, check_interval=4)
        tr.multi_tick_kernel = False
        advance = tr.run if multi else (lambda k: [tr.tick() for _ in range(k)])
        advance(60)
        robot.set_demonstration_states(_demo_path(900, seed=3) + np.array([[10.0, -20.0]]))
        advance(60)
        snaps.append(_snapshot(env, robot, tr))
    (a, rows_a, tab_a), (b, rows_b, tab_b) = snaps
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert rows_a == rows_b and np.array_equal(tab_a, tab_b)
