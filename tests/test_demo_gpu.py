"""Demonstrations for a batch of envs against oracle/demo_oracle.py, which is pinned to the unmodified
reference's planner and augmentation with their random draws recorded (tests/test_oracle_demo.py).

Bars: every random draw and the streams' positions bit-exact; elite selection, float32 mean / std refit and the best path exact on
the device's own rollouts; the demonstration path teacher-forced per step to 1e-5; demonstration sets, replay rows and the
nearest-demonstration term bit-exact / 1e-12 against the reference's expressions."""
import numpy as np
import pytest
import torch

from oracle import demo_oracle as dm
from oracle import env_oracle as eo
from oracle.mt19937 import LegacyMT19937
from oracle.robot_oracle import RobotOracle

pytestmark = pytest.mark.gpu
ENV_SEED, ROBOT_SEED = 905, 1234
P, T, E = 100, 200, 10


def next_words(bank, k=2):
    """The next k raw words of every stream WITHOUT advancing it (state saved and restored)."""
    saved = (bank.mt.clone(), bank.pos.clone(), bank.has_gauss.clone(), bank.gauss.clone())
    w = bank.draw_u32(k).cpu().numpy().view(np.uint32)
    bank.mt.copy_(saved[0]); bank.pos.copy_(saved[1]); bank.has_gauss.copy_(saved[2]); bank.gauss.copy_(saved[3])
    return w


def test_batched_planner_vs_oracle(pkg, env_golden):
    maps = (env_golden["speed"], env_golden["angle"])
    n = 5
    env = pkg.Environment(num_envs=n, seed=ENV_SEED, maps=maps)
    env.reset()
    state_before = env._state.clone()
    rngs = []
    for i in range(n):
        r = LegacyMT19937(ENV_SEED + i)
        eo.set_init_and_goal(r)
        eo.random_init_state(r, env.robot_init_region[i].cpu().numpy())
        rngs.append(r)
    goal, region = env.goal_state.cpu().numpy(), env.robot_init_region.cpu().numpy()
    t, _ = env._cem_workspace()
    act = lambda: t["actions"].reshape(T, 2, P, n).permute(3, 2, 0, 1).cpu().numpy()      # [n][P][T][2]

    # ---- iteration 0: start draw and +-5 actions from the env's own stream, one launch of P*n rollouts
    env._get_demonstration_batched(0, 1, finish=False)
    start64 = t["start64"].t().cpu().numpy()
    a0 = act()
    for i in range(n):
        assert (eo.random_init_state(rngs[i], region[i]) == start64[i]).all()
        assert (dm.draw_iteration_actions(rngs[i], 0) == a0[i]).all()
    assert torch.equal(env._state, state_before)                       # planning does not move the robots
    # the planner's rollouts ARE the rollout kernel's (parity-tested per step in test_env_gpu): same finals as Environment.rollout
    big = pkg.Environment(num_envs=P * n, seed=1, maps=maps)
    big._state[0].copy_(t["start_x"].repeat(P)); big._state[1].copy_(t["start_y"].repeat(P))
    big.rollout(t["actions"], record=False)
    assert torch.equal(big._state[0], t["x"]) and torch.equal(big._state[1], t["y"])
    means = {}

    def check_refit(actions):
        fx, fy = t["x"].reshape(P, n).cpu().numpy().astype(np.float64), t["y"].reshape(P, n).cpu().numpy().astype(np.float64)
        rew = t["rewards"].cpu().numpy()
        for i in range(n):
            ref_rew = -np.sqrt((fy[:, i] - goal[i, 1]) ** 2 + (fx[:, i] - goal[i, 0]) ** 2)
            np.testing.assert_allclose(rew[i], ref_rew, rtol=1e-15)
            idx, mean, std, best = dm.refit(actions[i], rew[i])
            assert (t["elite"][i].cpu().numpy() == idx).all() and int(t["best"][i]) == best
            np.testing.assert_array_equal(t["mean"][:, :, i].cpu().numpy(), mean)       # float32, numpy's operation order
            np.testing.assert_allclose(t["std"][:, :, i].cpu().numpy(), std, rtol=2e-7, atol=1e-9)
            means[i] = (t["mean"][:, :, i].cpu().numpy(), t["std"][:, :, i].cpu().numpy())
    check_refit(a0)

    # ---- iteration 1: normal(mean[step], std[step]) draws with the device's own mean / std
    env._get_demonstration_batched(1, 2, finish=False)
    a1 = act()
    for i in range(n):
        assert (dm.draw_iteration_actions(rngs[i], 1, *means[i]) == a1[i]).all()
    check_refit(a1)

    # ---- the rest, and the demonstration
    states, actions = env._get_demonstration_batched(2, 4, finish=True)
    a3 = act()
    S, A = states.cpu().numpy(), actions.cpu().numpy()
    words = next_words(env._bank)
    for i in range(n):
        for it in (2, 3):                                              # 2 x 100 x 200 x 2 normals whatever their parameters
            dm.draw_iteration_actions(rngs[i], it, *means[i])
        assert (words[:, i] == [rngs[i].random_uint32(), rngs[i].random_uint32()]).all(), i     # the stream ends where the reference's does
        best = int(t["best"][i])
        assert (A[i] == a3[i][best]).all()
        assert (S[i, 0] == start64[i].astype(np.float32)).all()
        for k in range(T - 1):                                         # teacher-forced per step against the oracle's dynamics
            ref = eo.dynamics_scalar(maps[0], maps[1], S[i, k].astype(np.float64), A[i, k].astype(np.float64))
            assert (np.abs(S[i, k + 1] - ref) <= 1e-5 * np.maximum(1.0, np.abs(ref))).all(), (i, k)
        end = eo.dynamics_scalar(maps[0], maps[1], S[i, -1].astype(np.float64), A[i, -1].astype(np.float64))
        assert np.linalg.norm(end - goal[i]) < 0.75 * np.linalg.norm(S[i, 0] - goal[i])    # it is a plan towards the goal
    # the public call does all of it in one go
    env2 = pkg.Environment(num_envs=n, seed=ENV_SEED, maps=maps)
    env2.reset()
    s2, a2 = env2.get_demonstration()
    assert torch.equal(s2, states) and torch.equal(a2, actions)


def test_batched_process_demonstration_vs_oracle(pkg, env_golden):
    maps = (env_golden["speed"], env_golden["angle"])
    n = 4
    env = pkg.Environment(num_envs=n, seed=ENV_SEED, maps=maps)
    env.reset()
    torch.manual_seed(0)
    robot = pkg.Robot(env.goal_state, hidden=32, layers=2, seed=ROBOT_SEED, buffer_size=5000)
    goal = env.goal_state.cpu().numpy()
    rngs = [LegacyMT19937(ROBOT_SEED + i) for i in range(n)]
    per_demo = 200 + 3 * (199 * 6 + 1)
    held = [[] for _ in range(n)]
    rows_expected = []
    for d in range(2):
        S, A = env.get_demonstration()
        robot.process_demonstration(S, A, None)
        Sn, An = S.cpu().numpy(), A.cpu().numpy()
        sets, count = robot.demonstration_sets()
        assert (count.cpu().numpy() == (d + 1) * per_demo).all()
        for i in range(n):
            aug = dm.augment(rngs[i], Sn[i], An[i])
            held[i] += list(Sn[i].astype(np.float64)) + list(aug)
            # (bit-identical except where the device's float64 log / sqrt inside legacy_gauss differ from glibc's in the last bit)
            np.testing.assert_allclose(sets[i, :(d + 1) * per_demo].cpu().numpy(), np.asarray(held[i]), rtol=1e-15, atol=0)
            assert (sets[i, :(d + 1) * per_demo].cpu().numpy() == np.asarray(held[i])).mean() > 0.999
            rows_expected += dm.demonstration_rows(Sn[i], An[i], goal[i])
    words = next_words(robot._bank)
    for i in range(n):
        assert (words[:, i] == [rngs[i].random_uint32(), rngs[i].random_uint32()]).all()
    rb = robot.memory
    assert len(rb) == 2 * n * 199 == len(rows_expected)
    np.testing.assert_array_equal(rb.s[:len(rb)].cpu().numpy(), np.array([r[0] for r in rows_expected], dtype=np.float32))
    np.testing.assert_array_equal(rb.a[:len(rb)].cpu().numpy(), np.array([r[1] for r in rows_expected], dtype=np.float32))
    np.testing.assert_array_equal(rb.r[:len(rb)].cpu().numpy(), np.array([r[2] for r in rows_expected], dtype=np.float64).astype(np.float32))
    np.testing.assert_array_equal(rb.s2[:len(rb)].cpu().numpy(), np.array([r[3] for r in rows_expected], dtype=np.float32))
    np.testing.assert_array_equal(rb.notdone[:len(rb)].cpu().numpy() < 0.5, np.array([r[4] for r in rows_expected]))
    assert not robot.goal_reached.any()

    # ---- the proximity term of the reward now looks at every env's OWN set (robot.py:753-760), exact nearest state
    robot._demo_flag.fill_(1)
    rs = np.random.RandomState(3)
    for trial in range(6):
        if trial < 3:          # near the demonstrated path (with the augmentation's spread), far from it, at the world's borders
            q = np.stack([np.asarray(held[i])[rs.randint(len(held[i]))] + rs.normal(0, [0.2, 3.0, 15.0][trial], 2) for i in range(n)])
        else:
            q = rs.uniform(0, 98.9999, (n, 2)) * ([1, 1] if trial == 3 else [0, 1] if trial == 4 else [1, 0])
        q = np.clip(q, 0, 98.9999).astype(np.float32)
        prev = rs.uniform(10, 90, (n, 2)).astype(np.float32)
        robot.process_transition(torch.from_numpy(prev).cuda(), torch.zeros((n, 2), device="cuda"), torch.from_numpy(q).cuda(), None, push=False)
        rew = robot._reward64.cpu().numpy()
        for i in range(n):
            o = RobotOracle(goal[i])
            o.demonstration_states, o.demo_flag = held[i], True
            ref = o.compute_reward([q[i].astype(np.float64)])
            assert abs(rew[i] - ref) <= 1e-12 * max(1.0, abs(ref)), (trial, i, rew[i], ref)
        robot._hist_count.zero_(); robot._goal_reached.zero_(); robot._stuck_flag.zero_()

    # ---- a third demonstration processed AFTER the demo phase: rows carry the proximity term (robot.py:709 with demo_flag set)
    S, A = env.get_demonstration()
    before = len(rb)
    robot.process_demonstration(S, A, None)
    Sn, An = S.cpu().numpy(), A.cpu().numpy()
    for i in range(n):
        held[i] += list(Sn[i].astype(np.float64)) + list(dm.augment(rngs[i], Sn[i], An[i]))
        rows = dm.demonstration_rows(Sn[i], An[i], goal[i], demo_flag=True, demo_set=held[i])
        got = rb.r[before + i * 199: before + (i + 1) * 199].cpu().numpy()
        np.testing.assert_allclose(got, np.array([r[2] for r in rows], dtype=np.float64).astype(np.float32), rtol=1e-6)
    assert np.abs(got).max() < 1e3


def test_demonstration_rows_wrap_like_a_sequential_push(pkg, env_golden):
    """More demonstration rows than the ring holds: what survives is what a row-by-row push (robot.py:79-96) would leave."""
    maps = (env_golden["speed"], env_golden["angle"])
    n, cap = 3, 350                                          # 3 x 199 = 597 rows into 350 slots
    env = pkg.Environment(num_envs=n, seed=ENV_SEED, maps=maps)
    env.reset()
    torch.manual_seed(0)
    robot = pkg.Robot(env.goal_state, hidden=32, layers=2, seed=ROBOT_SEED, buffer_size=cap)
    rb = robot.memory
    k = torch.arange(40, dtype=torch.float32, device="cuda")
    rb.push(torch.stack([k, k], 1), torch.stack([k, k], 1), k, torch.stack([k, k], 1), torch.zeros(40, dtype=torch.bool, device="cuda"))
    S, A = env.get_demonstration()
    robot.process_demonstration(S, A, None)
    ring = np.full((cap, 2), np.nan, dtype=np.float32)
    pos = 0
    for row in [np.array([v, v], np.float32) for v in range(40)] + [s for i in range(n) for s in S[i, :-1].cpu().numpy()]:
        ring[pos] = row
        pos = (pos + 1) % cap
    assert len(rb) == cap and rb.position == pos
    np.testing.assert_array_equal(rb.s.cpu().numpy(), ring)


def test_batched_loop_buys_demonstrations(pkg, env_golden):
    """BatchedTrainer(demonstrations=True, scheduler=True): the first three ticks buy a demonstration per env (planner +
    process_demonstration), the fourth leaves the demo phase; afterwards the reward's proximity term uses the per-env sets, in the
    eager tick and in the graph / multi-tick forms alike."""
    maps = (env_golden["speed"], env_golden["angle"])
    n, K = 128, 8
    finals = []
    for form in ("eager", "graph8", "multi"):
        torch.manual_seed(1)
        env = pkg.Environment(num_envs=n, seed=ENV_SEED, maps=maps)
        robot = pkg.Robot(env.goal_state, hidden=64, layers=2, seed=ROBOT_SEED, buffer_size=n * 199 * 3 + 20000)
        robot.episodes_per_update = 1 << 30
        robot.td3_agent.precision = "f16"
        robot.td3_agent.actor_network.output_layer.weight.zero_()
        robot.td3_agent.sync_transposed()
        tr = pkg.BatchedTrainer(env, robot, noise="philox", graph=form != "eager", check_interval=K if form != "eager" else 1, fused=True,
                                scheduler=True, demonstrations=True)
        tr.multi_tick_kernel = form == "multi"
        tr.run(6 * K)
        assert (tr.demos_bought == 3).all() and (robot.demonstration_sets()[1] == 3 * 3785).all()
        assert int(robot.memory._total_dev) == 3 * n * 199 + int(tr.steps_bought.sum())
        assert robot.demo_flag.all() and tr._multi_tick_ok() == (form == "multi")
        finals.append((env._state.clone(), robot._reward64.clone(), tr.steps_bought.clone()))
    for f in finals[1:]:
        assert torch.equal(f[0], finals[0][0]) and torch.equal(f[1], finals[0][1]) and torch.equal(f[2], finals[0][2])
    # the shaped reward really is in use: -distance alone would be larger
    rew = finals[0][1].cpu().numpy()
    gd = -torch.linalg.norm(finals[0][0].t().double() - env.goal_state, dim=1).cpu().numpy()
    stepping = (robot._type == 0).cpu().numpy() & (rew != 50.0)            # stepped in the last tick, goal not reached
    assert stepping.sum() > n // 2
    assert (rew[stepping] <= gd[stepping] + 1e-9).all() and (rew[stepping] < gd[stepping] - 1e-6).any()
