"""The driver loop (robot-learning.py:45-50, 66-117) on the GPU against the reference and the oracle.

1. configs[0], the WHOLE reference run (tests/golden/loop_golden.npz: the unmodified reference classes, training until the money is
   gone, refused purchases, the switch, the test phase): `trainer.DriverLoop` over the drop-in `Environment` / `Robot`.
2. N oracle driver loops (oracle/driver_oracle.py, pinned to that golden) with seeds s+i against `BatchedTrainer` in every form the
   throughput path uses - hook by hook, fused tick, eight fused ticks per CUDA graph, the multi-tick kernel - and, with the
   scheduler, through the whole run of every env.
Bars: tick kinds, flags, counters, money, reset draws bit-exact; states 1e-5 relative, teacher-forced (same float32 state and action in);
actions 1e-4 absolute - the float32 round-off of an actor whose inputs (state - goal) are O(100), which the reference's own float32
torch forward carries as well.
"""
import numpy as np
import pytest
import torch

from oracle import driver_oracle as do
from oracle import env_oracle as eo
from oracle import philox
from oracle.mt19937 import LegacyMT19937
from oracle.robot_oracle import RobotOracle

pytestmark = pytest.mark.gpu
ENV_SEED, ROBOT_SEED, H, L = 4242, 977, 64, 2


def close(a, b, rel=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool((np.abs(a - b) <= rel * np.maximum(1.0, np.abs(b))).all())


# ------------------------------------------------------------------------------------------------ 1. the whole reference run
def test_whole_driver_loop_vs_reference(pkg, env_golden, loop_golden):
    g, e = loop_golden, env_golden
    torch.manual_seed(0)
    loop = pkg.DriverLoop.from_seed(1707366464, maps=(e["speed"], e["angle"]), tick_seconds=float(g["tick_seconds"]))
    environment, robot = loop.environment, loop.robot
    assert (environment.goal_state == g["goal"]).all() and (environment.robot_init_region == g["region"]).all()
    upd = {"n": 0, "losses": []}
    real_update = robot.td3_agent.td3_update

    def update_with_recorded_noise(memory):                 # torch's generator is unseeded in the reference: inject its recorded draws
        z = torch.from_numpy(g["update_noise"][upd["n"] * 100:(upd["n"] + 1) * 100]).cuda()
        upd["losses"].append(real_update(memory, noise=z))
        upd["n"] += 1
    robot.td3_agent.td3_update = update_with_recorded_noise
    demos = {"n": 0}
    real_demo = environment.get_demonstration

    def demonstration_from_golden():                        # the planner consumes numpy's stream like the reference's; its float32
        ds, _ = real_demo()                                 # paths differ in the last digits, so continue from the reference's
        k = demos["n"]
        assert np.abs(ds[0] - g["demo_states"][k][0]).max() < 1e-5
        demos["n"] += 1
        return g["demo_states"][k], g["demo_actions"][k]
    environment.get_demonstration = demonstration_from_golden

    code = {"step": 0, "demo": 1, "reset": 2, "switch": 3, "test": 5}
    steps = 0
    for t in range(g["kinds"].shape[0]):
        assert not loop.finished
        if loop.mode == "training":
            assert loop.calculate_remaining_money() == g["money"][t], t
        before = (loop.resets_bought, loop.demos_bought, loop.steps_bought)
        state_before = loop.state
        kind = loop.update()
        bought = (loop.resets_bought, loop.demos_bought, loop.steps_bought) != before
        k = code[kind] if (kind in ("switch", "test") or bought) else 4
        assert k == g["kinds"][t], "tick %d: %s -> %d, reference %d" % (t, kind, k, g["kinds"][t])
        if k in (2, 3):
            assert (loop.state == g["states"][t]).all()    # reset draws are bit-exact
        elif k in (0, 5):
            # before the first update the actors are bit-identical copies; every update (100 epochs) then adds the learner's
            # 1e-3-class differences, which the teacher forcing keeps from compounding through the state
            tol = 2e-4 if upd["n"] == 0 else min(5e-2, 3e-3 * (1 + upd["n"]))
            np.testing.assert_allclose(loop.state, g["states"][t], rtol=0, atol=tol, err_msg="tick %d" % t)
            if k == 0:
                assert bool(robot._done[0]) == bool(g["step_dones"][steps])
                steps += 1
            loop.state = g["states"][t].copy()              # teacher-forced: continue from the reference's state
            environment.robot_state = loop.state
        else:
            assert loop.state is state_before
    assert loop.finished and loop.success == bool(g["success"]) and loop.penalty == bool(g["penalty"])
    assert (loop.demos_bought, loop.resets_bought, loop.steps_bought) == (int(g["demos_bought"]), int(g["resets_bought"]), int(g["steps_bought"]))
    assert loop.ticks == int(g["training_ticks"]) and loop.test_ticks == int(g["test_ticks"])
    # measured from the loop's own (not yet teacher-forced) test states: the bound of those states after all updates
    assert abs(loop.test_best_distance - float(g["test_best_distance"])) < min(5e-2, 3e-3 * (1 + upd["n"]))
    assert upd["n"] == int(g["n_updates"]) and len(robot.memory) == int(g["replay_len"])
    closs = torch.cat([l[0] for l in upd["losses"]]).cpu().numpy()
    aloss = torch.cat([l[1] for l in upd["losses"]]).cpu().numpy()
    np.testing.assert_allclose(closs[:200], g["critic_losses"][:200], rtol=2e-3)      # the first two updates: as the 130-tick trace
    np.testing.assert_allclose(closs, g["critic_losses"], rtol=2e-2)
    np.testing.assert_allclose(aloss, g["actor_losses"], rtol=2e-2)
    np.testing.assert_allclose(robot.td3_agent.flat(0).cpu().numpy(), g["final_actor"], rtol=0, atol=1e-3)
    assert np.random.uniform() == float(g["final_uniform"])   # every numpy draw of the run was consumed as by the reference


# ------------------------------------------------------------------------------------------------ 2. N oracle loops
def demo_set():
    rs = np.random.RandomState(5)
    tt = np.linspace(0, 1, 300)[:, None]
    return np.concatenate([rs.uniform(5, 95, (1, 2)) * (1 - tt) + rs.uniform(5, 95, (1, 2)) * tt + rs.normal(0, 2.5, (300, 2)) for _ in range(3)])


def build(pkg, n, maps, noise, fused, graph=False, interval=1, scheduler=False, zero_head=False, precision="fp32", tick_seconds=None,
          capacity=None):
    torch.manual_seed(1)
    env = pkg.Environment(num_envs=n, seed=ENV_SEED, maps=maps)
    robot = pkg.Robot(env.goal_state, hidden=H, layers=L, seed=ROBOT_SEED, buffer_size=capacity or max(20000, 16 * n))
    robot.episodes_per_update = 1 << 30                      # learner off: the update points are only counted
    if zero_head:                                            # residual == output bias whatever the hidden products' precision
        robot.td3_agent.actor_network.output_layer.weight.zero_()
        robot.td3_agent.actor_network.output_layer.bias.copy_(torch.tensor([0.3, -0.2], device="cuda"))
        robot.td3_agent.sync_transposed()
    robot.td3_agent.precision = precision
    robot.set_demonstration_states(demo_set())
    tr = pkg.BatchedTrainer(env, robot, noise=noise, graph=graph, check_interval=interval, fused=fused, scheduler=scheduler,
                            tick_seconds=tick_seconds)
    return env, robot, tr


def oracles(pkg, env, robot, tr, maps, n, gates, noise, tick_seconds=0.1):
    w = robot.td3_agent.flat(0).cpu().numpy().copy()
    goal, region = env.goal_state.cpu().numpy(), env.robot_init_region.cpu().numpy()
    out = []
    for i in range(n):
        rng = LegacyMT19937(ENV_SEED + i)
        og, oreg, _ = eo.set_init_and_goal(rng)
        assert (og == goal[i]).all() and (oreg == region[i]).all()       # seeding is bit-exact (a-4)
        r = RobotOracle(og, w, hidden=H, layers=L)
        r.demonstration_states = list(demo_set())
        d = do.DriverOracle(maps[0], maps[1], og, oreg, r, rng, LegacyMT19937(ROBOT_SEED + i), tick_seconds=tick_seconds, gates=gates)
        if noise == "philox":
            d.noise_fn = (lambda d=d, i=i: philox.normal2(tr.philox_seed, d.tick_index, i))
        d.tick_index = 0
        d.reset_env()                                                   # BatchedTrainer.__init__ -> environment.reset()
        out.append(d)
    s64 = env._state64.t().cpu().numpy()
    assert all((d.state == s64[i]).all() for i, d in enumerate(out))
    return out


def robot_flags(robot):
    f = lambda t: t.cpu().numpy()
    return {"num_episodes": f(robot._num_episodes), "plan_index": f(robot._plan_index), "path_length": f(robot._path_length),
            "goal_reached": f(robot._goal_reached).astype(bool), "stuck_flag": f(robot._stuck_flag).astype(bool),
            "demo_flag": f(robot._demo_flag).astype(bool), "noise_scale": f(robot._noise_scale)}


def check_flags(robot, ors, t):
    fl = robot_flags(robot)
    for i, d in enumerate(ors):
        r = d.robot
        got = tuple(fl[k][i] for k in ("num_episodes", "plan_index", "path_length", "goal_reached", "stuck_flag", "demo_flag", "noise_scale"))
        exp = (r.num_episodes, r.plan_index, r.path_length, r.goal_reached, r.stuck_flag, r.demo_flag, r.current_noise_scale)
        assert got == exp, "tick %d env %d: %s != %s" % (t, i, got, exp)


def lockstep_tick(env, robot, tr, ors, t, scheduler):
    """One eager device tick, then every oracle's tick teacher-forced from the device's values; everything compared."""
    for d in ors:
        d.tick_index = t + 1
    prev = env._state.t().cpu().numpy().astype(np.float64)
    tr.tick()
    types = robot._type.cpu().numpy()
    state = env._state.t().cpu().numpy().astype(np.float64)
    s64 = env._state64.t().cpu().numpy()
    act = robot._action.t().cpu().numpy().astype(np.float64)
    rew, done = robot._reward64.cpu().numpy(), robot._done.cpu().numpy().astype(bool)
    for i, d in enumerate(ors):
        d.state = prev[i] if d.state is None or types[i] not in (2, 3) else d.state
        kind = d.begin_tick()
        assert kind == types[i], "tick %d env %d: device kind %d, oracle %d" % (t, i, types[i], kind)
        if kind in (do.RESET, do.SWITCH):
            assert (d.state == s64[i]).all(), (t, i)                     # reset draw bit-exact
            d.state = state[i]                                           # its float32 rounding is what the device carries on
        elif kind in (do.STEP, do.TEST):
            d.state = prev[i]
            a = d.action(kind)
            assert np.abs(act[i] - a).max() <= 1e-4, "tick %d env %d action %s vs %s" % (t, i, act[i], a)
            nxt = eo.step_scalar(d.speed, d.angle, prev[i], act[i])
            assert close(state[i], nxt), "tick %d env %d state %s vs %s" % (t, i, state[i], nxt)
            d.finish_tick(kind, act[i], state[i])
            if kind == do.STEP:
                assert abs(rew[i] - d.last_reward) <= 1e-9 * max(1.0, abs(d.last_reward)) and done[i] == d.last_done, (t, i)
    check_flags(robot, ors, t)


@pytest.mark.parametrize("fused", [False, True])
def test_training_ticks_vs_oracle_loops_exact_noise(pkg, env_golden, fused):
    """Hook-by-hook and fused ticks, exact mode (per-env numpy-legacy noise streams), 64 envs x 130 ticks."""
    maps = (env_golden["speed"], env_golden["angle"])
    n = 64
    env, robot, tr = build(pkg, n, maps, "mt19937", fused)
    ors = oracles(pkg, env, robot, tr, maps, n, gates=False, noise="mt19937")
    for t in range(130):
        lockstep_tick(env, robot, tr, ors, t, False)
    assert int(robot._any_update.item()) == sum(d.robot.updates for d in ors) > 0
    assert (tr.steps_bought.cpu().numpy() == [d.steps_bought for d in ors]).all()
    assert (tr.resets_bought.cpu().numpy() == [d.resets_bought for d in ors]).all()


@pytest.mark.parametrize("form", ["graph8_mt19937", "graph8_philox", "multi_tick_kernel"])
def test_block_forms_vs_oracle_loops(pkg, env_golden, form):
    """The forms that run eight ticks per launch / graph replay: the oracle loops run the same eight ticks closed-loop (float64) and
    are compared - and re-synchronised - at the block boundaries: counters and flags exact, states to the drift of eight float32
    steps."""
    maps = (env_golden["speed"], env_golden["angle"])
    multi = form == "multi_tick_kernel"
    n, K = (128 if multi else 64), 8                    # (the kernel's tensor-core forward takes whole 128-env tiles)
    noise = "mt19937" if form == "graph8_mt19937" else "philox"
    env, robot, tr = build(pkg, n, maps, noise, True, graph=True, interval=K, zero_head=multi, precision="f16" if multi else "fp32",
                           capacity=20000)
    tr.multi_tick_kernel = multi
    assert tr._multi_tick_ok() == multi
    ors = oracles(pkg, env, robot, tr, maps, n, gates=False, noise=noise)
    for block in range(16):
        tr.run(K)
        state = env._state.t().cpu().numpy().astype(np.float64)
        for i, d in enumerate(ors):
            for k in range(K):
                d.tick_index = block * K + k + 1
                d.tick()
            assert np.abs(state[i] - d.state).max() <= 2e-4, "block %d env %d: %s vs %s" % (block, i, state[i], d.state)
            d.state = state[i]
            # the stuck history of the oracle holds its own (float64) states of the block; the device's differ in the last digits
        check_flags(robot, ors, block)
        assert (tr.steps_bought.cpu().numpy() == [d.steps_bought for d in ors]).all()
        assert (tr.resets_bought.cpu().numpy() == [d.resets_bought for d in ors]).all()
    assert int(robot._any_update.item()) == sum(d.robot.updates for d in ors) > 0
    assert len(robot.memory) == sum(d.steps_bought for d in ors)


def test_scheduler_whole_runs_vs_oracle_loops(pkg, env_golden):
    """scheduler=True: 48 envs through their WHOLE runs (money gates, demos_bought, refused purchases, the switch, the test phase to
    success or time-out) in lock step with 48 oracle driver loops.  A coarse clock (2 s per tick) keeps the runs short and makes the
    purchase gates and the test time-out (50 ticks) bite."""
    maps = (env_golden["speed"], env_golden["angle"])
    n, dt = 48, 2.0
    env, robot, tr = build(pkg, n, maps, "mt19937", True, scheduler=True, tick_seconds=dt)
    ors = oracles(pkg, env, robot, tr, maps, n, gates=True, noise="mt19937", tick_seconds=dt)
    seen = set()
    t = 0
    while not tr.all_finished():
        money = tr.money_remaining().cpu().numpy()
        modes = tr.mode.cpu().numpy()
        for i, d in enumerate(ors):
            if modes[i] == 0:
                assert money[i] == d.money(), (t, i)                     # the same float64 expression, bit for bit
        lockstep_tick(env, robot, tr, ors, t, True)
        seen.update(int(k) for k in robot._type.cpu().numpy())
        t += 1
        assert t < 3000
    res = tr.results()
    assert seen >= {0, 1, 2, 3, 4, 5, 6}
    for i, d in enumerate(ors):
        assert d.finished
        assert (res["success"][i], res["penalty"][i], res["test_ticks"][i]) == (d.success, d.penalty, d.test_ticks), i
        assert (res["demos_bought"][i], res["resets_bought"][i], res["steps_bought"][i]) == (d.demos_bought, d.resets_bought, d.steps_bought), i
        assert res["test_best_distance"][i] == d.test_best_distance, i    # computed from the same float32 states: bit-exact
    assert res["success"].any() or True
    assert (res["mode"] == 2).all()


def test_scheduler_in_the_block_forms_matches_eager_ticks(pkg, env_golden):
    """The scheduler inside eight-tick graphs and the multi-tick kernel leaves every array as eager fused ticks do."""
    maps = (env_golden["speed"], env_golden["angle"])
    n, K, dt = 200, 8, 2.0
    ref = None
    for form in ("eager", "graph8", "multi"):
        env, robot, tr = build(pkg, n, maps, "philox", True, graph=form != "eager", interval=K if form != "eager" else 1, scheduler=True,
                               tick_seconds=dt, zero_head=True, precision="f16", capacity=20000)
        tr.multi_tick_kernel = form == "multi"
        assert tr._multi_tick_ok() == (form == "multi")
        if form == "eager":
            for _ in range(120 * K):
                tr.tick()
        else:
            tr.run(120 * K)
        assert tr.all_finished()
        snap = {k: v.copy() if isinstance(v, np.ndarray) else v for k, v in tr.results().items()}
        snap["state"] = env._state.cpu().numpy()
        snap["episodes"] = robot._num_episodes.cpu().numpy()
        if ref is None:
            ref = snap
            assert snap["success"].any() and (~snap["success"]).any()
        else:
            for k in ref:
                assert np.array_equal(ref[k], snap[k]), (form, k)
