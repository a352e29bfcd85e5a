"""Parity of the CUDA Environment path (through the C ABI) with the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): float states within 1e-5 relative -> |d| <= 1e-5 * max(1, |s'|);
init region / goal / reset draws bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import env_oracle as eo
from oracle.mt19937 import LegacyMT19937

pytestmark = pytest.mark.gpu

REL = 1e-5


def assert_state_close(dev, ref):
    dev = np.asarray(dev, dtype=np.float64)
    tol = REL * np.maximum(1.0, np.abs(ref))
    bad = np.abs(dev - ref) > tol
    assert not bad.any(), "max |d| %g at %s" % (np.abs(dev - ref).max(), np.argwhere(bad)[:5])


@pytest.fixture(scope="module")
def env4096(pkg, env_golden):
    return pkg.Environment(num_envs=4096, seed=1707366464, maps=(env_golden["speed"], env_golden["angle"]))


def test_dynamics_vs_reference_golden(env4096, env_golden):
    g = env_golden
    s = torch.from_numpy(g["dyn_states"]).cuda()
    a = torch.from_numpy(g["dyn_actions"]).cuda()
    out = env4096.dynamics(s, a).cpu().numpy()
    assert_state_close(out, g["dyn_next_f64act"])
    assert_state_close(out, g["dyn_next_f32act"])


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4096, 100003, 300000])
def test_step_vs_oracle(pkg, env_golden, n, variant):
    g = env_golden
    env = pkg.Environment(num_envs=n, seed=7, maps=(g["speed"], g["angle"]))
    env.step_variant = variant
    rs = np.random.RandomState(n)
    s = rs.uniform(0, 98.9999, (n, 2)).astype(np.float32)
    a = rs.uniform(-7.5, 7.5, (n, 2)).astype(np.float32)
    if n == 1:
        env.robot_state = s[0]
        out = env.step(a[0])[None]
    else:
        env.robot_state = torch.from_numpy(s).cuda()
        out = env.step(torch.from_numpy(a).cuda()).cpu().numpy()
    ref = eo.step_batch(g["speed"], g["angle"], s.astype(np.float64), a.astype(np.float64))
    assert_state_close(out, ref)


def test_nan_zero_and_clip_edges(pkg, env_golden):
    g = env_golden
    env = pkg.Environment(num_envs=8, seed=1, maps=(g["speed"], g["angle"]))
    s = np.array([[10, 20], [50, 50], [0, 0], [98.9999, 98.9999], [0, 98.9999], [5, 5], [5, 5], [5, 5]], dtype=np.float32)
    a = np.array([[np.nan, 1], [0, 0], [-5, -5], [5, 5], [-9, 9], [np.inf, 0], [0, -np.inf], [1, np.nan]], dtype=np.float32)
    env.robot_state = torch.from_numpy(s).cuda()
    out = env.step(torch.from_numpy(a).cuda()).cpu().numpy()
    ref = eo.step_batch(g["speed"], g["angle"], s.astype(np.float64), a.astype(np.float64))
    assert_state_close(out, ref)
    assert (out[0] == s[0]).all() and (out[7] == s[7]).all()          # NaN action keeps the state (environment.py:125)
    assert (out[1] == s[1]).all()                                      # zero action is the identity
    assert out.min() >= 0 and out.max() <= np.float32(98.9999)
    # pure dynamics propagates the NaN instead
    d = env.dynamics(torch.from_numpy(s).cuda(), torch.from_numpy(a).cuda()).cpu().numpy()
    assert np.isnan(d[0]).all()


def test_empty_batch_is_a_noop(pkg):
    L = pkg._lib.lib()
    env = pkg.Environment(num_envs=4, seed=1)
    rc = L.rtd3_env_step(env._handle, None, None, None, None, 0, 0, None)
    assert rc == 0


def test_mt_bank_raw_and_gauss(pkg):
    n = 37
    bank = pkg.MtBank(n, "cuda").seed(1707366464)
    k = 1400                                              # crosses two regenerations
    got = bank.draw_u32(k).cpu().numpy().view(np.uint32)
    for i in [0, 1, 17, 36]:
        m = LegacyMT19937(1707366464 + i)
        exp = np.array([m.random_uint32() for _ in range(k)], dtype=np.uint32)
        assert (got[:, i] == exp).all()
    gz = bank.draw_gauss(9).cpu().numpy()                 # odd count leaves a cached spare in the bank
    gz2 = bank.draw_gauss(4).cpu().numpy()
    for i in [0, 5, 36]:
        m = LegacyMT19937(1707366464 + i)
        for _ in range(k):
            m.random_uint32()
        exp = np.array([m.gauss() for _ in range(13)])
        np.testing.assert_allclose(np.concatenate([gz[:, i], gz2[:, i]]), exp, rtol=4e-16, atol=0)


def test_init_goal_region_and_reset_bit_exact_vs_golden(pkg, env_golden):
    g = env_golden
    n = g["goals"].shape[0]
    env = pkg.Environment(num_envs=n, seed=int(g["seed_base"]), maps=(g["speed"], g["angle"]))
    assert (env.goal_state.cpu().numpy() == g["goals"]).all()
    assert (env.robot_init_region.cpu().numpy() == g["regions"]).all()
    env.reset()
    assert (env._state64.t().cpu().numpy() == g["reset1"]).all()
    assert (env.robot_state.cpu().numpy() == g["reset1"].astype(np.float32)).all()
    mask = torch.zeros(n, dtype=torch.bool)
    mask[::2] = True
    env.reset(mask)                                       # only even envs redraw
    s64 = env._state64.t().cpu().numpy()
    assert (s64[::2] == g["reset2"][::2]).all() and (s64[1::2] == g["reset1"][1::2]).all()


def test_init_goal_region_vs_oracle_many_seeds(pkg):
    n = 1000
    env = pkg.Environment(num_envs=n, seed=12345)
    goals = env.goal_state.cpu().numpy()
    regions = env.robot_init_region.cpu().numpy()
    env.reset()
    starts = env._state64.t().cpu().numpy()
    for i in range(0, n, 13):
        rng = LegacyMT19937(12345 + i)
        goal, region, _ = eo.set_init_and_goal(rng)
        assert (goals[i] == goal).all() and (regions[i] == region).all()
        assert (starts[i] == eo.random_init_state(rng, region)).all()


def test_single_env_mirrors_numpy_global_stream(pkg, env_golden):
    g = env_golden
    np.random.seed(1707366464)                            # robot-learning.py:19-22
    env = pkg.Environment(maps=(g["speed"], g["angle"]))
    state = env.reset()
    assert (env.robot_init_region == g["kat_region"]).all()
    assert (env.goal_state == g["kat_goal"]).all()
    assert (state == g["kat_reset"]).all() and state.dtype == np.float64
    assert (env.reset() == g["reset2"][0]).all()
    # numpy's own stream has advanced exactly as under the reference
    rs = np.random.RandomState(1707366464)
    m = LegacyMT19937(1707366464)
    goal, region, _ = eo.set_init_and_goal(m)
    eo.random_init_state(m, region); eo.random_init_state(m, region)
    assert np.random.uniform() == m.random_double()
    # single-env numpy call surface
    nxt = env.step(np.array([3.0, -4.0]))
    assert isinstance(nxt, np.ndarray) and nxt.shape == (2,)
    ref = eo.step_scalar(g["speed"], g["angle"], g["reset2"][0].astype(np.float32).astype(np.float64), np.array([3.0, -4.0]))
    assert_state_close(nxt, ref)


def test_rollout_teacher_forced_vs_golden_trajectory(pkg, env_golden):
    g = env_golden
    T, m = g["roll_actions"].shape[:2]
    env = pkg.Environment(num_envs=m, seed=1, maps=(g["speed"], g["angle"]))
    for t in range(T):
        s = torch.from_numpy(g["roll_traj"][t].astype(np.float32)).cuda()
        a = torch.from_numpy(g["roll_actions"][t]).cuda()
        env.robot_state = s
        out = env.step(a).cpu().numpy()
        ref = eo.step_batch(g["speed"], g["angle"], g["roll_traj"][t].astype(np.float32).astype(np.float64),
                            g["roll_actions"][t].astype(np.float64))
        assert_state_close(out, ref)


@pytest.mark.parametrize("plain", [0, 1, 2])
@pytest.mark.parametrize("n,T", [(8, 32), (4096, 50), (1000, 17), (70000, 33), (65536, 40), (4096, 1000), (32, 16), (64, 15)])
def test_rollout_equals_repeated_steps(pkg, env_golden, n, T, plain):
    """All rollout kernels: 0 = automatic (warp-pair TMA kernel for latency-bound sizes, single-warp TMA for large ones),
    1 = per-thread cp.async ring (also the fallback for n % 4 != 0), 2 = single-warp TMA kernel everywhere."""
    g = env_golden
    if T == 1000 and plain == 1:
        pytest.skip("long case once per TMA variant")
    pkg._lib.lib().rtd3_env_force_plain_rollout(plain)
    try:
        _rollout_vs_steps(pkg, g, n, T)
    finally:
        pkg._lib.lib().rtd3_env_force_plain_rollout(0)


def _rollout_vs_steps(pkg, g, n, T):
    env_a = pkg.Environment(num_envs=n, seed=3, maps=(g["speed"], g["angle"]))
    env_b = pkg.Environment(num_envs=n, seed=3, maps=(g["speed"], g["angle"]))
    env_a.reset(); env_b.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    planes = (torch.rand((T, 2, n), device="cuda", generator=gen) * 15 - 7.5)
    if T >= 17:                                            # NaN / inf actions inside a chunk: the careful path
        planes[5, 0, n // 2] = float("nan")
        planes[16, 1, 0] = float("inf")
        planes[3, 1, n - 1] = float("nan")
    traj = env_a.rollout(planes.permute(0, 2, 1))
    step_every = 1 if T <= 60 else 37
    for t in range(T):
        if T > 60 and t % step_every != 0:
            env_b.step(planes[t].t())
            continue
        st = env_b.step(planes[t].t())
        assert torch.equal(traj[t], st), t               # same kernel arithmetic: bit-identical
    assert torch.equal(env_a.robot_state, env_b.robot_state)
    env_c = pkg.Environment(num_envs=n, seed=3, maps=(g["speed"], g["angle"]))
    env_c.reset()
    assert env_c.rollout(planes, record=False) is None
    assert torch.equal(env_c.robot_state, env_a.robot_state)


def test_rollout_closed_loop_close_to_reference_float64(pkg, env_golden):
    """Closed loop, float32 device state vs float64 reference: equal to ~1e-4 until a cell boundary flips."""
    g = env_golden
    T, m = g["roll_actions"].shape[:2]
    env = pkg.Environment(num_envs=m, seed=1, maps=(g["speed"], g["angle"]))
    env.robot_state = torch.from_numpy(g["roll_traj"][0].astype(np.float32)).cuda()
    traj = env.rollout(torch.from_numpy(g["roll_actions"]).cuda()).cpu().numpy()
    err = np.abs(traj - g["roll_traj"][1:]).max(axis=(1, 2))
    assert err[0] <= 1e-5 * 100
    assert np.median(err) < 1e-3


def test_full_size_properties(pkg):
    """BASELINE config sizes: invariants that need no oracle."""
    n = 1 << 20
    env = pkg.Environment(num_envs=n, seed=99)
    env.reset()
    s0 = env.robot_state.clone()
    reg = env.robot_init_region
    assert bool(((s0[:, 0] >= reg[:, 0].float()) & (s0[:, 0] <= reg[:, 1].float())).all())
    assert bool(((s0[:, 1] >= reg[:, 2].float()) & (s0[:, 1] <= reg[:, 3].float())).all())
    g = env.goal_state
    mid = torch.stack([(reg[:, 0] + reg[:, 1]) * 0.5, (reg[:, 2] + reg[:, 3]) * 0.5], dim=1)
    assert bool((torch.linalg.norm(g - mid, dim=1) >= 90).all())
    zero = torch.zeros((2, n), device="cuda")
    # zero action: identity up to the boundary clip (a reset can land in (98.9999, 100), environment.py:117)
    assert torch.equal(env.step(zero.t()), s0.clamp(max=float(np.float32(98.9999))))
    s0 = env.robot_state.clone()
    a = torch.rand((2, n), device="cuda") * 20 - 10
    s1 = env.step(a.t()).clone()
    assert float(s1.min()) >= 0 and float(s1.max()) <= float(np.float32(98.9999))
    moved = torch.linalg.norm((s1 - s0).double(), dim=1)
    assert float(moved.max()) <= 5 * np.sqrt(2) + 1e-4     # speed <= 1, |a| <= 5 per axis
    env.robot_state = s0
    env.step_variant = 2
    s2 = env.step(a.t())
    assert torch.equal(s1, s2)                            # both table paths agree bit for bit


def test_rollout_host_equals_device_rollout(pkg, env_golden):
    g = env_golden
    n, T = 4096, 100
    env_a = pkg.Environment(num_envs=n, seed=3, maps=(g["speed"], g["angle"]))
    env_b = pkg.Environment(num_envs=n, seed=3, maps=(g["speed"], g["angle"]))
    env_a.reset(); env_b.reset()
    h_act = (torch.rand((T, 2, n)) * 15 - 7.5).pin_memory()
    out = env_a.rollout_host(h_act, chunks=7, mode="staged")           # staged three-stream pipeline
    ref = env_b.rollout(h_act.cuda())
    assert torch.equal(out.cuda().permute(0, 2, 1), ref)
    assert torch.equal(env_a.robot_state, env_b.robot_state)
    out2 = torch.empty((T, 2, n)).pin_memory()
    assert env_a.rollout_host(h_act, out2, chunks=1, mode="staged") is out2
    ref2 = env_b.rollout(h_act.cuda())
    assert torch.equal(out2.cuda().permute(0, 2, 1), ref2)
    out3 = torch.empty((T, 2, n)).pin_memory()
    before = pkg._lib.launch_count()
    assert env_a.rollout_host(h_act, out3, mode="zero_copy") is out3     # one launch, the kernel itself crosses PCIe both ways
    assert pkg._lib.launch_count() - before == 1
    ref3 = env_b.rollout(h_act.cuda())
    assert torch.equal(out3.cuda().permute(0, 2, 1), ref3)
    assert torch.equal(env_a.robot_state, env_b.robot_state)
    # pinned default: rtd3_env_rollout_host - the copy-engine pipeline as one cached CUDA graph; replayed, and with ragged slices
    for kw in ({}, {}, {"chunks": 7}, {"chunks": 1}, {"chunks": 1000}, {"mode": "graph", "chunks": 5}, {"mode": "graph_out", "chunks": 3}):
        out6 = torch.empty((T, 2, n)).pin_memory() if kw else out3
        out6.zero_()
        assert env_a.rollout_host(h_act, out6, **kw) is out6
        ref6 = env_b.rollout(h_act.cuda())
        assert torch.equal(out6.cuda().permute(0, 2, 1), ref6), kw
        assert torch.equal(env_a.robot_state, env_b.robot_state)
    small = pkg.Environment(num_envs=6, seed=3, maps=(g["speed"], g["angle"]))      # n % 4 != 0: the cp.async rollout kernel in the graph
    small_b = pkg.Environment(num_envs=6, seed=3, maps=(g["speed"], g["angle"]))
    h_small = (torch.rand((33, 2, 6)) * 15 - 7.5).pin_memory()
    assert torch.equal(small.rollout_host(h_small, chunks=4).cuda().permute(0, 2, 1), small_b.rollout(h_small.cuda()))
    out5 = torch.empty((T, 2, n)).pin_memory()
    assert env_a.rollout_host(h_act, out5, chunks=6, mode="hybrid") is out5   # H2D slices by the copy engine, trajectory written to the host by the kernel
    ref5 = env_b.rollout(h_act.cuda())
    assert torch.equal(out5.cuda().permute(0, 2, 1), ref5)
    with pytest.raises(ValueError):
        env_a.rollout_host(h_act.clone(), mode="zero_copy")
    pageable = h_act.clone()                                             # a pageable input falls back to the staged pipeline
    out4 = env_a.rollout_host(pageable, chunks=3)
    ref4 = env_b.rollout(h_act.cuda())
    assert torch.equal(out4.cuda().permute(0, 2, 1), ref4)
