"""Parity of the CUDA replay ring + residual-TD3 learner (through the C ABI) with the reference goldens and the oracle.

Tolerances (BASELINE.json north_star): replay indices bit-exact; losses and Q-values within 1e-3 relative after one update.
"""
import numpy as np
import pytest
import torch

from oracle import td3_oracle as to
from oracle.mt19937 import LegacyMT19937

pytestmark = pytest.mark.gpu
REL = 1e-3


def fill_replay(pkg, g, seed=None):
    rb = pkg.ReplayBuffer(10000, seed=seed)
    n = g["rep_s"].shape[0]
    cu = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).cuda()
    rb.push(cu(g["rep_s"]), cu(g["rep_a"]), cu(g["rep_r"]), cu(g["rep_s2"]), torch.from_numpy(g["rep_d"]).cuda())
    assert len(rb) == n
    return rb


def make_agent(pkg, g, stage="w0", **kw):
    agent = pkg.TD3(pkg.Residual_Actor_Network(), pkg.Residual_Critic_Network(), pkg.Residual_Critic_Network(), **kw)
    names = ["actor", "critic1", "critic2", "t_actor", "t_critic1", "t_critic2"]
    for net, name in enumerate(names):
        agent.flat(net).copy_(torch.from_numpy(g[stage + "_" + name]).cuda())
    return agent


def flat(agent, net):
    return agent.flat(net).cpu().numpy()


def update_agrees(new, old, ref_new, ref_old, frac_ok=0.995):
    d, dr = (new - old).astype(np.float64), (ref_new - ref_old).astype(np.float64)
    assert np.abs(d - dr).max() <= 2.05e-5        # never more than one full Adam step (2*lr) apart
    assert (np.abs(d - dr) <= 2e-6).mean() >= frac_ok


def test_replay_ring_push_wrap_and_gather(pkg, replay_golden):
    g = replay_golden
    rb = pkg.ReplayBuffer(16, seed=0)
    for k in range(40):                                   # single-row pushes, like robot.py:675
        rb.push(np.array([k, k]), np.array([k, -k]), float(k), np.array([k, k + 1]), False)
    assert (rb.s.cpu().numpy() == g["ring_states"]).all()
    assert rb.position == int(g["ring_position"]) and len(rb) == 16
    assert rb.sample(17) is None and bool(g["under_filled_is_none"])
    rb2 = pkg.ReplayBuffer(16, seed=0)                    # same rows in batched pushes of 7 (wraps twice)
    for k0 in range(0, 40, 8):
        k = torch.arange(k0, k0 + 8, dtype=torch.float32, device="cuda")
        rb2.push(torch.stack([k, k], 1), torch.stack([k, -k], 1), k, torch.stack([k, k + 1], 1), torch.zeros(8, dtype=torch.bool, device="cuda"))
    assert torch.equal(rb2.s, rb.s) and torch.equal(rb2.a, rb.a) and torch.equal(rb2.r, rb.r) and torch.equal(rb2.s2, rb.s2)
    assert rb2.position == rb.position


def test_sample_indices_bit_exact_vs_reference(pkg, replay_golden):
    g = replay_golden
    ci = 0
    while "case_%d" % ci in g:
        n, B, seed = [int(v) for v in g["case_%d" % ci]]
        rb = pkg.ReplayBuffer(10000, seed=seed)
        k = torch.arange(n, dtype=torch.float32, device="cuda")
        rb.push(torch.stack([k, torch.full_like(k, 0.5)], 1), torch.stack([torch.ones_like(k), k], 1), -k,
                torch.stack([k + 1, torch.full_like(k, 0.25)], 1), (torch.arange(n, device="cuda") % 50) == 49)
        idx = rb.sample_indices(B, 3).cpu().numpy()       # three consecutive draws in one launch
        assert (idx == g["idx_%d" % ci]).all()
        s, a, r, s2, nd = rb.gather(torch.from_numpy(g["idx_%d" % ci][2]).cuda())
        assert (s.cpu().numpy() == g["rows_s_%d" % ci]).all() and (a.cpu().numpy() == g["rows_a_%d" % ci]).all()
        assert (r.cpu().numpy() == g["rows_r_%d" % ci]).all() and (s2.cpu().numpy() == g["rows_s2_%d" % ci]).all()
        assert ((nd.cpu().numpy() < 0.5) == g["rows_d_%d" % ci]).all()
        ci += 1
    assert ci == 5


def test_sample_uses_numpy_global_stream_when_unseeded(pkg):
    rb = pkg.ReplayBuffer(1000)
    k = torch.arange(300, dtype=torch.float32, device="cuda")
    rb.push(torch.stack([k, k], 1), torch.stack([k, k], 1), k, torch.stack([k, k], 1), torch.zeros(300, dtype=torch.bool, device="cuda"))
    np.random.seed(5)
    s, a, r, s2, d = rb.sample(64)                        # numpy outputs, like the reference
    np.random.seed(5)
    exp = np.random.choice(300, 64, replace=False)
    assert (s[:, 0] == exp).all() and d.dtype == np.bool_
    m = LegacyMT19937(5)
    m.choice_no_replace(300, 64)
    np.random.seed(5)
    rb.sample(64)
    assert np.random.uniform() == m.random_double()       # numpy's stream advanced exactly as under the reference


def test_forward_matches_oracle(pkg, td3_golden):
    g = td3_golden
    agent = make_agent(pkg, g)
    rs = np.random.RandomState(0)
    for B in (1, 7, 100, 1000):
        x = rs.uniform(-50, 50, (B, 2)).astype(np.float32)
        y = agent.actor_network(torch.from_numpy(x).cuda()).cpu().numpy()
        ref, _ = to.actor_forward(g["w0_actor"], x)
        np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-4)
        s = rs.uniform(0, 99, (B, 2)).astype(np.float32)
        a = rs.uniform(-5, 5, (B, 2)).astype(np.float32)
        q = agent.critic_network_2(torch.from_numpy(s).cuda(), torch.from_numpy(a).cuda()).cpu().numpy()
        ref, _ = to.critic_forward(g["w0_critic2"], s, a)
        np.testing.assert_allclose(q, ref, rtol=1e-4, atol=1e-3)


def test_train_critic_one_step_vs_reference(pkg, td3_golden):
    g = td3_golden
    agent, rb = make_agent(pkg, g), fill_replay(pkg, g, seed=0)
    B = g["idx_critic"].shape[0]
    q_out = torch.zeros((2, B), device="cuda")
    y_out = torch.zeros((B,), device="cuda")
    l1, l2 = agent.train_critic(rb, noise=torch.from_numpy(g["noise_critic"]).cuda(), idx=torch.from_numpy(g["idx_critic"]).cuda(),
                                q_out=q_out, y_out=y_out)
    np.testing.assert_allclose([l1, l2], g["critic_losses"], rtol=REL)
    update_agrees(flat(agent, 1), g["w0_critic1"], g["w1_critic1"], g["w0_critic1"])
    update_agrees(flat(agent, 2), g["w0_critic2"], g["w1_critic2"], g["w0_critic2"])
    assert (flat(agent, 0) == g["w1_actor"]).all() and (flat(agent, 4) == g["w1_t_critic1"]).all()
    # Q-values after the update, the north-star observable
    s = torch.from_numpy(g["rep_s"][g["idx_critic"]].astype(np.float32)).cuda()
    a = torch.from_numpy(g["rep_a"][g["idx_critic"]].astype(np.float32)).cuda()
    np.testing.assert_allclose(agent.critic_network_1(s, a).cpu().numpy(), g["q1_after_critic"], rtol=REL, atol=1e-3)
    np.testing.assert_allclose(agent.critic_network_2(s, a).cpu().numpy(), g["q2_after_critic"], rtol=REL, atol=1e-3)
    # targets and pre-update Q agree with the oracle
    orc = to.TD3Oracle(g["w0_actor"], g["w0_critic1"], g["w0_critic2"], g["w0_t_actor"], g["w0_t_critic1"], g["w0_t_critic2"])
    ro = to.ReplayOracle(10000)
    for k in range(g["rep_s"].shape[0]):
        ro.push(g["rep_s"][k], g["rep_a"][k], g["rep_r"][k], g["rep_s2"][k], g["rep_d"][k])
    orc.train_critic(*ro.gather(g["idx_critic"]), g["noise_critic"])
    np.testing.assert_allclose(y_out.cpu().numpy(), orc.last_targets[:, 0], rtol=REL, atol=1e-3)


def test_train_actor_and_soft_update_vs_reference(pkg, td3_golden):
    g = td3_golden
    agent, rb = make_agent(pkg, g, "w1"), fill_replay(pkg, g, seed=0)
    la = agent.train_actor(rb, idx=torch.from_numpy(g["idx_actor"]).cuda())
    np.testing.assert_allclose(la, float(g["actor_loss"]), rtol=REL)
    update_agrees(flat(agent, 0), g["w1_actor"], g["w2_actor"], g["w1_actor"])
    assert (flat(agent, 1) == g["w2_critic1"]).all()      # the actor loss must not touch critic 1 (robot.py:393-395)
    assert float(agent.grads.abs().max()) == 0.0          # gradients are re-zeroed by the optimiser pass
    agent.soft_update(agent.target_actor, agent.actor_network, agent.tau)
    agent.soft_update(agent.target_critic_network_1, agent.critic_network_1, agent.tau)
    agent.soft_update(agent.target_critic_network_2, agent.critic_network_2, agent.tau)
    for net, k in ((3, "t_actor"), (4, "t_critic1"), (5, "t_critic2")):
        np.testing.assert_allclose(flat(agent, net), g["w2_" + k], rtol=0, atol=1.2e-7)


def test_td3_update_six_epochs_vs_reference(pkg, td3_golden):
    g = td3_golden
    agent, rb = make_agent(pkg, g), fill_replay(pkg, g, seed=0)
    agent.train_critic(rb, noise=torch.from_numpy(g["noise_critic"]).cuda(), idx=torch.from_numpy(g["idx_critic"]).cuda())
    agent.train_actor(rb, idx=torch.from_numpy(g["idx_actor"]).cuda())
    for t, s in ((agent.target_actor, agent.actor_network), (agent.target_critic_network_1, agent.critic_network_1),
                 (agent.target_critic_network_2, agent.critic_network_2)):
        agent.soft_update(t, s, agent.tau)
    agent.num_epochs = 6
    closs, aloss = agent.td3_update(rb, noise=torch.from_numpy(g["upd_noise"]).cuda(), idx=torch.from_numpy(g["upd_idx"]).cuda())
    np.testing.assert_allclose(closs.cpu().numpy(), g["upd_critic_losses"], rtol=REL)
    np.testing.assert_allclose(aloss.cpu().numpy(), g["upd_actor_losses"], rtol=REL)
    for net, k in ((0, "actor"), (1, "critic1"), (2, "critic2")):
        w = flat(agent, net)
        np.testing.assert_allclose(w, g["w3_" + k], rtol=0, atol=2.5e-5)
        assert (np.abs(w - g["w3_" + k]) <= 2e-6).mean() > 0.99
    for net, k in ((3, "t_actor"), (4, "t_critic1"), (5, "t_critic2")):
        np.testing.assert_allclose(flat(agent, net), g["w3_" + k], rtol=0, atol=1e-6)
    assert agent.steps.cpu().tolist() == [4, 7]


@pytest.mark.parametrize("B,H,L", [(256, 256, 2), (100, 200, 3), (37, 64, 1), (2048, 256, 2), (5000, 128, 4)])
def test_update_vs_oracle_other_shapes(pkg, B, H, L):
    """Benchmark shapes (2x256, B=256 / large batch) and ragged ones against the numpy oracle: losses after one critic and one
    actor step, and the parameter updates."""
    rs = np.random.RandomState(B)
    mk = lambda i, o: to.kaiming_uniform_params(rs, i, H, L, o) + rs.normal(0, 0.01, to.param_count(i, H, L, o)).astype(np.float32)
    w = [mk(2, 2), mk(4, 1), mk(4, 1)]
    wt = [x + rs.normal(0, 0.003, x.shape).astype(np.float32) for x in w]
    agent = pkg.TD3(pkg.Residual_Actor_Network(H, L), pkg.Residual_Critic_Network(H, L), pkg.Residual_Critic_Network(H, L), batch_size=B)
    for net, x in enumerate(w + wt):
        agent.flat(net).copy_(torch.from_numpy(x).cuda())
    n = max(B, 6000)
    s = rs.uniform(0, 98.9999, (n, 2)).astype(np.float32)
    a = rs.uniform(-5, 5, (n, 2)).astype(np.float32)
    s2 = np.clip(s + a, 0, 98.9999).astype(np.float32)
    r = (-np.linalg.norm(s2 - np.array([80, 20], np.float32), axis=1)).astype(np.float32)
    d = (np.arange(n) % 50) == 49
    rb = pkg.ReplayBuffer(10000, seed=1)
    rb.push(torch.from_numpy(s).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(s2).cuda(), torch.from_numpy(d).cuda())
    idx_c = rs.permutation(n)[:B]
    idx_a = rs.permutation(n)[:B]
    noise = rs.normal(size=(B, 2)).astype(np.float32)
    orc = to.TD3Oracle(w[0], w[1], w[2], wt[0], wt[1], wt[2], hidden=H, layers=L)
    ol1, ol2 = orc.train_critic(s[idx_c], a[idx_c], r[idx_c], s2[idx_c], d[idx_c], noise)
    ola = orc.train_actor(s[idx_a])
    l1, l2 = agent.train_critic(rb, noise=torch.from_numpy(noise).cuda(), idx=torch.from_numpy(idx_c).cuda())
    la = agent.train_actor(rb, idx=torch.from_numpy(idx_a).cuda())
    np.testing.assert_allclose([l1, l2, la], [ol1, ol2, ola], rtol=REL)
    update_agrees(flat(agent, 1), w[1], orc.critic1, w[1], 0.99)
    update_agrees(flat(agent, 2), w[2], orc.critic2, w[2], 0.99)
    update_agrees(flat(agent, 0), w[0], orc.actor, w[0], 0.99)


def test_underfilled_replay_raises_like_the_reference(pkg):
    agent = pkg.TD3(pkg.Residual_Actor_Network(64, 1), pkg.Residual_Critic_Network(64, 1), pkg.Residual_Critic_Network(64, 1))
    rb = pkg.ReplayBuffer(100, seed=0)
    with pytest.raises(TypeError):
        agent.train_critic(rb)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 64, 100, 255, 256, 257, 511, 513, 1000, 4095, 4096, 4097, 10000, 20011])
def test_sample_indices_exact_for_many_sizes(pkg, n):
    """Power-of-two neighbourhoods exercise the mask-level crossings of random_interval; several consecutive draws exercise the
    hand-over of the stream position between samples and across MT19937 regenerations."""
    rb = pkg.ReplayBuffer(32768, seed=n)
    k = torch.arange(n, dtype=torch.float32, device="cuda")
    rb.push(torch.stack([k, k], 1), torch.stack([k, k], 1), k, torch.stack([k, k], 1), torch.zeros(n, dtype=torch.bool, device="cuda"))
    rs = np.random.RandomState(n)
    B = max(1, min(n, 97))
    got = rb.sample_indices(B, 7).cpu().numpy()
    exp = np.stack([rs.choice(n, B, replace=False) for _ in range(7)])
    assert (got == exp).all()
    got2 = rb.sample_indices(n, 2).cpu().numpy()          # full permutations, stream continues
    exp2 = np.stack([rs.choice(n, n, replace=False) for _ in range(2)])
    assert (got2 == exp2).all()


def test_pipelined_update_equals_unpipelined(pkg, td3_golden):
    """td3_update drawing its index sets in chunks on a side stream (overlapped with training) gives exactly the result of
    drawing all of them up front: same MT19937 stream, same epochs."""
    g = td3_golden
    outs = []
    for chunk in (0, 10):
        agent, rb = make_agent(pkg, g), fill_replay(pkg, g, seed=42)
        agent.num_epochs, agent.sample_chunk_epochs = 30, chunk
        noise = torch.randn((30, agent.batch_size, 2), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
        closs, aloss = agent.td3_update(rb, noise=noise)
        outs.append((closs.clone(), aloss.clone(), agent.params.clone(), rb.sample_indices(5, 1).clone()))
    # losses are summed over CTAs with float atomics (order varies run to run); parameters are deterministic
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-5, atol=0)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-5, atol=0)
    assert torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][3], outs[1][3])             # the RNG stream ends at the same position


def test_pipelined_update_numpy_global_stream(pkg, td3_golden):
    g = td3_golden
    agent, rb = make_agent(pkg, g), fill_replay(pkg, g, seed=None)     # unseeded: numpy's global stream, like the reference
    agent.num_epochs = 20
    np.random.seed(123)
    agent.td3_update(rb)
    after = np.random.uniform()
    rs = np.random.RandomState(123)
    for _ in range(30):                                                # 20 critic + 10 actor draws (robot.py:326, 382)
        rs.choice(len(rb), agent.batch_size, replace=False)
    assert after == rs.uniform()
