"""The tcgen05 / TMEM (TF32) large-batch learner steps against the fp32 FFMA learner and the numpy oracle.

Same weights, same replay rows, same index sets, same smoothing noise.  Bars (north star: losses and Q-values within 1e-3
relative after one update): losses / Q-values / targets 1e-3 relative; gradients 1e-2 of each tensor's largest entry
(the worst of 65 536 entries; typical entries agree to ~1e-3)
(TF32 operands, round to nearest, fp32 accumulation; the batch reduction over CTAs is an L2 reduce-add, so the
summation order differs from the fp32 path).
"""
import numpy as np
import pytest
import torch

from oracle import td3_oracle as to

pytestmark = pytest.mark.gpu
REL = 1e-3


def make(pkg, H, B, n_rows=20000, seed=0):
    torch.manual_seed(seed)
    agent = pkg.TD3(pkg.Residual_Actor_Network(H, 2), pkg.Residual_Critic_Network(H, 2), pkg.Residual_Critic_Network(H, 2), batch_size=B)
    with torch.no_grad():
        agent.params.add_(0.01 * torch.randn_like(agent.params))           # non-zero biases, targets != online nets
    agent.sync_transposed()
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    rb = pkg.ReplayBuffer(n_rows, seed=0)
    s = torch.rand((n_rows, 2), device="cuda", generator=g) * 98.9
    a = torch.rand((n_rows, 2), device="cuda", generator=g) * 10 - 5
    s2 = (s + a).clamp(0, 98.9)
    goal = torch.tensor([80.0, 20.0], device="cuda")
    r = -(s2 - goal).norm(dim=1)
    done = (torch.arange(n_rows, device="cuda") % 50) == 49
    rb.push(s, a, r, s2, done)
    idx = torch.randint(0, n_rows, (B,), device="cuda", generator=g, dtype=torch.int32)
    noise = torch.randn((B, 2), device="cuda", generator=g)
    return agent, rb, idx, noise


def net_slices(agent, net):
    """(name, slice) of every parameter tensor of arena network `net` inside the flat gradient buffer."""
    H = agent.hidden
    in_dim, out_dim = (2, 2) if net == 0 else (4, 1)
    off = agent._off[net]
    out = []
    for name, n in (("W0", in_dim * H), ("b0", H), ("W1", H * H), ("b1", H), ("Wout", out_dim * H), ("bout", out_dim)):
        out.append((name, slice(off, off + n)))
        off += n
    return out


def assert_grads_close(agent, g_tc, g_ref, nets, tol=1e-2):
    for net in nets:
        for name, sl in net_slices(agent, net):
            a, b = g_tc[sl], g_ref[sl]
            scale = float(b.abs().max())
            err = float((a - b).abs().max())
            assert scale > 0, (net, name)
            assert err <= tol * scale, (net, name, err, scale)


def critic_pass(agent, rb, idx, noise):
    B = idx.numel()
    loss2 = torch.zeros(2, device="cuda")
    q = torch.zeros((2, B), device="cuda")
    y = torch.zeros((B,), device="cuda")
    agent.grads.zero_()
    agent._critic_step(rb, idx, noise, loss2, q, y, apply=False)
    torch.cuda.synchronize()
    g = agent.grads.clone()
    agent.grads.zero_()
    return loss2.cpu().numpy(), q.cpu().numpy(), y.cpu().numpy(), g


@pytest.mark.parametrize("H,B", [(256, 8192), (256, 2000), (128, 1041), (256, 64)])
def test_critic_step_tf32_matches_fp32_and_oracle(pkg, H, B):
    agent, rb, idx, noise = make(pkg, H, B)
    agent.tc_min_batch = 1
    l_ref, q_ref, y_ref, g_ref = critic_pass(agent, rb, idx, noise)
    agent.precision = "tf32"
    l_tc, q_tc, y_tc, g_tc = critic_pass(agent, rb, idx, noise)
    assert not np.array_equal(q_tc, q_ref)                                  # it really took the tensor-core path
    np.testing.assert_allclose(l_tc, l_ref, rtol=REL)
    np.testing.assert_allclose(y_tc, y_ref, rtol=REL, atol=REL * float(np.abs(y_ref).max()))
    np.testing.assert_allclose(q_tc, q_ref, rtol=REL, atol=REL * float(np.abs(q_ref).max()))
    assert_grads_close(agent, g_tc, g_ref, (1, 2), tol=1e-2 if B >= 1024 else 3e-2)
    assert float(g_tc[:agent._off[1]].abs().max()) == 0.0                   # the actor's gradient slot is untouched
    # the numpy oracle on the same rows (robot.py:312-366 restated), losses only: it is the fp32 path's own reference
    o = to.TD3Oracle(*(agent.flat(k).cpu().numpy() for k in range(6)), hidden=H, layers=2)
    s, a, r, s2, nd = (t.cpu().numpy() for t in rb.gather(idx))
    lo = o.train_critic(s, a, r, s2, nd < 0.5, noise.cpu().numpy())
    np.testing.assert_allclose(l_tc, lo, rtol=REL)
    np.testing.assert_allclose(y_tc, o.last_targets[:, 0], rtol=REL, atol=REL * float(np.abs(y_ref).max()))


def actor_pass(agent, rb, idx):
    loss1 = torch.zeros(1, device="cuda")
    agent.grads.zero_()
    agent._actor_step(rb, idx, loss1)
    torch.cuda.synchronize()
    g = agent.grads.clone()
    agent.grads.zero_()
    return float(loss1.cpu()[0]), g


@pytest.mark.parametrize("H,B", [(256, 8192), (256, 2000), (128, 1041), (256, 64)])
def test_actor_step_tf32_matches_fp32_and_oracle(pkg, H, B):
    agent, rb, idx, _ = make(pkg, H, B, seed=3)
    agent.tc_min_batch = 1
    l_ref, g_ref = actor_pass(agent, rb, idx)
    agent.precision = "tf32"
    l_tc, g_tc = actor_pass(agent, rb, idx)
    assert l_tc != l_ref
    with torch.no_grad():                                                    # the loss is a mean of signed Q-values: scale = mean |Q|
        s_rows = rb.gather(idx)[0]
        agent.precision = "fp32"
        q_scale = float(agent.critic_network_1(s_rows, agent.actor_network(s_rows)).abs().mean())
        agent.precision = "tf32"
    np.testing.assert_allclose(l_tc, l_ref, rtol=REL, atol=REL * q_scale)
    assert_grads_close(agent, g_tc, g_ref, (0,), tol=1e-2 if B >= 1024 else 3e-2)     # a single 64-row tile does not average
    assert float(g_tc[agent._off[1]:].abs().max()) == 0.0                   # critic gradients are discarded (robot.py:356)
    o = to.TD3Oracle(*(agent.flat(k).cpu().numpy() for k in range(6)), hidden=H, layers=2)
    s = rb.gather(idx)[0].cpu().numpy()
    np.testing.assert_allclose(l_tc, o.train_actor(s), rtol=REL, atol=REL * q_scale)


def test_td3_update_tf32_tracks_fp32(pkg):
    """Six epochs of robot.py:272-285 (critic steps, actor steps + Polyak on even epochs) in both precisions from the same
    state: every loss within 1e-3, parameters within a fraction of the accumulated Adam steps."""
    H, B, E = 256, 4096, 6
    agents = []
    for prec in ("fp32", "tf32"):
        agent, rb, _, _ = make(pkg, H, B)
        agent.num_epochs = E
        agent.precision = prec
        g = torch.Generator(device="cuda").manual_seed(7)
        idx = torch.randint(0, len(rb), (E + 3, B), device="cuda", generator=g, dtype=torch.int32)
        noise = torch.randn((E, B, 2), device="cuda", generator=g)
        closs, aloss = agent.td3_update(rb, noise=noise, idx=idx)
        agents.append((agent, closs.cpu().numpy().copy(), aloss.cpu().numpy().copy()))
    (a32, c32, l32), (atc, ctc, ltc) = agents
    assert int(atc.steps[1]) == E and int(atc.steps[0]) == 3
    np.testing.assert_allclose(ctc, c32, rtol=REL)
    np.testing.assert_allclose(ltc, l32, rtol=REL)
    d = (atc.params - a32.params).abs().max()
    assert float(d) <= E * 2e-5                                             # never more than the Adam steps taken (lr 1e-5) apart
    # the tensor-core operand copies followed the optimiser: a tf32 forward agrees with the fp32 forward of the SAME agent
    x = torch.rand((512, 2), device="cuda") * 50
    y_tc = atc.actor_network(x)
    atc.precision = "fp32"
    y_32 = atc.actor_network(x)
    assert float((y_tc - y_32).abs().max()) <= 4e-3 * float(y_32.abs().max())
