"""The tcgen05 / TMEM (TF32) large-batch forward against the fp32 FFMA path: same weights, same inputs, TF32 tolerance."""
import numpy as np
import pytest
import torch

from oracle import td3_oracle as to

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,L,B", [(256, 2, 128), (256, 2, 1000), (256, 3, 4096), (128, 2, 300), (64, 4, 129), (256, 2, 65536), (96, 2, 700), (160, 3, 257), (192, 2, 20000)])
def test_tf32_forward_matches_fp32_forward(pkg, H, L, B):
    torch.manual_seed(0)
    agent = pkg.TD3(pkg.Residual_Actor_Network(H, L), pkg.Residual_Critic_Network(H, L), pkg.Residual_Critic_Network(H, L))
    with torch.no_grad():
        agent.params.add_(0.01 * torch.randn_like(agent.params))       # non-zero biases
    agent.sync_transposed()
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.rand((B, 2), device="cuda", generator=g) * 100 - 50
    s = torch.rand((B, 2), device="cuda", generator=g) * 99
    a = torch.rand((B, 2), device="cuda", generator=g) * 10 - 5
    ref_a = agent.actor_network(xa)
    ref_q = agent.critic_network_1(s, a)
    ref_t = agent.target_critic_network_2(s, a)
    agent.precision = "tf32"
    out_a = agent.actor_network(xa)
    out_q = agent.critic_network_1(s, a)
    out_t = agent.target_critic_network_2(s, a)
    for out, ref in ((out_a, ref_a), (out_q, ref_q), (out_t, ref_t)):
        scale = float(ref.abs().max())
        err = float((out - ref).abs().max())
        assert err <= 4e-3 * (L - 1) * scale, (err, scale)   # TF32 operand truncation (2^-10 per product) compounds per tensor-core layer
        assert not torch.equal(out, ref)                               # it really took the tensor-core path
    # and against the numpy oracle (fp32) for a few rows
    w = agent.flat(0).cpu().numpy()
    ref_np, _ = to.actor_forward(w, xa[:64].cpu().numpy(), H, L)
    np.testing.assert_allclose(out_a[:64].cpu().numpy(), ref_np, rtol=0, atol=4e-3 * (L - 1) * float(np.abs(ref_np).max()))


def test_tf32_weights_follow_the_optimiser(pkg):
    """After an optimiser step the chunk-major copy is rebuilt before the next tensor-core forward."""
    H, L, B = 256, 2, 256
    torch.manual_seed(0)
    agent = pkg.TD3(pkg.Residual_Actor_Network(H, L), pkg.Residual_Critic_Network(H, L), pkg.Residual_Critic_Network(H, L), batch_size=B)
    agent.precision = "tf32"
    n = 2000
    rb = pkg.ReplayBuffer(4000, seed=0)
    s = torch.rand((n, 2), device="cuda") * 98
    a = torch.rand((n, 2), device="cuda") * 10 - 5
    rb.push(s, a, -s[:, 0], (s + a).clamp(0, 98.9), torch.zeros(n, dtype=torch.bool, device="cuda"))
    x = torch.rand((512, 2), device="cuda") * 50
    y0 = agent.actor_network(x).clone()
    agent.actor_lr = 1e-2                                              # a visible step
    agent.train_critic(rb)
    agent.train_actor(rb)
    y1 = agent.actor_network(x)
    agent.precision = "fp32"
    y1_ref = agent.actor_network(x)
    assert float((y1 - y0).abs().max()) > 1e-3
    assert float((y1 - y1_ref).abs().max()) <= 4e-3 * float(y1_ref.abs().max())


@pytest.mark.parametrize("H,B", [(256, 128), (256, 1000), (256, 65536), (128, 300), (64, 129), (96, 700), (160, 257), (192, 20000), (224, 5000), (256, 8192)])
def test_f16_resident_weight_forward_matches_fp32_forward(pkg, H, B):
    """precision = "f16": fp16 operands (the 11-bit significand of TF32), hidden weight resident in shared memory, two accumulator
    buffers in TMEM (tile i's epilogue under tile i+1's products).  All six networks, ragged last tiles, 1..4 tiles per CTA."""
    L = 2
    torch.manual_seed(0)
    agent = pkg.TD3(pkg.Residual_Actor_Network(H, L), pkg.Residual_Critic_Network(H, L), pkg.Residual_Critic_Network(H, L))
    with torch.no_grad():
        agent.params.add_(0.01 * torch.randn_like(agent.params))       # non-zero biases, targets differ from the online nets
    agent.sync_transposed()
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.rand((B, 2), device="cuda", generator=g) * 200 - 100     # |state - goal| up to the world size
    s = torch.rand((B, 2), device="cuda", generator=g) * 99
    a = torch.rand((B, 2), device="cuda", generator=g) * 10 - 5
    nets = (lambda: agent.actor_network(xa), lambda: agent.critic_network_1(s, a), lambda: agent.critic_network_2(s, a),
            lambda: agent.target_actor(xa), lambda: agent.target_critic_network_1(s, a), lambda: agent.target_critic_network_2(s, a))
    refs = [f().clone() for f in nets]
    agent.precision = "f16"
    outs = [f().clone() for f in nets]
    for out, ref in zip(outs, refs):
        scale = float(ref.abs().max())
        err = float((out - ref).abs().max())
        assert err <= 4e-3 * scale, (err, scale)
        assert not torch.equal(out, ref)                               # it really took the tensor-core path
    assert not torch.equal(outs[1], outs[2])                           # the two critics read their own weights
    w = agent.flat(0).cpu().numpy()
    ref_np, _ = to.actor_forward(w, xa[:64].cpu().numpy(), H, L)
    np.testing.assert_allclose(outs[0][:64].cpu().numpy(), ref_np, rtol=0, atol=4e-3 * float(np.abs(ref_np).max()))


def test_f16_weights_follow_the_optimiser_and_the_update_graph(pkg):
    """The fp16 copies are rebuilt before td3_update returns (a captured tick graph cannot do that bookkeeping itself)."""
    H, L, B = 256, 2, 256
    torch.manual_seed(0)
    agent = pkg.TD3(pkg.Residual_Actor_Network(H, L), pkg.Residual_Critic_Network(H, L), pkg.Residual_Critic_Network(H, L), batch_size=B)
    agent.precision = "f16"
    agent.num_epochs = 4
    n = 2000
    rb = pkg.ReplayBuffer(4000, seed=0)
    s = torch.rand((n, 2), device="cuda") * 98
    a = torch.rand((n, 2), device="cuda") * 10 - 5
    rb.push(s, a, -s[:, 0], (s + a).clamp(0, 98.9), torch.zeros(n, dtype=torch.bool, device="cuda"))
    x = torch.rand((512, 2), device="cuda") * 50
    y0 = agent.actor_network(x).clone()
    agent.actor_lr = 1e-2                                              # a visible step
    agent.td3_update(rb)
    assert not agent._h_stale                                          # current without another forward() call
    out = torch.empty((512, 2), device="cuda")
    pkg._lib.check(pkg._lib.lib().rtd3_mlp_forward_f16(H, L, 0, pkg._lib.ptr(agent.params), pkg._lib.ptr(agent.params_h), pkg._lib.ptr(x),
                                                       pkg._lib.ptr(out), 512, pkg._lib.stream_ptr(agent.device)))
    agent.precision = "fp32"
    y1_ref = agent.actor_network(x)
    assert float((out - y0).abs().max()) > 1e-3
    assert float((out - y1_ref).abs().max()) <= 4e-3 * float(y1_ref.abs().max())
