"""CPU restatement (numpy, float32) of the reference's replay buffer and residual-TD3 learner.  TEST INFRASTRUCTURE.

Follows /root/reference/robot.py:
  ReplayBuffer.push / sample / __len__          robot.py:58-124
  Residual_Actor_Network / _Critic_Network      robot.py:128-206   (2->H->H->H->2 and 4->H->H->H->1, ReLU, linear head)
  TD3.train_critic                              robot.py:312-366
  TD3.train_actor                               robot.py:369-398
  TD3.soft_update                               robot.py:293-310
  TD3.td3_update (epoch loop, delayed actor)    robot.py:258-285
torch (ATen linear / relu / MSELoss / autograd / optim.Adam) is the third-party holder of the arithmetic; its published
Adam update (torch/optim/adam.py, single-tensor form, no amsgrad / weight decay) is restated in `adam_step`.
Pinned against tests/golden/td3_golden.npz, which the unmodified reference produced with injected weights, captured
replay indices and captured target-policy noise (torch's RNG is never seeded by the reference).

Parameter vectors are flat float32 arrays in torch's `parameters()` order:
  [W1 (H x in) row-major, b1 (H), W2 (H x H), b2, ..., Wout (out x H), bout].
"""
import numpy as np

from .mt19937 import LegacyMT19937

F32 = np.float32


# ---------------------------------------------------------------------------------------------------- replay ring
class ReplayOracle:
    """robot.py:58-124 with SoA float32 rows like the device ring (the reference stores tuples and casts at robot.py:329-333)."""

    def __init__(self, capacity):
        self.capacity = capacity
        self.s = np.zeros((capacity, 2), F32)
        self.a = np.zeros((capacity, 2), F32)
        self.r = np.zeros((capacity,), F32)
        self.s2 = np.zeros((capacity, 2), F32)
        self.done = np.zeros((capacity,), bool)
        self.size = 0
        self.position = 0

    def push(self, state, action, reward, next_state, done):          # robot.py:79-96
        if self.size < self.capacity:
            self.size += 1
        p = self.position
        self.s[p], self.a[p], self.r[p], self.s2[p], self.done[p] = state, action, reward, next_state, done
        self.position = (self.position + 1) % self.capacity

    def __len__(self):
        return self.size

    def sample_indices(self, rng: LegacyMT19937, batch_size):         # robot.py:108-111
        if self.size < batch_size:
            return None
        return rng.choice_no_replace(self.size, batch_size)

    def gather(self, idx):                                             # robot.py:113-115 (+ casts of robot.py:329-333)
        return self.s[idx], self.a[idx], self.r[idx], self.s2[idx], self.done[idx]


# ---------------------------------------------------------------------------------------------------- MLPs
def layer_dims(in_dim, hidden, layers, out_dim):
    dims = [in_dim] + [hidden] * layers + [out_dim]
    return list(zip(dims[:-1], dims[1:]))


def param_count(in_dim, hidden, layers, out_dim):
    return sum(i * o + o for i, o in layer_dims(in_dim, hidden, layers, out_dim))


def unpack(flat, in_dim, hidden, layers, out_dim):
    """Views (W [out,in], b [out]) per layer into the flat vector."""
    out, off = [], 0
    for i, o in layer_dims(in_dim, hidden, layers, out_dim):
        W = flat[off:off + i * o].reshape(o, i)
        off += i * o
        b = flat[off:off + o]
        off += o
        out.append((W, b))
    return out


def mlp_forward(flat, x, in_dim, hidden, layers, out_dim):
    """robot.py:153-159 / 193-200.  Returns (y, cache of layer inputs and pre-activations)."""
    acts = [x.astype(F32)]
    ps = unpack(flat, in_dim, hidden, layers, out_dim)
    h = acts[0]
    for li, (W, b) in enumerate(ps):
        z = h @ W.T + b
        if li < len(ps) - 1:
            h = np.maximum(z, F32(0))
            acts.append(h)
        else:
            return z.astype(F32), acts


def mlp_backward(flat, acts, dy, in_dim, hidden, layers, out_dim):
    """Gradient of sum(dy * y) w.r.t. the flat parameters and the input (what autograd computes)."""
    ps = unpack(flat, in_dim, hidden, layers, out_dim)
    g = np.zeros_like(flat)
    gs = unpack(g, in_dim, hidden, layers, out_dim)
    d = dy.astype(F32)
    for li in range(len(ps) - 1, -1, -1):
        W, _ = ps[li]
        gW, gb = gs[li]
        gW[...] = d.T @ acts[li]
        gb[...] = d.sum(axis=0)
        d = d @ W
        if li > 0:
            d = d * (acts[li] > 0)
    return g, d


def actor_forward(flat, x, hidden=200, layers=3):
    return mlp_forward(flat, x, 2, hidden, layers, 2)


def critic_forward(flat, s, a, hidden=200, layers=3):
    return mlp_forward(flat, np.concatenate([s, a], axis=1), 4, hidden, layers, 1)     # robot.py:195


# ---------------------------------------------------------------------------------------------------- Adam / Polyak
class AdamState:
    def __init__(self, n):
        self.m = np.zeros(n, F32)
        self.v = np.zeros(n, F32)
        self.t = 0


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def adam_step(p, g, st: AdamState, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (robot.py:237-239), in place on p, in the operation order of torch's CPU kernels (pinned bit for bit to
    torch.optim.Adam by tests/test_oracle_td3.py): exp_avg.lerp_ is ONE fused multiply-add; exp_avg_sq.mul_(b2) is rounded, then
    addcmul_ adds ((1-b2)*g)*g with the last product fused into the sum; addcdiv_ adds (value*exp_avg)/denom."""
    st.t += 1
    st.m[...] = _fma32(g - st.m, F32(1 - b1), st.m)                   # exp_avg.lerp_(grad, 1-beta1)
    st.v[...] = _fma32(F32(1 - b2) * g, g, st.v * F32(b2))            # exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    bc1 = 1 - b1 ** st.t
    bc2 = 1 - b2 ** st.t
    step_size = lr / bc1
    denom = np.sqrt(st.v) / F32(np.sqrt(bc2)) + F32(eps)
    p[...] = p + (F32(-step_size) * st.m) / denom


def soft_update(target, source, tau=0.001):                           # robot.py:307-310
    target[...] = target * F32(1.0 - tau) + source * F32(tau)


# ---------------------------------------------------------------------------------------------------- TD3
class TD3Oracle:
    def __init__(self, actor, critic1, critic2, t_actor=None, t_critic1=None, t_critic2=None, hidden=200, layers=3,
                 actor_lr=1e-5, critic_lr=1e-5, gamma=0.99, tau=0.001, policy_noise=0.2, noise_clip=0.5,
                 policy_update_delay=2, max_action=5.0):
        self.H, self.L = hidden, layers
        cp = lambda a: np.array(a, dtype=F32, copy=True)
        self.actor, self.critic1, self.critic2 = cp(actor), cp(critic1), cp(critic2)
        self.t_actor = cp(actor if t_actor is None else t_actor)              # copy.deepcopy, robot.py:232-234
        self.t_critic1 = cp(critic1 if t_critic1 is None else t_critic1)
        self.t_critic2 = cp(critic2 if t_critic2 is None else t_critic2)
        self.opt_a, self.opt_c1, self.opt_c2 = AdamState(self.actor.size), AdamState(self.critic1.size), AdamState(self.critic2.size)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.gamma, self.tau = gamma, tau
        self.policy_noise, self.noise_clip, self.delay, self.max_action = policy_noise, noise_clip, policy_update_delay, max_action

    # robot.py:312-366 ; `noise` is the captured randn_like draw (unit normal, [B,2])
    def train_critic(self, s, a, r, s2, done, noise):
        s, a, s2 = s.astype(F32), a.astype(F32), s2.astype(F32)
        r = r.astype(F32).reshape(-1, 1)
        notdone = (1 - done.astype(F32)).reshape(-1, 1)                        # robot.py:333
        eps = np.clip(noise.astype(F32) * F32(self.policy_noise), -self.noise_clip, self.noise_clip).astype(F32)
        na, _ = actor_forward(self.t_actor, s2, self.H, self.L)
        na = np.clip(na + eps, -self.max_action, self.max_action).astype(F32)
        q1t, _ = critic_forward(self.t_critic1, s2, na, self.H, self.L)
        q2t, _ = critic_forward(self.t_critic2, s2, na, self.H, self.L)
        y = (r + F32(self.gamma) * np.minimum(q1t, q2t) * notdone).astype(F32)  # robot.py:345
        losses = []
        B = s.shape[0]
        for flat, opt in ((self.critic1, self.opt_c1), (self.critic2, self.opt_c2)):
            q, acts = critic_forward(flat, s, a, self.H, self.L)
            diff = q - y
            losses.append(float(np.mean(diff * diff, dtype=F32)))
            g, _ = mlp_backward(flat, acts, (F32(2.0 / B) * diff), 4, self.H, self.L, 1)
            adam_step(flat, g, opt, lr=self.critic_lr)
        self.last_targets = y
        return losses[0], losses[1]

    # robot.py:369-398
    def train_actor(self, s):
        s = s.astype(F32)
        B = s.shape[0]
        act, a_acts = actor_forward(self.actor, s, self.H, self.L)             # fed the raw replay state, robot.py:386
        q, c_acts = critic_forward(self.critic1, s, act, self.H, self.L)
        loss = float(-np.mean(q, dtype=F32))
        dq = np.full_like(q, F32(-1.0 / B))
        _, dx = mlp_backward(self.critic1, c_acts, dq, 4, self.H, self.L, 1)   # critic-1 parameter grads are discarded
        g, _ = mlp_backward(self.actor, a_acts, dx[:, 2:4], 2, self.H, self.L, 2)
        adam_step(self.actor, g, self.opt_a, lr=self.actor_lr)
        return loss

    def polyak_all(self):                                                      # robot.py:283-285
        soft_update(self.t_actor, self.actor, self.tau)
        soft_update(self.t_critic1, self.critic1, self.tau)
        soft_update(self.t_critic2, self.critic2, self.tau)

    # robot.py:258-285 with the sampled indices and noise supplied (in the order the reference consumes them)
    def td3_update(self, replay: ReplayOracle, idx_seq, noise_seq, num_epochs):
        it_idx, it_noise = iter(idx_seq), iter(noise_seq)
        c_losses, a_losses = [], []
        for epoch in range(num_epochs):
            s, a, r, s2, d = replay.gather(next(it_idx))
            c_losses.append(self.train_critic(s, a, r, s2, d, next(it_noise)))
            if epoch % self.delay == 0:
                s, _, _, _, _ = replay.gather(next(it_idx))
                a_losses.append(self.train_actor(s))
                self.polyak_all()
        return np.array(c_losses), np.array(a_losses)


def kaiming_uniform_params(rs: np.random.RandomState, in_dim, hidden, layers, out_dim):
    """Init of robot.py:161-165: W ~ U(+-sqrt(6/fan_in)), b = 0 (values differ from torch's RNG; distribution only)."""
    parts = []
    for i, o in layer_dims(in_dim, hidden, layers, out_dim):
        bound = np.sqrt(6.0 / i)
        parts.append(rs.uniform(-bound, bound, (o, i)).astype(F32).reshape(-1))
        parts.append(np.zeros(o, F32))
    return np.concatenate(parts)
