"""env.reset timing (development aid): all envs / random 10 % masks, per call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
n = 65536
env = rt.Environment(num_envs=n, seed=1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def t(fn):
    torch.cuda.synchronize(); ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize(); return ev[0].elapsed_time(ev[1]) * 1e3
print("pos after init: min %d max %d" % (int(env._bank.pos.min()), int(env._bank.pos.max())))
print("reset(all) calls 1..8 (us):", [round(t(lambda: env.reset())) for _ in range(8)])
ts = [t(lambda: env.reset()) for _ in range(200)]
print("reset(all) 200 calls: median %.1f us, max %.1f us, calls > 50 us: %d" % (sorted(ts)[100], max(ts), sum(x > 50 for x in ts)))
g = torch.Generator(device="cuda").manual_seed(0)
masks = [(torch.rand(n, device="cuda", generator=g) < 0.1) for _ in range(200)]
ts = [t(lambda m=m: env.reset(mask=m)) for m in masks]
print("reset(10%% mask) 200 calls: median %.1f us, max %.1f us, calls > 50 us: %d" % (sorted(ts)[100], max(ts), sum(x > 50 for x in ts)))
print("pos now: min %d max %d" % (int(env._bank.pos.min()), int(env._bank.pos.max())))
