"""The bench's rollout launch (4096 envs x 1000 steps) a few times, for ncu (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
n, T = 4096, 1000
env = rt.Environment(num_envs=n, seed=1707366464, maps=rt.synthetic_maps(0))
env.reset()
acts = [torch.rand((T, 2, n), device="cuda") * 15 - 7.5 for _ in range(6)]
for k in range(8):
    env.rollout(acts[k % 6])
torch.cuda.synchronize()
print("ok")
