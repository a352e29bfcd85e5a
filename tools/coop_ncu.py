"""One short cooperative update (B=256, 2x256, E epochs, no graph) for an ncu capture of td3_update_coop_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtd3_b200 as rt

B, H, L, E = int(os.environ.get("B", 256)), 256, 2, int(os.environ.get("E", 2))
dev = torch.device("cuda", 0)
n = 10000
g = torch.Generator(device=dev).manual_seed(0)
s = torch.rand((n, 2), device=dev, generator=g) * 98.9999
a = torch.rand((n, 2), device=dev, generator=g) * 10 - 5
s2 = (s + a).clamp(0, 98.9999)
r = -torch.linalg.norm(s2 - torch.tensor([80., 20.], device=dev), dim=1)
rb = rt.ReplayBuffer(n, device=dev, seed=0)
rb.push(s, a, r, s2, (torch.arange(n, device=dev) % 50) == 49)
idx = torch.randint(0, n, (E + (E + 1) // 2, B), device=dev, dtype=torch.int32)
torch.manual_seed(0)
ag = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B, num_epochs=E)
ag.update_kernel = os.environ.get("KERNEL", "coop")
for _ in range(3):
    ag.td3_update(rb, idx=idx, use_graph=False)
torch.cuda.synchronize()
print("done")
