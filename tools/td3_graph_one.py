"""One td3_update (graph) at B=256 2x256 after warm-up, for an ncu per-node launch list (use --cache-control none)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
B, H, L, E = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torch.manual_seed(0)
ag = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B, num_epochs=E)
n = 10000
rb = rt.ReplayBuffer(n, seed=0)
s = torch.rand((n, 2), device="cuda") * 98; a = torch.rand((n, 2), device="cuda") * 10 - 5
rb.push(s, a, -s[:, 0], (s + a).clamp(0, 98.9), torch.zeros(n, dtype=torch.bool, device="cuda"))
idx = torch.randint(0, n, (E + (E + 1) // 2, B), device="cuda", dtype=torch.int32)
for _ in range(3):
    ag.td3_update(rb, idx=idx)
torch.cuda.synchronize()
