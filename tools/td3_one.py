"""A few eager critic / actor steps (for an ncu launch list of the learner's kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt

def main(B, H, L, reps=3):
    torch.manual_seed(0)
    ag = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B)
    n = 10000
    rb = rt.ReplayBuffer(n, seed=0)
    s = torch.rand((n, 2), device="cuda") * 98; a = torch.rand((n, 2), device="cuda") * 10 - 5
    rb.push(s, a, -s[:, 0], (s + a).clamp(0, 98.9), torch.zeros(n, dtype=torch.bool, device="cuda"))
    idx = torch.randint(0, n, (B,), device="cuda", dtype=torch.int32)
    noise = torch.randn((B, 2), device="cuda")
    loss2 = torch.zeros(2, device="cuda"); loss1 = torch.zeros(1, device="cuda")
    ag.sync_transposed()
    for _ in range(reps):
        ag._critic_step(rb, idx, noise, loss2)
        ag._actor_step(rb, idx, loss1)
    torch.cuda.synchronize()

if __name__ == "__main__":
    for B in [int(x) for x in sys.argv[1].split(",")]:
        main(B, int(sys.argv[2]) if len(sys.argv) > 2 else 256, int(sys.argv[3]) if len(sys.argv) > 3 else 2)
