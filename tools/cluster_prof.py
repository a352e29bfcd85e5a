"""Stage stamps (clock64 of CTA 0) of the cluster critic / actor kernels, warm L2, eager launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
from rtd3_b200 import _lib

def main(B, H, L):
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
    bc = torch.zeros(256, dtype=torch.int64, device="cuda"); ba = torch.zeros(256, dtype=torch.int64, device="cuda")
    for _ in range(5):
        ag._critic_step(rb, idx, noise, loss2); ag._actor_step(rb, idx, loss1)
    _lib.lib().rtd3_debug_cluster_prof(bc.data_ptr(), ba.data_ptr())
    ag._critic_step(rb, idx, noise, loss2); ag._actor_step(rb, idx, loss1)
    torch.cuda.synchronize()
    _lib.lib().rtd3_debug_cluster_prof(None, None)
    for name, b in (("critic", bc), ("actor", ba)):
        v = b.cpu().tolist(); k = v[0]; st = [(x >> 48, x & ((1 << 48) - 1)) for x in v[1:1 + k]]
        if k < 2:
            print(name, "no stamps"); continue
        print("B=%d H=%d L=%d %s: %d stamps, total %d cycles" % (B, H, L, name, k, st[-1][1] - st[0][1]))
        print("   code:delta ", " ".join("%d:%d" % (st[i + 1][0], st[i + 1][1] - st[i][1]) for i in range(k - 1)))
        tot = {}
        for i in range(k - 1):
            tot[st[i + 1][0]] = tot.get(st[i + 1][0], 0) + st[i + 1][1] - st[i][1]
        print("   per code   ", " ".join("%d:%d" % kv for kv in sorted(tot.items())))

if __name__ == "__main__":
    main(256, 256, 2)
    main(100, 200, 3)
