"""Per-call device times of the learner's pieces (eager launches, CUDA events, warm L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
from rtd3_b200 import _lib

def timeit(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

def main(B=256, H=256, L=2):
    torch.manual_seed(0)
    ag = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B)
    n = 10000
    rb = rt.ReplayBuffer(n, seed=0)
    s = torch.rand((n, 2), device="cuda") * 98; a = torch.rand((n, 2), device="cuda") * 10 - 5
    rb.push(s, a, -s[:, 0], (s + a).clamp(0, 98.9), torch.zeros(n, dtype=torch.bool, device="cuda"))
    idx = torch.randint(0, n, (B,), device="cuda", dtype=torch.int32)
    noise = torch.randn((B, 2), device="cuda")
    loss2 = torch.zeros(2, device="cuda"); loss1 = torch.zeros(1, device="cuda")
    x = torch.rand((B, 2), device="cuda")
    ag.sync_transposed()
    L_ = _lib.lib(); sp = _lib.stream_ptr()
    print("B=%d H=%d L=%d  cluster path %d, max active clusters %d" % (B, H, L, L_.rtd3_td3_cluster_supported(ag._handle, B), L_.rtd3_td3_cluster_occupancy(ag._handle, B)))
    print("  forward(actor)            %.1f us" % timeit(lambda: ag.forward(0, x)))
    print("  critic step (fwd/bwd+wgrad+adam) %.1f us" % timeit(lambda: ag._critic_step(rb, idx, noise, loss2)))
    print("  actor step (fwd/bwd+wgrad)       %.1f us" % timeit(lambda: ag._actor_step(rb, idx, loss1)))
    print("  adam+polyak               %.1f us" % timeit(lambda: ag._adam(0b111, 0b111)))
    print("  empty launch (sync_transposed) %.1f us" % timeit(lambda: ag.sync_transposed()))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for B in (16, 32, 64, 128, 256, 512):
            main(B, 256, 2)
        main(100, 200, 3)
        sys.exit(0)
    main()
    main(100, 200, 3)
    main(8192, 256, 2)
