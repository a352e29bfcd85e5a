"""Quick timing of the TD3 update path on one GPU (development aid; bench.py carries the reported numbers)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtd3_b200 as rt

def run(B, H, L, epochs=100, reps=3, sampler=True, precision="fp32"):
    torch.manual_seed(0)
    agent = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B, num_epochs=epochs)
    agent.precision = precision
    n = 10000
    rb = rt.ReplayBuffer(n, seed=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    s = torch.rand((n, 2), device="cuda", generator=g) * 98.9999
    a = torch.rand((n, 2), device="cuda", generator=g) * 10 - 5
    s2 = (s + a).clamp(0, 98.9999)
    r = -torch.linalg.norm(s2 - torch.tensor([80., 20.], device="cuda"), dim=1)
    d = (torch.arange(n, device="cuda") % 50) == 49
    rb.push(s, a, r, s2, d)
    count = epochs + (epochs + 1) // 2
    idx = rb.sample_indices(min(B, n), count) if not sampler else None
    if B > n:
        idx = torch.randint(0, n, (count, B), device="cuda", dtype=torch.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(reps):
        torch.cuda.synchronize()
        rt._lib.launch_count_reset()
        ev[0].record()
        ii = idx if idx is not None else rb.sample_indices(B, count)
        ev[1].record()
        agent.td3_update(rb, idx=ii)
        ev[2].record()
        torch.cuda.synchronize()
        ts, tu = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(precision, "B=%d H=%d L=%d: sampler %.3f ms, update %.3f ms for %d epochs -> %.1f us/epoch, %.0f updates/s (update only), %.0f with sampler; launches %d"
          % (B, H, L, ts, tu, epochs, tu * 1e3 / epochs, epochs / (tu * 1e-3), epochs / ((tu + ts) * 1e-3), rt._lib.launch_count()))

if __name__ == "__main__":
    run(100, 200, 3)
    run(256, 256, 2)
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        for B, H, L in ((64, 256, 2), (128, 256, 2), (512, 256, 2), (1024, 256, 2), (256, 128, 2), (256, 256, 3)):
            run(B, H, L, sampler=False)
        sys.exit(0)
    run(8192, 256, 2, epochs=20)
    for B in (1024, 2048, 4096, 8192, 16384, 65536):
        run(B, 256, 2, epochs=20, sampler=False)
        run(B, 256, 2, epochs=20, sampler=False, precision="tf32")
