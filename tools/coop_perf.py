"""us per epoch of td3_update: per-step kernels vs the persistent cooperative kernel (update only, index sets given)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtd3_b200 as rt

def main():
    dev = torch.device("cuda", 0)
    for (B, H, L, E) in ((256, 256, 2, 100), (100, 200, 3, 100), (512, 256, 2, 100), (64, 256, 2, 100), (2048, 256, 2, 20)):
        n = 10000
        g = torch.Generator(device=dev).manual_seed(0)
        s = torch.rand((n, 2), device=dev, generator=g) * 98.9999
        a = torch.rand((n, 2), device=dev, generator=g) * 10 - 5
        s2 = (s + a).clamp(0, 98.9999)
        r = -torch.linalg.norm(s2 - torch.tensor([80., 20.], device=dev), dim=1)
        rb = rt.ReplayBuffer(n, device=dev, seed=0)
        rb.push(s, a, r, s2, (torch.arange(n, device=dev) % 50) == 49)
        idx = torch.randint(0, n, (E + (E + 1) // 2, B), device=dev, dtype=torch.int32)
        for kernel in ("steps", "coop"):
            torch.manual_seed(0)
            ag = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B, num_epochs=E)
            ag.update_kernel = kernel
            for _ in range(2):
                ag.td3_update(rb, idx=idx)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                ag.td3_update(rb, idx=idx)
            e1.record()
            torch.cuda.synchronize()
            print("B=%d %dx%d %s: %.1f us/epoch" % (B, L, H, kernel, e0.elapsed_time(e1) * 1e3 / (reps * E)), flush=True)
            if kernel == "coop" and os.environ.get("RTD3_COOP_PROF"):
                buf = torch.zeros(256, dtype=torch.int64, device=dev)
                rt._lib.lib().rtd3_debug_coop_prof(rt._lib.ptr(buf))
                ag.td3_update(rb, idx=idx, use_graph=False)
                torch.cuda.synchronize()
                rt._lib.lib().rtd3_debug_coop_prof(None)
                full = buf.cpu().numpy()
                ts = full[128:]
                ts = ts[ts > 0]
                # per tile of block 0: entry, [panel: start, B issued, A built, data landed], compute done, epilogue done
                print("   block 0 tile phases (us since first stamp):", " ".join("%.1f" % ((v - ts[0]) / 1e3) for v in ts[:42]))
                st = full[:128]
                st = st[st > 0]
                # stamps alternate: arrival of block 0 at barrier k, release from barrier k
                work = [(st[i] - (st[i - 1] if i else st[0])) / 1e3 for i in range(0, len(st), 2)]
                wait = [(st[i + 1] - st[i]) / 1e3 for i in range(0, len(st) - 1, 2)]
                print("   last epoch, block 0: work before each barrier (us):", " ".join("%.1f" % w for w in work))
                print("   barrier wait (us):", " ".join("%.1f" % w for w in wait), " total %.1f" % ((st[-1] - st[0]) / 1e3), flush=True)

if __name__ == "__main__":
    main()
