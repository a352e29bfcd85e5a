"""Data-parallel learner check, run under torchrun with N ranks (one per GPU):
every rank holds the same weights and the same replay content, takes its 1/N slice of one global minibatch, and the flat
gradient buffer is all-reduced over NCCL before Adam.  Checks: (1) replicas stay bit-identical across ranks, (2) the
result matches a single-GPU step on the whole minibatch to summation-order round-off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import rtd3_b200 as rt

def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("RTD3_HANG_DUMP_S", "120")), exit=True)     # a hung rank prints where it is stuck
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, L, B = 256, 2, 256
    torch.manual_seed(0)
    def agent(pg, batch):
        torch.manual_seed(0 if pg is None else rank)      # ranks start from DIFFERENT weights: the constructor broadcasts rank 0's
        return rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=batch,
                      process_group=pg, num_epochs=6)
    n = 5000
    g = torch.Generator(device="cuda").manual_seed(1)
    s = torch.rand((n, 2), device="cuda", generator=g) * 98.9999
    a = torch.rand((n, 2), device="cuda", generator=g) * 10 - 5
    s2 = (s + a).clamp(0, 98.9999)
    r = -torch.linalg.norm(s2 - torch.tensor([80., 20.], device="cuda"), dim=1)
    d = (torch.arange(n, device="cuda") % 50) == 49
    rb = rt.ReplayBuffer(10000, seed=0)
    rb.push(s, a, r, s2, d)
    E = 6
    count = E + 3
    idx = torch.randint(0, n, (count, B), device="cuda", generator=g, dtype=torch.int32)
    noise = torch.randn((E, B, 2), device="cuda", generator=g)
    per = B // world
    dp = agent(dist.group.WORLD, per)
    closs, aloss = dp.td3_update(rb, noise=noise[:, rank * per:(rank + 1) * per].contiguous(), idx=idx[:, rank * per:(rank + 1) * per].contiguous())
    # losses are per-shard means: their average over ranks is the global-batch loss
    cl = closs.clone(); dist.all_reduce(cl); cl /= world
    al = aloss.clone(); dist.all_reduce(al); al /= world
    mine = dp.params.clone()
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool(torch.equal(lo, hi))
    ok = True
    if rank == 0:
        single = agent(None, B)
        c1, a1 = single.td3_update(rb, noise=noise, idx=idx)
        dl = float((cl - c1).abs().max() / c1.abs().max()); da = float((al - a1).abs().max() / a1.abs().max())
        dparam = float((single.params - mine).abs().max())
        frac = float(((single.params - mine).abs() <= 2e-6).float().mean())
        ok = identical and dl < 1e-3 and da < 1e-3 and dparam <= 6 * 2.05e-5 and frac > 0.98
        print("DP%d: replicas identical=%s  critic-loss rel diff %.2e  actor-loss rel diff %.2e  max |dparam| %.2e  frac equal %.4f -> %s"
              % (world, identical, dl, da, dparam, frac, "OK" if ok else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
