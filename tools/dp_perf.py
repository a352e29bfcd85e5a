"""Data-parallel TD3 epoch time under torchrun (development aid): B rows per rank, update only (index sets given), replicas checked.
RTD3_P2P_FUSE selects the form of the peer-memory step (2 inside the weight-gradient kernels, 1 all-reduce + optimiser kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import rtd3_b200 as rt

def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, E = 10000, 100
    g = torch.Generator(device="cuda").manual_seed(1 + rank)
    s = torch.rand((n, 2), device="cuda", generator=g) * 98.9999
    a = torch.rand((n, 2), device="cuda", generator=g) * 10 - 5
    s2 = (s + a).clamp(0, 98.9999)
    r = -torch.linalg.norm(s2 - torch.tensor([80., 20.], device="cuda"), dim=1)
    d = (torch.arange(n, device="cuda") % 50) == 49
    rb = rt.ReplayBuffer(n, seed=rank)
    rb.push(s, a, r, s2, d)
    for B, H, L in ((256, 256, 2), (100, 200, 3)):
        torch.manual_seed(rank)
        ag = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B,
                    process_group=dist.group.WORLD, num_epochs=E)
        idx = torch.randint(0, n, (E + E // 2, B), device="cuda", generator=g, dtype=torch.int32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(6):
            dist.barrier(); torch.cuda.synchronize()
            rt._lib.launch_count_reset()
            e0.record()
            ag.td3_update(rb, idx=idx)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep >= 2: best = min(best, float(t))
        lo, hi = ag.params.clone(), ag.params.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        fin = bool(torch.isfinite(ag.params).all())
        if rank == 0:
            print("DP%d fuse=%s B=%d/rank %dx%d: %.1f us/epoch (launches %d per update), replicas identical=%s finite=%s"
                  % (world, os.environ.get("RTD3_P2P_FUSE", "2"), B, L, H, best * 1e3 / E, rt._lib.launch_count(), bool(torch.equal(lo, hi)), fin), flush=True)
        if os.environ.get("RTD3_P2P_PROF") == "1" and rank == 0:
            import numpy as np
            buf = np.zeros((256, 2, 8), dtype=np.uint64)
            if rt._lib.lib().rtd3_debug_p2p_prof(buf.ctypes.data):
                t = buf.astype(np.int64)
                live = t[:, 0, 0] > 0
                t0 = t[live, :, 0].min()
                print("  stamps of the last fused weight-gradient launch (ns after the first block's start; blocks with stamps: %d)" % live.sum())
                for name, k in (("start", 0), ("reduced", 1), ("pushed", 2), ("summed", 5), ("end", 6)):
                    for th in (0, 1):
                        v = t[live, th, k]; v = v[v > 0] - t0
                        if v.size: print("    %-16s thread %d: min %6d  median %6d  max %6d  (n=%d)" % (name, th, v.min(), np.median(v), v.max(), v.size))
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
