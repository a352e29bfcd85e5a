"""Data-parallel learner + full loop under torchrun (one rank per GPU): timing and replica identity for both collectives.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dp_loop_check.py [nccl|p2p]

Every rank seeds torch DIFFERENTLY (the constructor's broadcast has to make the replicas equal), holds its own replay shard and env
shard, and runs (1) 100-epoch td3_update calls at B = 256 per rank, (2) the fused full loop at 8 192 envs per rank with the
asynchronous update check.  Prints one JSON line per measurement on rank 0; exits non-zero when the replicas differ."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import rtd3_b200 as rt
from rtd3_b200.trainer import replicas_identical


def main():
    import faulthandler
    arm = lambda: faulthandler.dump_traceback_later(int(os.environ.get("RTD3_HANG_DUMP_S", "150")), exit=True)   # a hung rank prints where it is stuck
    arm()
    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    envs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
    torch.manual_seed(1000 + rank)
    ok = True

    # ---- (1) DP epoch time, B = 256 per rank, 2 x 256
    arm()
    H, L, B, E = 256, 2, 256, 100
    agent = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L), batch_size=B,
                   num_epochs=E, device=dev, process_group=pg, dp_collective=mode)
    same0 = replicas_identical(agent, pg)
    n = 10000
    g = torch.Generator(device=dev).manual_seed(rank)
    s = torch.rand((n, 2), device=dev, generator=g) * 98.9999
    a = torch.rand((n, 2), device=dev, generator=g) * 10 - 5
    s2 = (s + a).clamp(0, 98.9999)
    r = -torch.linalg.norm(s2 - torch.tensor([80., 20.], device=dev), dim=1)
    rb = rt.ReplayBuffer(n, device=dev, seed=rank)
    rb.push(s, a, r, s2, (torch.arange(n, device=dev) % 50) == 49)
    rb.sampler = "philox"
    ms = []
    for rep in range(5):
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        agent.td3_update(rb)
        e1.record()
        torch.cuda.synchronize(dev)
        if rep:
            ms.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ms))], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    same1 = replicas_identical(agent, pg)
    ok = ok and same0 and same1
    if rank == 0:
        print(json.dumps({"what": "dp_epoch", "collective": mode, "world": world, "batch_per_rank": B, "us_per_epoch": float(t[0]) * 1e3 / E,
                          "replicas_identical_after_init": same0, "replicas_identical": same1}), flush=True)
    del agent, rb

    # ---- (2) full loop, fused tick, f16 forward, async update check
    rs = np.random.RandomState(0)
    tt = np.linspace(0, 1, 3785)[:, None]
    demos = np.concatenate([rs.uniform(5, 95, (1, 2)) * (1 - tt) + rs.uniform(5, 95, (1, 2)) * tt + rs.normal(0, 2.5, (3785, 2)) for _ in range(3)])
    for async_check in (False, True):
        arm()
        env = rt.Environment(num_envs=envs, seed=1707366464 + rank * envs, device=dev)
        robot = rt.Robot(env.goal_state, hidden=256, layers=2, seed=100 + rank, device=dev, process_group=pg, buffer_size=max(50000, 8 * envs),
                         dp_collective=mode)
        robot.td3_agent.precision = "f16"
        robot.td3_agent.batch_size = 256
        robot.td3_agent.num_epochs = 20
        robot.memory.sampler = "philox"
        robot.set_demonstration_states(demos)
        tr = rt.BatchedTrainer(env, robot, noise="philox", graph=True, check_interval=8, fused=True, async_check=async_check)
        warm = 0
        while warm < 16 or (robot.num_updates < 1 and warm < 400):
            tr.run(8)
            warm += 8
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ticks, upd0 = 480, robot.num_updates
        e0.record()
        tr.run(ticks)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same = replicas_identical(robot.td3_agent, pg)
        ok = ok and same
        if rank == 0:
            print(json.dumps({"what": "full_loop", "collective": mode, "world": world, "envs_per_gpu": envs, "async_check": async_check,
                              "us_per_tick": float(t[0]) * 1e3 / ticks, "updates_in_window": robot.num_updates - upd0,
                              "multi_tick_kernel": bool(tr._multi_tick_ok()), "replicas_identical": same}), flush=True)
        del tr, robot, env
        import gc
        gc.collect()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
