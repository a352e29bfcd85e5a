"""Full-loop tick: one launch per hook vs the fused tick (rtd3_tick_pre / forward / rtd3_tick_post), single-tick graph vs the
check_interval-tick graph of BatchedTrainer.run (development aid).  usage: tick_perf.py [envs ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rtd3_b200 as rt

def build(n, fused, noise, precision, updates=False):
    env = rt.Environment(num_envs=n, seed=1)
    robot = rt.Robot(env.goal_state, hidden=256, layers=2, seed=100, buffer_size=max(50000, 4 * n))
    robot.td3_agent.precision = precision; robot.td3_agent.batch_size = 256; robot.td3_agent.num_epochs = 20
    robot.memory.sampler = "philox"
    if not updates:
        robot.episodes_per_update = 10 ** 9
    m = int(os.environ.get("DEMOS", "512"))
    if m == 512:
        robot.set_demonstration_states(np.random.RandomState(0).uniform(0, 99, (512, 2)))
    else:                                                   # reference-like: three noisy paths across the world
        rs = np.random.RandomState(0)
        t = np.linspace(0, 1, m // 3 + 1)[:, None]
        robot.set_demonstration_states(np.concatenate([rs.uniform(5, 95, (1, 2)) * (1 - t) + rs.uniform(5, 95, (1, 2)) * t + rs.normal(0, 2.5, (t.shape[0], 2))
                                                       for _ in range(3)])[:m])
    return rt.BatchedTrainer(env, robot, noise=noise, graph=True, check_interval=8, fused=fused), robot

def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    fn(reps)
    e1.record(); t_issue = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, t_issue / reps * 1e6

for n in [int(a) for a in sys.argv[1:]] or [8192, 65536]:
    for precision in os.environ.get("PRECISIONS", "tf32,fp32").split(","):
        for label, fused, noise, multi in (("hook-by-hook, 1-tick graph", False, "randn", False), ("fused, 1-tick graph", True, "philox", False),
                                          ("fused, 8-tick graph", True, "philox", True), ("fused, 8-tick kernel", True, "philox", "kernel")):
            if multi == "kernel" and precision != "f16":
                continue
            for updates in (False, True):
                tr, robot = build(n, fused, noise, precision, updates)
                tr.multi_tick_kernel = multi == "kernel"
                loop = (lambda k: tr.run(k)) if multi else (lambda k: [tr.tick() for _ in range(k)])
                loop(160 if updates else 32)
                u0 = robot.num_updates
                dev, host = timed(loop, 480)
                print("n=%6d %s %-28s updates=%d: %7.1f us/tick (host issue %6.1f us/tick)  %.3e env-steps/s  [%d updates]"
                      % (n, precision, label, updates, dev, host, n / dev * 1e6, robot.num_updates - u0), flush=True)
                del tr, robot
