"""Why did the 8-tick-graph row slow down?  Counts captures / signature changes and times run() with pieces disabled (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rtd3_b200 as rt
from rtd3_b200 import trainer as T
n = 65536
caps = [0]
orig_capture = T._capture
def counting_capture(g):
    caps[0] += 1
    return orig_capture(g)
T._capture = counting_capture
def build():
    env = rt.Environment(num_envs=n, seed=1)
    robot = rt.Robot(env.goal_state, hidden=256, layers=2, seed=100, buffer_size=4 * n)
    robot.td3_agent.precision = "f16"; robot.td3_agent.batch_size = 256; robot.td3_agent.num_epochs = 20
    robot.memory.sampler = "philox"
    robot.set_demonstration_states(np.random.RandomState(0).uniform(0, 99, (512, 2)))
    return rt.BatchedTrainer(env, robot, noise="philox", graph=True, check_interval=8, fused=True), robot
for variant in ("current", "no_check", "no_prepare", "neither"):
    tr, robot = build()
    ag = robot.td3_agent
    sigs = []
    if variant in ("no_check", "neither"):
        tr._check_graphs = lambda: None
    else:
        oc = tr._check_graphs
        def logged():
            before = getattr(tr, "_graph_sig", None)
            oc()
            if tr._graph_sig != before: sigs.append(tr._graph_sig)
        tr._check_graphs = logged
    if variant in ("no_prepare", "neither"):
        real = ag.prepare_forward
        calls = [0]
        def once(b):
            if calls[0] == 0: real(b)
            calls[0] += 1
        ag.prepare_forward = once
    caps[0] = 0
    tr.run(160)
    while robot.num_updates < 1: tr.run(8)
    torch.cuda.synchronize(); c0 = caps[0]; u0 = robot.num_updates
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); tr.run(480); e1.record(); torch.cuda.synchronize()
    print("%-10s %.1f us/tick  captures in window %d (before %d)  updates %d  signature changes %d  t_stale=%s u_stale=%s h_stale=%s"
          % (variant, e0.elapsed_time(e1) / 480 * 1e3, caps[0] - c0, c0, robot.num_updates - u0, len(sigs), ag._t_stale, ag._u_stale, ag._h_stale), flush=True)
    for s in sigs[:6]: print("    sig", s)
    del tr, robot
