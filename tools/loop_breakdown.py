"""Where a full-loop tick's time goes at 65 536 envs (development aid): graph replay alone vs tick() vs tick() with updates."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rtd3_b200 as rt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
def build(updates):
    env = rt.Environment(num_envs=n, seed=1)
    robot = rt.Robot(env.goal_state, hidden=256, layers=2, seed=100, buffer_size=4 * n)
    robot.td3_agent.precision = "tf32"; robot.td3_agent.batch_size = 256; robot.td3_agent.num_epochs = 20
    robot.memory.sampler = "philox"
    if not updates:
        robot.episodes_per_update = 10 ** 9
    robot.set_demonstration_states(np.random.RandomState(0).uniform(0, 99, (512, 2)))
    tr = rt.BatchedTrainer(env, robot, noise="randn", graph=True, check_interval=8)
    for _ in range(16):
        tr.tick()
    torch.cuda.synchronize()
    return tr, robot
def timed(fn, reps=240):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); t_issue = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, t_issue / reps * 1e6
tr, robot = build(False)
print("graph replay only      : %.1f us/tick on the GPU, %.1f us/tick of host issue time" % timed(lambda: tr._graph.replay()))
print("tick(), no updates     : %.1f us/tick, host %.1f us/tick" % timed(tr.tick))
tr, robot = build(True)
while robot.num_updates < 1:
    tr.tick()
u0 = robot.num_updates
print("tick(), with updates   : %.1f us/tick, host %.1f us/tick" % timed(tr.tick), "updates", robot.num_updates - u0)
ag = robot.td3_agent
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ag.td3_update(robot.memory)
    torch.cuda.synchronize(); print("td3_update(memory) call: %.2f ms (E=%d, B=%d, rows %d, sampler %s)" % ((time.perf_counter() - t0) * 1e3, ag.num_epochs, ag.batch_size, len(robot.memory), robot.memory.sampler))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): ag.td3_update(robot.memory)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
