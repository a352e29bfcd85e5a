import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rtd3_b200 as rt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = rt.Environment(num_envs=n, seed=1)
robot = rt.Robot(env.goal_state, hidden=256, layers=2, seed=100, buffer_size=4 * n)
robot.episodes_per_update = 10 ** 9
robot.td3_agent.precision = sys.argv[2] if len(sys.argv) > 2 else "tf32"
robot.set_demonstration_states(np.random.RandomState(0).uniform(0, 99, (512, 2)))
fused = len(sys.argv) > 3 and sys.argv[3] in ("fused", "multi")
tr = rt.BatchedTrainer(env, robot, noise="philox" if fused else "randn", graph=False, fused=fused, check_interval=8)
tr.multi_tick_kernel = sys.argv[3] == "multi" if len(sys.argv) > 3 else False
ticks = int(sys.argv[4]) if len(sys.argv) > 4 else 12
if tr.multi_tick_kernel:
    tr.run(ticks)
else:
    for _ in range(ticks):
        tr.tick()
torch.cuda.synchronize()
print("ok")
