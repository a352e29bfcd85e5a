"""Tiny invocations of the main kernels for compute-sanitizer (memcheck / racecheck); development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtd3_b200 as rt
torch.manual_seed(0)
env = rt.Environment(num_envs=96, seed=3)
env.reset()
acts = torch.rand((40, 2, 96), device="cuda") * 15 - 7.5
env.rollout(acts)
env.step(acts[0].t())
env.reset(mask=torch.arange(96, device="cuda") % 3 == 0)
H = 128
agent = rt.TD3(rt.Residual_Actor_Network(H, 2), rt.Residual_Critic_Network(H, 2), rt.Residual_Critic_Network(H, 2), batch_size=64, num_epochs=2)
rb = rt.ReplayBuffer(500, seed=0)
s = torch.rand((300, 2), device="cuda") * 98
a = torch.rand((300, 2), device="cuda") * 10 - 5
rb.push(s, a, -s[:, 0], (s + a).clamp(0, 98.9), torch.zeros(300, dtype=torch.bool, device="cuda"))
agent.td3_update(rb, use_graph=False)                      # fp32 learner + exact sampler
agent.precision = "tf32"; agent.tc_min_batch = 1
agent.td3_update(rb, use_graph=False)                      # tcgen05 learner (one 64-row tile)
x = torch.rand((200, 2), device="cuda")
agent.actor_network(x)                                     # tcgen05 forward (2 tiles)
robot = rt.Robot(env.goal_state, hidden=64, layers=2, seed=1, buffer_size=2000)
robot.demo_grid_min_points = 1
robot.set_demonstration_states(np.random.RandomState(0).uniform(0, 99, (300, 2)))
robot._demo_flag.fill_(1)
st = env.robot_state.clone()
robot.process_transition(st, acts[1].t(), st, None)        # grid search path
torch.cuda.synchronize()
print("sanitize smoke ok")
