"""process_transition with the reference's 11 355 demonstration states: full sweep vs the candidate lists of rtd3_demo_lists (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtd3_b200 as pkg
def run(n, m, min_points, name, near=False):
    rs = np.random.RandomState(0)
    t = np.linspace(0, 1, max(m, 3) // 3 + 1)[:, None]
    demos = np.concatenate([np.array([[5.0, 80.0]]) * (1 - t) + np.array([[90.0, 15.0]]) * t + rs.normal(0, s, (t.shape[0], 2)) for s in (0.5, 2.5, 6.0)])[:m]
    goals = rs.uniform(5, 95, (n, 2))
    if near:
        nxt = torch.from_numpy((demos[rs.randint(0, max(m, 1), n)] + rs.normal(0, 1.5, (n, 2))).clip(0, 98.9).astype(np.float32)).cuda()
    else:
        nxt = torch.rand((n, 2), device="cuda") * 98.9
    act = torch.zeros((n, 2), device="cuda")
    robot = pkg.Robot(torch.from_numpy(goals).cuda(), seed=5, buffer_size=200000)
    robot.demo_grid_min_points = min_points
    if m:
        robot.set_demonstration_states(demos)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        robot.set_demonstration_states(demos)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        if robot._demo_cells is not None:
            sz = (robot._demo_cells[1:] - robot._demo_cells[:-1]).float()
            print("           lists: built in %.2f ms, %d entries, mean %.1f max %d per cell" % (build_ms, robot._demo_list.shape[0], sz.mean().item(), int(sz.max())))
    robot._demo_flag.fill_(1)
    for _ in range(3):
        robot.process_transition(nxt, act, nxt, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        robot.process_transition(nxt, act, nxt, None)
    e1.record(); torch.cuda.synchronize()
    print("%-10s n=%6d m=%6d %s: %.1f us per process_transition" % (name, n, m, "near demos" if near else "uniform   ", e0.elapsed_time(e1) * 1e3 / 20))
for n in (4096, 65536):
    run(n, 0, 1, "no demos")
    for m in (64, 1000, 11355):
        run(n, m, 10 ** 9, "full sweep")
        run(n, m, 1, "lists")
    run(n, 11355, 1, "lists", near=True)
