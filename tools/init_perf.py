import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
for n in (4096, 65536, 1 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    env = rt.Environment(num_envs=n, seed=1)
    torch.cuda.synchronize(); print("Environment(%d): %.1f ms" % (n, (time.perf_counter() - t0) * 1e3))
