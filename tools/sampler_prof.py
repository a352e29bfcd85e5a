import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
n = 10000
rb = rt.ReplayBuffer(n, seed=0)
k = torch.arange(n, dtype=torch.float32, device="cuda")
rb.push(torch.stack([k, k], 1), torch.stack([k, k], 1), k, torch.stack([k, k], 1), torch.zeros(n, dtype=torch.bool, device="cuda"))
for B, count in ((256, 150), (8192, 30)):
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rb.sample_indices(B, count); e1.record(); torch.cuda.synchronize()
    print(B, count, "%.3f ms" % e0.elapsed_time(e1))
