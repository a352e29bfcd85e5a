"""Exact index draws (150 x B of n) a few times: timing and an ncu target."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, rtd3_b200 as rt
n, B, count = 10000, 256, 150
rb = rt.ReplayBuffer(n, seed=0)
z = torch.zeros((n, 2), device="cuda")
rb.push(z, z, z[:, 0], z, torch.zeros(n, dtype=torch.bool, device="cuda"))
ref = np.random.RandomState(0)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx = rb.sample_indices(B, count)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    want = np.stack([ref.choice(n, B, replace=False) for _ in range(count)])
    print("rep %d: %.3f ms, bit-exact %s" % (rep, (t1 - t0) * 1e3, bool((idx.cpu().numpy() == want).all())))
