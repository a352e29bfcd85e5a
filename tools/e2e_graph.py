"""rollout_host: the graph-replayed copy-engine pipeline (rtd3_env_rollout_host) against the other modes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtd3_b200 as rt
N, T = 4096, 1000
env = rt.Environment(N, seed=0, maps=rt.synthetic_maps(0))
g = torch.Generator().manual_seed(0)
acts = [(torch.rand((T, 2, N), generator=g) * 15 - 7.5).pin_memory() for _ in range(2)]
outs = [torch.empty((T, 2, N)).pin_memory() for _ in range(2)]
def timed(fn, reps=20):
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps): fn(k)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
d = torch.empty((T, 2, N), device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both(k):
    with torch.cuda.stream(s1): d.copy_(acts[k % 2], non_blocking=True)
    with torch.cuda.stream(s2): outs[k % 2].copy_(d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    torch.cuda.synchronize()
t = timed(both); print("raw H2D || D2H %.3f ms" % (t * 1e3))
for mode, chunk_list in (("zero_copy", (None,)), ("staged", (8,)), ("graph", (4, 8, 12, 16, 24, 32, 64)), ("graph_in", (8, 16, 32)), ("graph_out", (8, 16, 32))):
    for ch in chunk_list:
        t = timed(lambda k: env.rollout_host(acts[k % 2], outs[k % 2], chunks=ch, mode=mode))
        print("%-9s chunks %4s: %.3f ms -> %.3e env-steps/s" % (mode, ch, t * 1e3, N * T / t), flush=True)
ref = rt.Environment(N, seed=0, maps=rt.synthetic_maps(0))
e2 = rt.Environment(N, seed=0, maps=rt.synthetic_maps(0))
o = e2.rollout_host(acts[0], outs[0])
print("graph result equals device rollout:", bool(torch.equal(o.cuda().permute(0, 2, 1), ref.rollout(acts[0].cuda()))))
