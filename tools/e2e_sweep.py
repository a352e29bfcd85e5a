"""rollout_host chunk-count sweep + raw PCIe copy rates (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtd3_b200 as rt
N, T = 4096, 1000
env = rt.Environment(N, seed=0, maps=rt.synthetic_maps(0))
g = torch.Generator().manual_seed(0)
acts = (torch.rand((T, 2, N), generator=g) * 15 - 7.5).pin_memory()
out = torch.empty((T, 2, N)).pin_memory()
d = torch.empty((T, 2, N), device="cuda")
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = timed(lambda: d.copy_(acts, non_blocking=True)); print("H2D alone %.3f ms %.1f GB/s" % (t * 1e3, acts.numel() * 4 / t / 1e9))
t = timed(lambda: out.copy_(d, non_blocking=True)); print("D2H alone %.3f ms %.1f GB/s" % (t * 1e3, acts.numel() * 4 / t / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(acts, non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
t = timed(both); print("H2D || D2H %.3f ms %.1f GB/s each" % (t * 1e3, acts.numel() * 4 / t / 1e9))
for chunks in (2, 4, 8, 12, 16, 24, 32):
    t = timed(lambda: env.rollout_host(acts, out, chunks=chunks))
    print("chunks %2d: %.3f ms -> %.3e env-steps/s" % (chunks, t * 1e3, N * T / t))

# zero-copy: the rollout kernel reads the pinned host actions / writes the pinned host trajectory itself (UVA)
L = rt._lib.lib()
def zc():
    rt._lib.check(L.rtd3_env_rollout(env._handle, rt._lib.ptr(env._state[0]), rt._lib.ptr(env._state[1]), rt._lib.ptr(acts), rt._lib.ptr(out), N, T,
                                     rt._lib.stream_ptr(env.device)), "zc")
try:
    t = timed(zc)
    print("zero-copy both: %.3f ms -> %.3e env-steps/s" % (t * 1e3, N * T / t))
    def zc_in():
        rt._lib.check(L.rtd3_env_rollout(env._handle, rt._lib.ptr(env._state[0]), rt._lib.ptr(env._state[1]), rt._lib.ptr(acts), rt._lib.ptr(d), N, T,
                                         rt._lib.stream_ptr(env.device)), "zc")
    t = timed(zc_in); print("zero-copy in only: %.3f ms" % (t * 1e3))
    def zc_out():
        rt._lib.check(L.rtd3_env_rollout(env._handle, rt._lib.ptr(env._state[0]), rt._lib.ptr(env._state[1]), rt._lib.ptr(d), rt._lib.ptr(out), N, T,
                                         rt._lib.stream_ptr(env.device)), "zc")
    t = timed(zc_out); print("zero-copy out only: %.3f ms" % (t * 1e3))
    env2 = rt.Environment(N, seed=0, maps=rt.synthetic_maps(0)); env3 = rt.Environment(N, seed=0, maps=rt.synthetic_maps(0))
    rt._lib.check(L.rtd3_env_rollout(env2._handle, rt._lib.ptr(env2._state[0]), rt._lib.ptr(env2._state[1]), rt._lib.ptr(acts), rt._lib.ptr(out), N, T, rt._lib.stream_ptr(env.device)), "zc")
    ref = env3.rollout(acts.cuda())
    torch.cuda.synchronize()
    print("zero-copy result equals device rollout:", bool(torch.equal(out.cuda().permute(0, 2, 1), ref)))
except Exception as e:
    print("zero-copy failed:", e)

# hybrid: the kernel reads the pinned actions itself (zero-copy in), trajectory slices go back through the copy engine
s_out = torch.cuda.Stream()
def hybrid(chunks, reverse=False):
    main = torch.cuda.current_stream()
    s_out.wait_stream(main)
    bounds = [round(c * T / chunks) for c in range(chunks + 1)]
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if not reverse:
            rt._lib.check(L.rtd3_env_rollout(env._handle, rt._lib.ptr(env._state[0]), rt._lib.ptr(env._state[1]), rt._lib.ptr(acts[lo:hi]), rt._lib.ptr(d[lo:hi]), N, hi - lo,
                                             rt._lib.stream_ptr(env.device)), "hy")
            e = torch.cuda.Event(); e.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e)
                out[lo:hi].copy_(d[lo:hi], non_blocking=True)
        else:      # staged H2D, zero-copy out
            with torch.cuda.stream(s_out):
                d[lo:hi].copy_(acts[lo:hi], non_blocking=True)
                e = torch.cuda.Event(); e.record(s_out)
            main.wait_event(e)
            rt._lib.check(L.rtd3_env_rollout(env._handle, rt._lib.ptr(env._state[0]), rt._lib.ptr(env._state[1]), rt._lib.ptr(d[lo:hi]), rt._lib.ptr(out[lo:hi]), N, hi - lo,
                                             rt._lib.stream_ptr(env.device)), "hy")
    main.wait_stream(s_out)
for chunks in (4, 8, 16):
    t = timed(lambda: hybrid(chunks)); print("hybrid zero-copy in + D2H slices, %2d chunks: %.3f ms -> %.3e env-steps/s" % (chunks, t * 1e3, N * T / t))
    t = timed(lambda: hybrid(chunks, True)); print("hybrid H2D slices + zero-copy out, %2d chunks: %.3f ms -> %.3e env-steps/s" % (chunks, t * 1e3, N * T / t))
