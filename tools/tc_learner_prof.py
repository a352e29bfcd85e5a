"""A few eager tf32 learner steps at B=8192, 2x256 for ncu (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import rtd3_b200 as pkg
from test_tc_learner_gpu import make

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
agent, rb, idx, noise = make(pkg, 256, B)
agent.precision = "tf32"
loss2 = torch.zeros(2, device="cuda"); loss1 = torch.zeros(1, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    ev[0].record()
    agent._critic_step(rb, idx, noise, loss2)
    ev[1].record()
    agent._actor_step(rb, idx, loss1)
    agent._adam(nets=0b001, polyak=0b111)
    ev[2].record()
    torch.cuda.synchronize()
    print("critic step + adam %.1f us, actor step + adam/polyak %.1f us" % (ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3))

# phase timeline of CTA 0 of the critic kernel (clock64 stamps, see LT_STAMP in rtd3_tc_learner.cu)
import ctypes
L = ctypes.CDLL(pkg._lib.LIB_PATH)
L.rtd3_debug_lt_prof(1, None)
agent._critic_step(rb, idx, noise, loss2)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
L.rtd3_debug_lt_prof(0, buf)
t = list(buf)
names = ["load_small", "layer0", "gemm handoff", "epilogue(wait+ld)", "out+rowlogic", "bwd_out", "dw(relayout+mma)", "drain", "gemm handoff", "epilogue dX", "colsum", "reduce_small"]
prev = t[0]
print("phase timeline (cycles):")
for p in range(5):
    n = 12 if p >= 3 else 5
    line = []
    for k in range(n):
        v = t[1 + p * 16 + k]
        line.append("%s %d" % (names[k], v - prev))
        prev = v
    print(" pass", p, "|", "; ".join(line))
print(" total", t[90] - t[0], "teardown", t[91] - t[90])
