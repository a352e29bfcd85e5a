"""fp32 FFMA forward vs tcgen05 TF32 (streamed weights) and f16 (resident weights) forwards, device time per call (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtd3_b200 as rt
from rtd3_b200 import _lib
H, L = 256, 2
agent = rt.TD3(rt.Residual_Actor_Network(H, L), rt.Residual_Critic_Network(H, L), rt.Residual_Critic_Network(H, L))
agent.sync_transposed(); agent.precision = "tf32"; agent._sync_chunk_major(); agent._sync_half()
for B in (8192, 65536, 1 << 20):
    x = torch.rand((B, 2), device="cuda"); y = torch.empty((B, 2), device="cuda")
    Lb = _lib.lib(); sp = _lib.stream_ptr()
    def f32(): _lib.check(Lb.rtd3_mlp_forward(agent._handle, 0, _lib.ptr(agent.params), _lib.ptr(agent.params_t), _lib.ptr(x), _lib.ptr(y), B, sp))
    def t32(): _lib.check(Lb.rtd3_mlp_forward_tf32(H, L, 1, 0, _lib.ptr(agent.params), _lib.ptr(agent.params_u), _lib.ptr(x), _lib.ptr(y), B, sp))
    def f16(): _lib.check(Lb.rtd3_mlp_forward_f16(H, L, 0, _lib.ptr(agent.params), _lib.ptr(agent.params_h), _lib.ptr(x), _lib.ptr(y), B, sp))
    for name, fn in (("fp32", f32), ("tf32", t32), ("f16 ", f16)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        flops = 2.0 * B * (2 * H + H * H + 2 * H)
        print("B=%7d %s: %8.1f us  %.1f TFLOP/s" % (B, name, us, flops / us / 1e6))
