"""Debug aid: compare the tf32 learner's gradients with the fp32 path tensor by tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import rtd3_b200 as pkg
from test_tc_learner_gpu import make, critic_pass, actor_pass, net_slices

H, B = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
agent, rb, idx, noise = make(pkg, H, B)
agent.tc_min_batch = 1
l_ref, q_ref, y_ref, g_ref = critic_pass(agent, rb, idx, noise)
agent.precision = "tf32"
l_tc, q_tc, y_tc, g_tc = critic_pass(agent, rb, idx, noise)
print("loss", l_ref, l_tc)
for net in (1, 2):
    for name, sl in net_slices(agent, net):
        a, b = g_tc[sl].double(), g_ref[sl].double()
        print(net, name, "ref max %.4g tc max %.4g err %.4g nonzero_tc %d/%d" % (b.abs().max(), a.abs().max(), (a - b).abs().max(), int((a != 0).sum()), a.numel()))
        if name == "W1":
            A, Bm = a.view(H, H), b.view(H, H)
            print("   vs transpose err %.4g" % (A.t() - Bm).abs().max())
            print("   ratio sample", (A[:4, :4] / Bm[:4, :4]).cpu().numpy())
            blk = (A - Bm).abs().view(H // 32, 32, H // 32, 32).amax(dim=(1, 3))
            print("   per 32x32 block max err:\n", blk.cpu().numpy().round(1))
la_ref, ga_ref = None, None
agent.precision = "fp32"
la_ref, ga_ref = actor_pass(agent, rb, idx)
agent.precision = "tf32"
la_tc, ga_tc = actor_pass(agent, rb, idx)
print("actor loss", la_ref, la_tc)
for name, sl in net_slices(agent, 0):
    a, b = ga_tc[sl].double(), ga_ref[sl].double()
    print(0, name, "ref max %.4g tc max %.4g err %.4g" % (b.abs().max(), a.abs().max(), (a - b).abs().max()))
