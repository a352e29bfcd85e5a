import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.td3_perf import run
import sys
B, H, L, E = [int(v) for v in sys.argv[1:5]]
run(B, H, L, epochs=E, reps=2, sampler=False)
