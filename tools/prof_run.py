"""One small pipeline call for ncu captures (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_bench import make
import fbgnn as F
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
nG = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = make("c1270", nG)
for _ in range(2):
    r = m.run(B, 0.10, want_counters=True)
print(r["counters"])
