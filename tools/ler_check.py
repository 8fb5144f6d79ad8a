"""Lab: block-error counts of the n1270 nG=3 pipeline in exact and fast arithmetic on the same frames."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_bench import make
import fbgnn as F
ctx = F.default_context()
pts = [(0.13, 40000, 705/5000), (0.12, 40000, 139/5000), (0.11, 200000, 106/25000), (0.10, 600000, 100/275000)]
for p, frames, pub in pts:
    row = {"p": p, "frames": frames, "published": pub}
    for mode in ("exact", "fast"):
        ctx.set_math(mode)
        m = make("c1270", 3, skip=True); m.seed = 4242
        k = fl = s0 = 0
        for _ in range(frames // 20000):
            c = m.run(20000, p, want_flags=False, want_diff=False, want_counters=True)["counters"]
            k += int(c[2]); fl += int(c[1]); s0 += int(c[3])
        row[mode] = {"block": k, "rate": k / frames, "flagged": fl, "stage0_fail": s0}
    print(json.dumps(row), flush=True)
ctx.set_math("exact")
