"""Block-error counts of the n1270 nG=3 pipeline on the SAME frames in exact arithmetic, in SFU arithmetic and in SFU
arithmetic with the tensor-core feedback GNN, next to the published rates (examples/n1270.ipynb).
    python tools/ler_check.py [scale]      (scale multiplies the frame counts; default 1)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_bench import make
import fbgnn as F
ctx = F.default_context()
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
pts = [(0.13, 40000, 705/5000), (0.12, 40000, 139/5000), (0.11, 200000, 106/25000), (0.10, 600000, 100/275000)]
for p, frames, pub in pts:
    frames = int(frames * scale) // 20000 * 20000
    row = {"p": p, "frames": frames, "published": pub}
    for mode, math, gemm in (("exact", "exact", "fma"), ("fast", "sfu", "fma"), ("sfu_tc", "sfu", "tf32x3")):
        ctx.set_math(math)
        os.environ["FBGNN_LAB_GNN_GEMM"] = gemm
        m = make("c1270", 3, skip=True); m.seed = 4242
        k = fl = s0 = 0
        for _ in range(frames // 20000):
            c = m.run(20000, p, want_flags=False, want_diff=False, want_counters=True)["counters"]
            k += int(c[2]); fl += int(c[1]); s0 += int(c[3])
        row[mode] = {"block": k, "rate": k / frames, "flagged": fl, "stage0_fail": s0}
    print(json.dumps(row), flush=True)
ctx.set_math("exact")
