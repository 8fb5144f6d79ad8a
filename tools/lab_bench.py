"""Kernel-variant lab (development aid): times the headline pipeline pieces for the library named by FBGNN_LIB and
prints counters as an exactness checksum.   FBGNN_LIB=.../libfbgnn_expN.so python tools/lab_bench.py [B]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_bench import make
import fbgnn as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = F.default_context()
out = {"lib": os.path.basename(os.environ.get("FBGNN_LIB", "libfbgnn.so"))}
t = {}
for nG in (0, 1, 3):
    m = make("c1270", nG)
    for _ in range(2):
        m.run(B, 0.10, want_flags=False, want_diff=False)
    ctx.sync()
    ctx.timer_start()
    reps = 4
    for _ in range(reps):
        m.run(B, 0.10, want_flags=False, want_diff=False)
    t[nG] = ctx.timer_stop() / reps
    r = m.run(B, 0.10, want_counters=True)
    out[f"counters_nG{nG}"] = r["counters"].tolist()
out["us_per_frame_stage0"] = round(t[0] / B * 1e3, 4)
out["us_per_frame_gnn_plus_bp16"] = round((t[1] - t[0]) / B * 1e3, 4)
out["frames_per_s_nG3"] = round(B / (t[3] * 1e-3))
print(json.dumps(out))
