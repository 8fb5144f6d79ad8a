"""One small pipeline call in SFU arithmetic for ncu captures (development aid)."""
import os, sys
os.environ["FBGNN_MATH"] = "sfu"
sys.argv = sys.argv[:1] + sys.argv[1:]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_run.py")).read())
