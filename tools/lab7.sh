export FBGNN_MATH=sfu
python tools/lab_bench.py
FBGNN_FPX_MIN_ITER=16 python tools/lab_bench.py
unset FBGNN_MATH
python tools/lab_bench.py
FBGNN_FPX_MIN_ITER=16 python tools/lab_bench.py
