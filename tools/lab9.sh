export FBGNN_MATH=sfu
python tools/lab_bench.py
FBGNN_BP4_SMEM_PAD=12000 python tools/lab_bench.py
FBGNN_BP4_SMEM_PAD=30000 python tools/lab_bench.py
