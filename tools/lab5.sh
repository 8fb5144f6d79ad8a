L=$PWD/feedback-gnn_b200/fbgnn
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu 2>&1 | tail -3
export FBGNN_MATH=sfu
python tools/lab_bench.py
for t in $LABTAGS; do FBGNN_LIB=$L/libfbgnn_$t.so python tools/lab_bench.py; done
unset FBGNN_MATH
python tools/lab_bench.py
