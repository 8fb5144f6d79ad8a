L=$PWD/feedback-gnn_b200/fbgnn
export FBGNN_MATH=sfu
python tools/lab_bench.py
for t in $LABTAGS; do FBGNN_LIB=$L/libfbgnn_$t.so python tools/lab_bench.py; done
