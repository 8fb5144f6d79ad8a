L=feedback-gnn_b200/fbgnn
python tools/lab_bench.py
for t in g1l1 g2l1 g3l1 g6l1; do FBGNN_LIB=$PWD/$L/libfbgnn_$t.so python tools/lab_bench.py; done
for th in 128 192 320 384 512; do echo threads $th; FBGNN_BP4_THREADS=$th python tools/lab_bench.py; done
