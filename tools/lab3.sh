L=$PWD/feedback-gnn_b200/fbgnn
export FBGNN_MATH=sfu
python tools/lab_bench.py
for t in gnn5 gnn6; do FBGNN_LIB=$L/libfbgnn_$t.so python tools/lab_bench.py; done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_gnn" -s 1 -c 1 -o gpurun_out/r02c_gnn_sfu -f python tools/prof_run_sfu.py 8192 1 2>&1 | tail -1
