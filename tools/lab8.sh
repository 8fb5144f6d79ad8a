export FBGNN_LAB_GNN_GEMM=tf32x3
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_gnn_tc" -s 1 -c 1 -o gpurun_out/r02_gnn_tc -f python tools/prof_run_sfu.py 32768 1 2>&1 | tail -1
