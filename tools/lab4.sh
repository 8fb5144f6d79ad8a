L=$PWD/feedback-gnn_b200/fbgnn
export FBGNN_MATH=sfu
python tools/lab_bench.py
FBGNN_LIB=$L/libfbgnn_stable.so python tools/lab_bench.py
unset FBGNN_MATH
python tools/ler_check.py | tee gpurun_out/r02_ler_exact_vs_sfu.txt
FBGNN_LIB=$L/libfbgnn_stable.so python tools/ler_check.py | tee gpurun_out/r02_ler_exact_vs_sfu_stable.txt
python -m pytest tests/test_gpu_sfu.py -x -q -m gpu 2>&1 | tail -5
