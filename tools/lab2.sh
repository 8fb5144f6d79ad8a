python tools/dump_sfu_tables.py || exit 1
cp tests/golden/sfu_b200_*.xz gpurun_out/
timeout 900 python -m pytest tests/test_sfu_oracle.py tests/test_gpu_sfu.py tests/test_gpu_parity.py tests/test_gnn_bp4.py -x -q 2>&1 | tail -8
python tools/lab_bench.py
FBGNN_MATH=sfu python tools/lab_bench.py
