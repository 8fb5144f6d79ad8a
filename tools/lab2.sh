python tools/dump_sfu_tables.py
cp tests/golden/sfu_b200_*.xz gpurun_out/
python tools/lab_bench.py
FBGNN_MATH=sfu python tools/lab_bench.py
