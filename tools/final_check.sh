timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gnn_bp4.py tests/test_training.py -m gpu -x -q > gpurun_out/sanitizer_new.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_new.txt
