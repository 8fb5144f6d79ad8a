#!/bin/bash
# compute-sanitizer passes over the round's kernels (both arithmetics); output -> gpurun_out/r02_sanitizer.txt
out=gpurun_out/r02_sanitizer.txt
: > $out
run() {   # tool, math, pytest args...
    tool=$1; shift; math=$1; shift
    echo "## compute-sanitizer --tool $tool   FBGNN_MATH=$math   pytest $*" >> $out
    FBGNN_MATH=$math timeout 1500 compute-sanitizer --tool $tool --error-exitcode 0 python -m pytest "$@" -q -m gpu -x 2>&1 \
        | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|hazard|Invalid|Error:|error" | sort | uniq -c | tail -12 >> $out
    echo >> $out
}
SEL='(steane or rsurf3 or gb48 or c882) and not published'
# exact arithmetic: the parity tests as they are; SFU arithmetic: tests/test_gpu_sfu.py (the same tests under its fixture)
run memcheck exact tests/test_gpu_parity.py tests/test_gpu_early_stop.py -k "$SEL"
run memcheck exact tests/test_gpu_sfu.py -k "bitexact_sfu and not c1270"
run racecheck exact tests/test_gpu_parity.py -k "(bp4_layer or pipeline_bitexact or pipeline_packed or bp2_layer) and (gb48 or c882 or rsurf3)"
run racecheck exact tests/test_gpu_sfu.py -k "bp4_layer_bitexact_sfu and (gb48 or c882)"
run synccheck exact tests/test_gpu_parity.py -k "(bp4_layer or pipeline_bitexact or pipeline_packed) and (gb48 or c882)"
run memcheck exact tests/test_gpu_parity.py -k "larger_than_shared_memory"
run memcheck exact tests/test_gpu_gnn_tc.py -k "tolerance or bitexact and not headline"
run racecheck exact tests/test_gpu_gnn_tc.py -k "tolerance and c882 and sum"
cat $out
