#!/bin/bash
# Refresh the round's profile artefacts under gpurun_out/ (copy what is wanted into profiles/).
#   profile_round.sh sfu|exact : ncu --set full capture of the hot kernels at the bench's B = 32768 in that arithmetic
#   profile_round.sh bench     : launch list of the bench command (durations only) + the bench line itself
case "$1" in
sfu)   FBGNN_LAB_GNN_GEMM=tf32x3 timeout 900 ncu --set full --clock-control none -k regex:'k_bp4|k_gnn' -s 3 -c 3 -o gpurun_out/r02_headline_sfu -f \
           python tools/prof_run_sfu.py 32768 1 > /dev/null 2>&1 ;;
exact) timeout 900 ncu --set full --clock-control none -k regex:'k_bp4|k_gnn' -s 3 -c 3 -o gpurun_out/r02_headline_exact -f \
           python tools/prof_run.py 32768 1 > /dev/null 2>&1 ;;
bench) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
           python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
       timeout 600 python bench.py --steps 10 2>&1 | tail -1 > gpurun_out/r02_bench_1gpu.json ;;
esac
ls -la gpurun_out/
