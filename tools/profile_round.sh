#!/bin/bash
# Refresh the round's profile artefacts under gpurun_out/ (copy what is wanted into profiles/):
#   1. launch list of the bench command (durations only), 2. one --set full capture of the two hot kernels.
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames-per-step 8192 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_bp4|k_gnn' -s 4 -c 4 -o gpurun_out/headline_full -f \
    python tools/prof_run.py 2368 1 > /dev/null 2>&1
ls -la gpurun_out/
