#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "tiled or effective_size or slm" > gpurun_out/c5_tests.log 2>&1; tail -4 gpurun_out/c5_tests.log
for k in 6 7 8; do PB200_TILED_K=$k timeout 100 python tools/apply_c3.py 14 20 2>&1 | tail -1 | tee -a gpurun_out/c5_c3.jsonl; done
PB200_TILED=0 timeout 100 python tools/apply_c3.py 14 20 2>&1 | tail -1 | tee -a gpurun_out/c5_c3.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_tiled -s 2 -c 2 -f -o gpurun_out/c5_prof_tiled python tools/apply_c3.py 14 6 > gpurun_out/c5_ncu.log 2>&1; tail -2 gpurun_out/c5_ncu.log
ls -la gpurun_out | tail -5
