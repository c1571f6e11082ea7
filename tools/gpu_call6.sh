#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -k "tiled or c3 or lindblad or leakage or all_basis" > gpurun_out/c6_tests.log 2>&1; tail -6 gpurun_out/c6_tests.log
for t in 1 2 0; do PB200_TILED=$t timeout 100 python tools/apply_c3.py 14 20 2>&1 | tail -1 | tee -a gpurun_out/c6_c3.jsonl; done
timeout 200 python tools/run_c3.py 14 2>&1 | tail -2 | tee -a gpurun_out/c6_c3.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_multilevel -s 2 -c 2 -f -o gpurun_out/c6_prof_mlrb python tools/apply_c3.py 14 6 > gpurun_out/c6_ncu.log 2>&1; tail -2 gpurun_out/c6_ncu.log
