#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "tiled or c3 or lindblad or leakage or all_basis or final_state" > gpurun_out/c4_tests.log 2>&1; tail -12 gpurun_out/c4_tests.log
for cfg in "1 0" "1 1" "0 0"; do set -- $cfg; echo "TILED=$1 BIG=$2"; PB200_TILED=$1 PB200_TILED_BIG=$2 timeout 200 python tools/run_c3.py 14 2>&1 | tail -2 | tee -a gpurun_out/c4_c3.jsonl; done
