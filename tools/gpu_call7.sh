#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -k "tiled or c3 or leakage or all_basis" > gpurun_out/c7_tests.log 2>&1; tail -4 gpurun_out/c7_tests.log
timeout 100 python tools/apply_c3.py 14 20 2>&1 | tail -1 | tee -a gpurun_out/c7_c3.jsonl
timeout 200 python tools/run_c3.py 14 2>&1 | tail -3 | tee -a gpurun_out/c7_c3.jsonl
