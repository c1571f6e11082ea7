#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -k "slm or xy or forwarding or final_state" > gpurun_out/c3_tests.log 2>&1; tail -8 gpurun_out/c3_tests.log
timeout 300 python tools/run_c3.py 14 > gpurun_out/c3_c3.jsonl 2>&1; cat gpurun_out/c3_c3.jsonl | tail -3
