#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "forwarding" > gpurun_out/c1_tests.log 2>&1; tail -5 gpurun_out/c1_tests.log
timeout 240 python tools/fwd_sweep.py 16 18 20 21 22 > gpurun_out/c1_fwd.jsonl 2> gpurun_out/c1_fwd.err; cat gpurun_out/c1_fwd.jsonl; tail -3 gpurun_out/c1_fwd.err
for b in 8 16 32; do timeout 120 python tools/run_c4.py 64 $b 2>&1 | tail -1 | tee -a gpurun_out/c1_c4.jsonl; done
