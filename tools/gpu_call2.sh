#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "forwarding" > gpurun_out/c2_tests.log 2>&1; tail -8 gpurun_out/c2_tests.log
timeout 300 python tools/fwd_sweep.py 20 0:0:3:11 1:0:3:11 1:16:3:11 1:8:3:11 1:24:3:11 1:3:3:11 1:11:3:11 1:27:3:11 1:7:3:11 1:15:3:11 1:31:3:11 1:0:2:11 1:24:2:11 1:11:2:11 1:27:2:11 0:0:2:11 1:0:3:12 1:16:3:12 1:3:3:12 > gpurun_out/c2_fwd20.jsonl 2> gpurun_out/c2_fwd.err; cat gpurun_out/c2_fwd20.jsonl; tail -3 gpurun_out/c2_fwd.err
timeout 200 python tools/fwd_sweep.py 22 0:0:3:11 1:0:3:11 1:27:3:11 1:0:3:12 1:3:3:12 > gpurun_out/c2_fwd22.jsonl 2>> gpurun_out/c2_fwd.err; cat gpurun_out/c2_fwd22.jsonl
timeout 100 python tools/fwd_sweep.py 18 0:0:3:11 1:0:3:11 1:27:3:11 1:11:3:11 > gpurun_out/c2_fwd18.jsonl 2>> gpurun_out/c2_fwd.err; cat gpurun_out/c2_fwd18.jsonl
