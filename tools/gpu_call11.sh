#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "1 0" "1 32" "0 64"; do set -- $cfg; for n in 20 24; do echo -n "N=$n STREAM=$1 DBG=$2: "; PB200_STREAM=$1 PB200_DBG=$2 timeout 100 python tools/apply_only.py $n 40 2>&1 | tail -1; done; done | tee gpurun_out/c11_sweep.txt
PB200_STREAM=1 PB200_DBG=32 timeout 200 python -m pytest tests -m gpu -q -k "apply_h_uniform or apply_h_local or blockade_sweep" 2>&1 | tail -2 | tee -a gpurun_out/c11_sweep.txt
PB200_DBG=64 timeout 200 python -m pytest tests -m gpu -q -k "apply_h_uniform or blockade_sweep" 2>&1 | tail -2 | tee -a gpurun_out/c11_sweep.txt
for cfg in "1 32" "0 64" "0 0"; do set -- $cfg; echo "quick_bench STREAM=$1 DBG=$2"; PB200_STREAM=$1 PB200_DBG=$2 timeout 100 python tools/fwd_sweep.py 20 0:0:3:11 2>&1 | tail -1; done | tee -a gpurun_out/c11_sweep.txt
