#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/c8_c4_launches.csv python tools/run_c4.py 64 64 > gpurun_out/c8_ncu.log 2>&1; tail -2 gpurun_out/c8_ncu.log
python tools/ncu_summary.py launches gpurun_out/c8_c4_launches.csv gpurun_out/c8_c4_launches_summary.json | head -60
