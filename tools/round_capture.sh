#!/bin/bash
# One GPU call: the whole GPU test suite, the formal bench lines and the ncu captures that profiles/ is built from.
#   gpurun --timeout 1500 -- 'bash tools/round_capture.sh r01'
R=${1:-r01}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/${R}_tests.log 2>&1; tail -3 gpurun_out/${R}_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; tail -c 600 gpurun_out/${R}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/${R}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 0 > gpurun_out/${R}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_d2_rb -s 3000 -c 3 -f -o gpurun_out/${R}_prof_stage python bench.py --steps 1 --warmup 0 > gpurun_out/${R}_ncu2.log 2>&1
timeout 200 python tools/run_c4.py 64 64 > gpurun_out/${R}_c4.json 2>&1; tail -2 gpurun_out/${R}_c4.json
timeout 200 python tools/run_c3.py 14 > gpurun_out/${R}_c3.json 2>&1; tail -2 gpurun_out/${R}_c3.json
timeout 200 python tools/obs_timing.py > gpurun_out/${R}_obs.json 2>&1; tail -3 gpurun_out/${R}_obs.json
ls -la gpurun_out | tail -12
