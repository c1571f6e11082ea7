#!/bin/bash
# One GPU call that produces the round's evidence:  gpurun --timeout 3000 -- 'bash tools/round_capture.sh r02'
R=${1:-r02}
O=gpurun_out
python -m pytest tests -m gpu -q -rs > $O/${R}_gpu_tests.log 2>&1; tail -4 $O/${R}_gpu_tests.log
python bench.py --steps 5 --warmup 3 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; tail -c 600 $O/${R}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2>> $O/${R}_bench_n1.err
# launch list of one bench step (C2 only): every launch with its device time (cold-cache, serialised: shares only)
PB200_BENCH_SKIP_C4=1 PB200_BENCH_SKIP_C5=1 PB200_BENCH_SKIP_C3=1 PB200_BENCH_SKIP_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 16000 -c 800 --csv \
    --log-file $O/${R}_launches.csv python bench.py --steps 1 --warmup 3 > /dev/null 2>&1
# the dominant kernel, full set, 3 launches from the middle of a step (sweep section: chi history, accumulator on even orders)
ncu --set full --clock-control none --import-source on -k regex:stage_d2_taylor -s 2000 -c 4 \
    -o $O/${R}_prof_taylor_stage python tools/taylor_run.py 20 1 > $O/${R}_ncu_taylor_stage.log 2>&1
# the same kernel at C5 size
ncu --set full --clock-control none -k regex:stage_d2_taylor -s 2000 -c 2 -o $O/${R}_prof_taylor_stage_n24 python tools/taylor_run.py 24 1 > $O/${R}_ncu_taylor_stage_n24.log 2>&1
