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
# the dominant kernel, full set: FULL=1 re-captures N = 20 / N = 24 (profiles/r02_taylor_stage_kernel*_summary.json)
if [ "${FULL:-0}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:stage_d2_taylor -s 2000 -c 4 \
    -o $O/${R}_prof_taylor_stage python tools/taylor_run.py 20 1 > $O/${R}_ncu_taylor_stage.log 2>&1
ncu --set full --clock-control none -k regex:stage_d2_taylor -s 2000 -c 2 -o $O/${R}_prof_taylor_stage_n24 python tools/taylor_run.py 24 1 > $O/${R}_ncu_taylor_stage_n24.log 2>&1
fi
# the batched (per-trajectory table) variant on a C4 device batch of 64 trajectories
PB200_BENCH_C4_TRAJ=64 ncu --set full --clock-control none -k regex:stage_d2_taylor -s 3000 -c 2 -o $O/${R}_prof_taylor_stage_c4 python tools/run_c4.py > $O/${R}_ncu_taylor_stage_c4.log 2>&1
# compute-sanitizer on the Taylor stage kernels (short sequences)
compute-sanitizer --tool memcheck python tools/sanitize_taylor.py > $O/${R}_compute_sanitizer_memcheck_taylor.log 2>&1; tail -3 $O/${R}_compute_sanitizer_memcheck_taylor.log
compute-sanitizer --tool racecheck python tools/sanitize_taylor.py > $O/${R}_compute_sanitizer_racecheck_taylor.log 2>&1; tail -3 $O/${R}_compute_sanitizer_racecheck_taylor.log
