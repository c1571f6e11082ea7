set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; tail -5 gpurun_out/r02b_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; tail -c 3000 gpurun_out/r02b_bench_n1.json; tail -3 gpurun_out/r02b_bench_n1.err
python tools/l2hint_ab.py 20 22 24 > gpurun_out/r02b_l2hint.jsonl 2>&1; cat gpurun_out/r02b_l2hint.jsonl
PB200_BENCH_C4_TRAJ=64 ncu --set full --clock-control none --import-source on -k regex:stage_d2_rb_kernel -s 200 -c 2 -o gpurun_out/r02b_prof_c4 python tools/run_c4.py > gpurun_out/r02b_ncu_c4.log 2>&1; tail -2 gpurun_out/r02b_ncu_c4.log
for h in 0 3; do PB200_L2HINT=$h ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum --clock-control none -k regex:stage_d2_rb_kernel -s 5 -c 3 --csv --log-file gpurun_out/r02b_n24_hint$h.csv python tools/apply_only.py 24 > /dev/null 2>&1; done
compute-sanitizer --tool racecheck python tools/sanitize_c2.py > gpurun_out/r02b_racecheck.log 2>&1; tail -5 gpurun_out/r02b_racecheck.log
compute-sanitizer --tool memcheck python tools/sanitize_c2.py > gpurun_out/r02b_memcheck.log 2>&1; tail -5 gpurun_out/r02b_memcheck.log
