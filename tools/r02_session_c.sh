set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; tail -5 gpurun_out/r02c_tests.log
python tools/mag4_ab.py > gpurun_out/r02c_mag4.jsonl 2>&1; cat gpurun_out/r02c_mag4.jsonl
PB200_BENCH_C4_TRAJ=64 python tools/run_c4.py > gpurun_out/r02c_c4.json 2>&1; cat gpurun_out/r02c_c4.json
python tools/run_c3.py 14 > gpurun_out/r02c_c3.jsonl 2>&1; cat gpurun_out/r02c_c3.jsonl
python tools/quick_bench.py 24 > gpurun_out/r02c_c5.jsonl 2>&1; cat gpurun_out/r02c_c5.jsonl
