#!/bin/bash
# session: parity tests (staging off / on), staging sweep on bench, ncu capture of K1
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s1_tests_off.log
VSRT_STAGE_NODES=64 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s1_tests_on.log
VAR=VSRT_STAGE_NODES VALUES="0 64 256 1024" bash tools/sweep_env.sh > gpurun_out/s1_sweep_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_traverse -s 3 -c 1 -o gpurun_out/prof_traverse_r1d -f python bench.py --steps 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/s1_ncu.log 2>&1
cat gpurun_out/s1_tests_off.log gpurun_out/s1_tests_on.log gpurun_out/s1_sweep_stage.log
