#!/bin/bash
# parity tests, bench breakdown, one ncu --set full capture of K1 (name given as $1)
bash tools/gpu_quick.sh
ncu --set full --clock-control none --import-source on -k regex:k_traverse -s 3 -c 1 -o gpurun_out/prof_traverse_$1 -f python bench.py --steps 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-200
