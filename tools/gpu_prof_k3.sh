#!/bin/bash
# ncu --set full capture of K3 (name $1) and the launch list of one bench run
ncu --set full --clock-control none --import-source on -k regex:k_compact -s 3 -c 1 -o gpurun_out/prof_compact_$1 -f python bench.py --steps 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_k3_$1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_$1.log 2>&1
tail -1 gpurun_out/ncu_k3_$1.log | cut -c1-200
