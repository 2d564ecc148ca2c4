#!/bin/bash
# parity tests, bench breakdown, one ncu --set full capture of K1 (name given as $1; $2 = "notests" skips the quick check)
# The EXACT pass is queued behind a device-side gate after every hot launch, so K1 is picked by its mangled name (Lb0 = EXACT false).
[ "$2" = "notests" ] || bash tools/gpu_quick.sh
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_traverseILi${VSRT_BENCH_MODE:-1}ELi96ELb0 -s 2 -c 1 -o gpurun_out/prof_traverse_$1 -f python bench.py --steps 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-200
