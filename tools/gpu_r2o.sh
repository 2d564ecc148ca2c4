#!/bin/bash
# memcheck / racecheck over the kernels added in round 2
mkdir -p gpurun_out
K="binned or ray_order or node_visit or packed_pipeline or formation_store or kat_vector or treelet_histogram"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/r2o_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2o_memcheck.log | tail -8
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "binned_variant_stages or ray_order or node_visit" > gpurun_out/r2o_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2o_racecheck.log | tail -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_baseline_scale.py -m gpu -x -q -k "stack_overflow or stack_entries" > gpurun_out/r2o_memcheck2.log 2>&1; echo "memcheck2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2o_memcheck2.log | tail -4
