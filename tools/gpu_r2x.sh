#!/bin/bash
# round 2, session 2: one-pass scan -- full GPU suite, bench A/B against the previous commit's numbers, memcheck over the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f scan %.4f k3 %.3f step %.3f value %.1f M launches %d" % (b["k_traverse"], b["scan"], b["k_compact"], d["ms_per_step"], d["value"]/1e6, d["gpu_launches"]))'
B="python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for i in 1 2; do echo -n "bench: "; $B 2>/dev/null | python -c "$J"; done
K="kat or random_scenes or packed_pipeline or treelet_histogram or empty_and_ragged or formation_store"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/r2x_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2x_memcheck.log | tail -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kat or treelet_histogram" > gpurun_out/r2x_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r2x_racecheck.log | tail -5
