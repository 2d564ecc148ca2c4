#!/bin/bash
# round 2, session 2: shared-memory carve-out of K3 (driver default 132 KB for 92 KB used)
mkdir -p gpurun_out
J='import json,sys
d=json.loads(sys.stdin.read()); b=d["roofline"]["step_breakdown_ms"]; print("k1 %.3f k3 %.3f step %.3f value %.1f M" % (b["k_traverse"], b["k_compact"], d["ms_per_step"], d["value"]/1e6))'
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-incoherent --e2e-steps 1"
for CV in -1 auto 57 -1 auto; do
  if [ $CV = auto ]; then unset VSRT_K3_CARVEOUT; else export VSRT_K3_CARVEOUT=$CV; fi
  echo -n "bench k3 carveout=$CV: "; $B 2>/dev/null | python -c "$J"
done
unset VSRT_K3_CARVEOUT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2ad_tests.log
